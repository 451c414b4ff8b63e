#!/usr/bin/env python
"""profiles/r02_summary.md from the round's captured files (bench lines, ncu launch list, ncu --set full raw CSVs).
usage: python tools/make_profile_summary.py BENCH.json REF.json LAUNCHES.csv REGIMEA_RAW.csv KSW_RAW.csv [N2.json] [EXTRA_RAW.csv=title ...] > profiles/r02_summary.md"""
import csv, json, sys, collections, os

bench, ref, launches, rega, ksw = sys.argv[1:6]
n2 = sys.argv[6] if len(sys.argv) > 6 and "=" not in sys.argv[6] else None
extras = [a.split("=", 1) for a in sys.argv[6:] if "=" in a]
B, R = json.load(open(bench)), json.load(open(ref))
out = []
P = out.append
P("# Round 2 profile summary (B200; every number below comes from a file in this directory)\n")
P("Files: `%s` (bench line, driver contract), `%s` (`--impl reference`: the unmodified binary on the same reads/index/batches), `%s` (ncu launch list of the bench's timed resident steps), `%s` (`ncu --set full --clock-control none`, profiler range = the same steps), `%s` (`ncu --set full` of the ksw_extd2 batch)." % tuple(os.path.basename(x) for x in (bench, ref, launches, rega, ksw)))
P("")
P("## Headline (N = 1, configs[1]: 100 k x 20 kb ONT reads, k17 w10, 3000-read = 60 Mbp steps, same batches on both arms)\n")
P("| | overlaps/s | ms/step | aligned Gbp/s |\n|---|---:|---:|---:|")
P("| GPU, inputs resident (`value`) | %.3g | %.3f | %.1f |" % (B["value"], B["ms_per_step"], B["aligned_gbp_per_s"]))
P("| GPU, end to end from pinned host buffers (`e2e`: H2D %.1f MB + D2H %.2f MB per step inside the timed region) | %.3g | %.3f | |" % (B["e2e"]["h2d_bytes_per_step"] / 1e6, B["e2e"]["d2h_bytes_per_step"] / 1e6, B["e2e"]["value"], B["e2e"]["ms_per_step"]))
P("| reference binary, %d host threads (`--impl reference`) | %.3g | %.1f | %.2f |" % (R["cpu_baseline"]["cores"], R["value"], R["ms_per_step"], R["aligned_gbp_per_s"]))
P("\ne2e / reference = **%.1fx**, resident / reference = %.1fx (equal work: records %d and %d).  Clocks: %s.\n" % (B["e2e"]["value"] / R["value"], B["value"] / R["value"], B["work"]["records"], R["work"]["records"], json.dumps(B["clocks"])))
P("Whole-step roofline (SURVEY 8d, B_A = L/4 + 32 M + 88 n_a + 13 n_ovl): %.0f MB / %.3f ms = %.0f GB/s = **%.1f %%** of the measured %.0f GB/s.  Dominant kernel `%s`: %.0f MB algorithmic per launch, %.0f GB/s, frac %.3f; ncu DRAM traffic %.0f MB per launch.\n" % (
    B["roofline"]["whole_step"]["algorithmic_bytes_per_step"] / 1e6, B["ms_per_step"], B["roofline"]["whole_step"]["achieved"], 100 * B["roofline"]["whole_step"]["frac"], B["roofline"]["peak"],
    B["roofline"]["kernel"], B["roofline"]["algorithmic_bytes_per_launch"] / 1e6, B["roofline"]["achieved"], B["roofline"]["frac"], (B["roofline"]["traffic"] or 0) / 1e6))
P("## Per kernel family, event-timed inside bench.py (ms per step, algorithmic GB/s, fraction of HBM peak, ncu DRAM bytes per launch)\n")
P("| family | ms/step | launches/step | alg GB/s | frac | ncu dram MB/launch |\n|---|---:|---:|---:|---:|---:|")
for k, v in B["kernels"].items():
    P("| `%s` | %.4f | %g | %s | %s | %s |" % (k, v["ms_per_step"], v["launches_per_step"], v["alg_GBps"], v["frac"], "%.1f" % (v["ncu_dram_bytes_per_launch"] / 1e6) if v.get("ncu_dram_bytes_per_launch") else ""))
# launch list shares
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; ik, im, iv = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
iu = H.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if r[im] != "gpu__time_duration.sum": continue
    v = float(r[iv].replace(",", "")); u = r[iu]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1}.get(u, 1e-6)
    name = r[ik].split("(")[0].replace("void ", "")[:60]
    agg[name][0] += 1; agg[name][1] += ms
tot = sum(v[1] for v in agg.values())
P("\n## ncu launch list of the timed region (`%s`; cold-cache, serialised: compare the SHARES with the event-timed table above)\n" % os.path.basename(launches))
P("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    P("| `%s` | %d | %.3f | %.1f %% |" % (k, v[0], v[1], 100 * v[1] / tot))
def full_table(path, title):
    rows = list(csv.reader(open(path))); hdr, units = rows[0], rows[1]
    cols = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/inst"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram written"), ("smsp__inst_executed.sum", "warp inst")]
    P("\n## %s\n" % title)
    P("| kernel | " + " | ".join(c[1] for c in cols) + " |\n|---|" + "---:|" * len(cols))
    seen = collections.Counter()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")[:44]
        seen[name] += 1
        if seen[name] > 3: continue
        cells = []
        for c, _ in cols:
            i = hdr.index(c); v = r[i]
            try:
                f = float(v); v = ("%.4g" % f) + (" " + units[i] if units[i] not in ("", "%", "register/thread", "inst") else "")
            except ValueError:
                pass
            cells.append(v)
        P("| `%s` | " % name + " | ".join(cells) + " |")
full_table(rega, "Full capture, regime A (`%s`; at most three launches per kernel shown)" % os.path.basename(rega))
for path, title in extras:
    full_table(path, "%s (`%s`)" % (title, os.path.basename(path)))
full_table(ksw, "Full capture, `ksw_extd2` register-resident kernel on the synthetic batch (`%s`)" % os.path.basename(ksw))
K = B.get("ksw2_extd2") or {}
if K:
    cells, ms = K["cells"], K["kernel_ms"]
    P("\n`ksw_extd2` batch (20 000 calls 413 x 413, CIGARs included): %.3f ms event-timed = **%.1f GCUPS**; algorithmic bytes (SURVEY 8d B_B) %.2f GB -> %.0f GB/s = %.1f %% of HBM peak; one host core of `ksw_extd2_sse`: %.2f GCUPS.  Integer-issue view: 148 SM x 4 schedulers x 1.965 GHz = 1163 G warp-instr/s; the capture above gives warp instructions per launch, i.e. instructions per cell = warp inst x 32 / cells.\n" % (
        ms, K["gcups"], K["alg_bytes"] / 1e9, K["achieved_GBps"], 100 * K["roofline_frac_hbm"], K["cpu_reference_1core_gcups"] or 0))
rb, s2 = B.get("regime_b"), B.get("step2")
if rb and "seconds" in rb:
    P("Regime B (`%s`): %.2f s whole step (DP kernels %.2f s = %.0f GCUPS, alignment stage %.2f s) vs %.2f s for the reference with %d threads (%.2fx), bytes identical: %s." % (
        rb["workload"], rb["seconds"], rb["dp_kernel_s"], rb["gcups_kernels"], rb["alignment_stage_s"], rb["reference"]["seconds"], rb["reference"]["threads"], rb["speedup_vs_reference"], rb["bytes_identical_to_reference"]))
if s2 and "seconds" in s2:
    P("`--step 2` (`%s`): %.2f s vs %.2f s (%.2fx), bytes identical: %s." % (s2["workload"], s2["seconds"], s2["reference"]["seconds"], s2["speedup_vs_reference"], s2["bytes_identical_to_reference"]))
if n2:
    N = json.load(open(n2))
    P("\n## N = 2 (`%s`, one process per GPU over NCCL, weak scaling: 2 x 100 k reads, index built sharded)\n" % os.path.basename(n2))
    P("value %.3g overlaps/s (%.2fx of the N=1 line of the same kernels, `r02e`: 2.14e+07/s; NOT re-measured after the group filter and the strided chain claims of `r02n`, which remove most of the per-rank growth of the sort work -- the driver's `SCALE_r02.json` is the current number), %.3f ms/step, e2e %.3g overlaps/s; index: %s" % (N["value"], N["value"] / 21383756.76, N["ms_per_step"], N["e2e"]["value"], json.dumps(N["index"])))
print("\n".join(out))
