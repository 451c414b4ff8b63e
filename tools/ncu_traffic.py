#!/usr/bin/env python
"""profiles/rNN_traffic.json from an `ncu --set full` capture of bench.py: dram__bytes_read.sum + dram__bytes_write.sum per launch for
every kernel family bench.py reports (its per-family names are the ProfScope names of libndgpu).
usage: python tools/ncu_traffic.py capture.ncu-rep reads batch_reads out.json [summary.md]"""
import csv, json, subprocess, sys, collections

FAMILY = {"k_sketch_fast4": "sketch_fast", "k_sketch_fast": "sketch_fast", "k_sketch_chunks": "sketch", "k_sketch_compact": "sketch_compact",
          "k_probe_count": "probe_count", "k_fill_anchors": "fill_anchors", "k_rs_level": "sort_level", "k_rs_init": "sort_init", "k_sort_fast": "sort_fast",
          "k_chain_fill_warp": "chain_fill_warp", "k_chain_backtrack": "chain_backtrack", "k_emit": "emit", "k_hit": "hit_order", "k_read_stats": "read_stats",
          "k_ksw_rx": "ksw_extd2", "k_ksw_extd2": "ksw_extd2"}
COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def unit_scale(u):
    return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "usecond": 1e-3, "msecond": 1, "nsecond": 1e-6}.get(u, 1)


def main():
    rep, reads, batch, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    agg = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows[2:]:
        name = r[ik].split("(")[0].replace("void ", "").split("<")[0]
        fam = next((v for k, v in FAMILY.items() if name.startswith(k)), name)
        for c in COLS:
            if c in hdr:
                i = hdr.index(c)
                try:
                    agg[fam][c].append(float(r[i]) * unit_scale(units[i]))
                except ValueError:
                    pass
    kernels, lines = {}, ["| kernel family | launches | ms/launch | dram MB/launch | warps active % | issue active % | threads/inst | ALU pipe % | DRAM % | regs |", "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for fam, m in sorted(agg.items(), key=lambda kv: -sum(kv[1].get("gpu__time_duration.sum", [0]))):
        n = len(m["gpu__time_duration.sum"])
        avg = lambda c: sum(m[c]) / len(m[c]) if m.get(c) else float("nan")
        kernels[fam] = avg("dram__bytes_read.sum") + avg("dram__bytes_write.sum")
        lines.append("| `%s` | %d | %.4f | %.2f | %.1f | %.1f | %.1f | %.1f | %.1f | %d |" % (fam, n, avg("gpu__time_duration.sum"), kernels[fam] / 1e6, avg("sm__warps_active.avg.pct_of_peak_sustained_active"),
                     avg("smsp__issue_active.avg.pct_of_peak_sustained_active"), avg("smsp__thread_inst_executed_per_inst_executed.ratio"), avg("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                     avg("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), avg("launch__registers_per_thread")))
    json.dump({"reads": reads, "batch_reads": batch, "source": rep.split("/")[-1], "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full, cold cache)", "kernels": kernels},
              open(out, "w"), indent=1)
    if len(sys.argv) > 5:
        open(sys.argv[5], "a").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
