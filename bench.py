#!/usr/bin/env python
"""bench.py -- overlaps/s (+ aligned Gbp/s) of the read-overlap hot path on synthetic 20 kb ONT reads.

Workload = BASELINE.json configs[1]: 100k synthetic 20 kb ONT reads (15 % error: 5 % sub, 4 % ins,
6 % del), `-x ava-ont -k 17 -w 10`, self-overlap, .ovl step-1 records.  The index over all reads is
built once (reported separately); a *step* maps one query batch of `--batch-reads` reads against it.

  value : whole-job overlaps/s with the step's query batch already resident in HBM (kernels only,
          CUDA events on the launching stream, max over ranks)
  e2e   : the same through ndg_map_batch() with HOST buffers: H2D of the packed reads, all kernels,
          D2H of the hits and the host-side .ovl encoding inside the timed region
  roofline : dominant kernel of the step (largest accumulated device time), algorithmic bytes per
          DESIGN.md / SURVEY.md 8(d) divided by its event-timed duration, against MEASURED_PEAKS.json
  cpu_baseline : the unmodified reference (oracle/_ref/minimap2-nd, SSE path) on the box's host
          cores, on a bounded sample of the same workload (same read model and depth, smaller genome)

`--impl reference` times that reference binary as the measured arm.
N > 1: one process per GPU (torchrun), NCCL all-gather of the target sketches at start-up, then
independent per-GPU query batches (no data-path collective): weak scaling.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20240923
READ_LEN = 20000
ERR = dict(sub=0.05, ins=0.04, dele=0.06)
DEPTH = 40


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=100000, help="reads in the index (configs[1]: 100k)")
    ap.add_argument("--batch-reads", type=int, default=3000, help="query reads per step")
    ap.add_argument("--sub-batch-mbp", type=int, default=64, help="device sub-batch limit in Mbp (a step must fit one sub-batch)")
    ap.add_argument("--cpu-sample-reads", type=int, default=20000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", default="1,1")
    ap.add_argument("--no-ksw", action="store_true", help="skip the regime-B (ksw_extd2) kernel measurement")
    ap.add_argument("--no-regime-b", action="store_true", help="skip the regime-B (-c) whole-step measurement")
    ap.add_argument("--regime-b-full", action="store_true", help="regime B on all 997 reads of the reference's test data (GPU ~1 min, reference ~4 min)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu=0):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except Exception:  # noqa: BLE001
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def make_reads(n, genome_bp, first=0):
    from nextdenovo_b200 import synth
    g = synth.genome(genome_bp, SEED)
    return synth.reads(g, n, READ_LEN, seed=SEED, first=first, **ERR)


def run_reference(reads, threads):
    """Run the unmodified reference binary on `reads` (self-overlap, config-1 options).  Returns
    (records, sum_qspan, mapping_seconds, total_seconds)."""
    import nextdenovo_b200 as nd
    exe = os.path.join(ROOT, "oracle", "_ref", "minimap2-nd")
    if not os.path.exists(exe):
        raise RuntimeError("oracle/_ref/minimap2-nd missing (build() makes it when /root/reference is present)")
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "s.2bit")
        nd.write_2bit(p, reads)
        out = os.path.join(td, "s.ovl")
        t0 = time.time()
        r = subprocess.run([exe, "--step", "1", "-t", str(threads), "-I", "100G", "-x", "ava-ont", "-k", "17", "-w", "10", p, p, "-o", out],
                           stderr=subprocess.PIPE, text=True, check=True)
        total = time.time() - t0
        t_idx = t_map = None
        for line in r.stderr.splitlines():
            m = re.match(r"\[M::mm_idx_stat::([0-9.]+)\*", line)
            if m:
                t_idx = float(m.group(1))
            m = re.match(r"\[M::worker_pipeline::([0-9.]+)\*", line)
            if m:
                t_map = float(m.group(1))
        d = nd.decode_ovl8(open(out, "rb").read())
        recs = d.shape[0]
        sq = int((d[:, 3].astype(np.int64) - d[:, 2].astype(np.int64)).sum())
    return recs, sq, (t_map - t_idx) if (t_idx is not None and t_map is not None) else total, total


def usable_cpus():
    """CPUs this process may really use: affinity mask capped by the cgroup CPU quota (same rule as libndgpu's ndg_set_threads)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        if os.path.exists("/sys/fs/cgroup/cpu.max"):
            q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
            if q != "max":
                n = min(n, max(1, -(-int(q) // int(per))))
        else:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0 and per > 0:
                n = min(n, max(1, -(-q // per)))
    except Exception:  # noqa: BLE001
        pass
    return n


def regime_b_bench(eng, engine_mod, with_reference=True, full=False):
    """Regime B (`-c`, base-level alignment of every chain) on real reads: tests/golden/sub120.2bit (120 reads of the reference's
    test_data, 7346 overlaps; --regime-b-full: all 997 reads, 133100 overlaps) self-overlap with -x ava-ont.  Whole step through the
    public API (index build + mapping + alignment + .ovl bytes), checked against the golden bytes, next to the unmodified reference
    binary on the host cores for the same input."""
    import hashlib
    import nextdenovo_b200 as nd
    from nextdenovo_b200.options import MM_F_CIGAR
    g = os.path.join(ROOT, "tests", "golden")
    src = os.path.join(g, "reads_test.seed.2bit" if full else "sub120.2bit")
    gold = os.path.join(g, "G4_ava-ont_c.ovl" if full else "S6_sub120_ava-ont_c.ovl")
    reads = nd.read_2bit(src)
    io, mo = nd.IdxOpt(), nd.MapOpt()
    nd.set_opt("ava-ont", io, mo)
    mo.step, mo.minlen = 1, 500
    mo.flag |= MM_F_CIGAR
    eng.set_limits(64 << 20, 0)
    thr = eng.set_threads(0)
    out = {"workload": "%s self-overlap, --step 1 -c -x ava-ont (%d reads)" % (os.path.basename(src), reads.n), "host_threads": thr}
    for rep in range(2):                                    # first pass warms allocations (arena, pinned buffers)
        eng.prof_reset()
        t0 = time.time()
        ovl, st = engine_mod.run_step1(eng, reads, reads, io, mo)
        dt = time.time() - t0
    s = eng.aln_stats()
    out.update({"seconds": dt, "records": int(st["records"]), "overlaps_per_s": st["records"] / dt, "bytes_identical_to_reference": ovl == open(gold, "rb").read(),
                "md5": hashlib.md5(ovl).hexdigest(), "chains": s["n_chains"], "dp_calls": s["n_calls"], "rounds": s["n_rounds"], "launches": s["n_launches"],
                "cells": s["p_bytes"], "dp_kernel_s": s["dp_ms"] / 1e3, "dp_backend_wall_s": s["dp_wall_ms"] / 1e3, "alignment_stage_s": s["host_ms"] / 1e3,
                "gcups_kernels": s["p_bytes"] / max(s["dp_ms"], 1e-9) / 1e6, "gcups_stage": s["p_bytes"] / max(s["host_ms"], 1e-9) / 1e6})
    mo.flag &= ~MM_F_CIGAR
    eng.set_options(io, mo)
    if with_reference:
        exe = os.path.join(ROOT, "oracle", "_ref", "minimap2-nd")
        with tempfile.TemporaryDirectory() as td:
            o = os.path.join(td, "r.ovl")
            t0 = time.time()
            subprocess.run([exe, "--step", "1", "-c", "-t", str(usable_cpus()), "-x", "ava-ont", src, src, "-o", o], stderr=subprocess.DEVNULL, check=True)
            rt = time.time() - t0
            same = open(o, "rb").read() == ovl
        out["reference"] = {"seconds": rt, "threads": usable_cpus(), "same_bytes": same, "overlaps_per_s": st["records"] / rt}
        out["speedup_vs_reference"] = rt / dt
    return out


def step2_bench(eng, engine_mod, with_reference=True):
    """`--step 2` (cns_align) on the reference's test data (997 reads, self-overlap, the pipeline's options -k 17 -w 17 --minlen 2000
    --maxhan1 5000): whole step through the public API, .ovl and .bl checked against the golden bytes, next to the reference binary."""
    import nextdenovo_b200 as nd
    from nextdenovo_b200.options import parse_args
    g = os.path.join(ROOT, "tests", "golden")
    src = os.path.join(g, "reads_test.seed.2bit")
    reads = nd.read_2bit(src)
    args = ["-I", "6G", "--step", "2", "-x", "ava-ont", "-k", "17", "-w", "17", "--minlen", "2000", "--maxhan1", "5000"]
    io, mo, _, _ = parse_args(args + ["a", "b", "-o", "x"])
    eng.set_limits(64 << 20, 0)
    for rep in range(2):
        t0 = time.time()
        ovl, bl, st = engine_mod.run_step2(eng, reads, reads, io, mo)
        dt = time.time() - t0
    out = {"workload": "reads_test.seed.2bit self-overlap, " + " ".join(args), "seconds": dt, "records": st["records"],
           "bytes_identical_to_reference": ovl == open(os.path.join(g, "G5_step2.ovl"), "rb").read() and bl == open(os.path.join(g, "G5_step2.ovl.bl"), "rb").read()}
    if with_reference:
        exe = os.path.join(ROOT, "oracle", "_ref", "minimap2-nd")
        with tempfile.TemporaryDirectory() as td:
            o = os.path.join(td, "r.ovl")
            t0 = time.time()
            subprocess.run([exe] + args + ["-t", str(usable_cpus()), src, src, "-o", o], stderr=subprocess.DEVNULL, check=True)
            rt = time.time() - t0
        out["reference"] = {"seconds": rt, "threads": usable_cpus()}
        out["speedup_vs_reference"] = rt / dt
    return out


def cpu_sample(n):
    genome_bp = max(int(n * READ_LEN / DEPTH), 200000)
    return make_reads(n, genome_bp), "%d reads x %d bp at %dx depth (genome %d bp), self-overlap, mapping phase only" % (n, READ_LEN, DEPTH, genome_bp)


def ksw_cells(qlen, tlen, w):
    """in-band DP cells of one ksw_extd2 call (ref band rule: minimap2/ksw2_extd2_sse.c:128-133)"""
    r = np.arange(qlen + tlen - 1)
    st = np.maximum(np.maximum(0, r - qlen + 1), (r - w + 1) >> 1)
    en = np.minimum(np.minimum(tlen - 1, r), (r + w) >> 1)
    return int(np.maximum(en - st + 1, 0).sum())


def ksw_bench(eng, n_calls=20000, L=413, seed=SEED):
    """Regime B kernel alone: a batch of gap-filling style ksw_extd2 calls (mean 413 x 413, the mean call of
    `--step 1 -c` on reads_test), flag = KSW_EZ_APPROX_MAX, w = 3001, zdrop = 400, with CIGAR."""
    from nextdenovo_b200.engine import KSW_CALL_DTYPE
    rng = np.random.default_rng(seed)
    mat = np.zeros(25, dtype=np.int8)
    for i in range(4):
        for j in range(4):
            mat[i * 5 + j] = 2 if i == j else -4
        mat[i * 5 + 4] = -1
    mat[20:25] = -1
    calls = np.zeros(n_calls, dtype=KSW_CALL_DTYPE)
    chunks, off, cells, alg_bytes = [], 0, 0, 0
    cache = {}
    for i in range(n_calls):
        t = rng.integers(0, 4, L).astype(np.uint8)
        u = rng.random(L)
        qv = np.where((u >= 0.06) & (u < 0.11), (t + 1 + rng.integers(0, 3, L)) % 4, t).astype(np.uint8)[u >= 0.06]
        ins = np.nonzero(rng.random(qv.shape[0]) < 0.04)[0]
        qv = np.insert(qv, ins, rng.integers(0, 4, ins.shape[0]).astype(np.uint8))
        calls[i] = (len(qv), L, off, off + len(qv), 3001, 400, -1, 0x08)
        chunks += [qv, t]; off += len(qv) + L
        key = (len(qv), L)
        if key not in cache:
            cache[key] = ksw_cells(len(qv), L, 3001)
        cells += cache[key]
    seqs = np.concatenate(chunks)
    eng.ksw_extd2_batch(calls[:2000], seqs, mat, 4, 2, 24, 1)           # warm-up (allocations)
    eng.ksw_extd2_batch(calls, seqs, mat, 4, 2, 24, 1)
    eng.prof_reset()
    t0 = time.perf_counter()
    res, cigs = eng.ksw_extd2_batch(calls, seqs, mat, 4, 2, 24, 1)
    wall = time.perf_counter() - t0
    ms = eng.prof()["ksw_extd2"][0]
    n_cig = int(res["n_cigar"].sum())
    alg_bytes = int(2 * (calls["qlen"].sum() + calls["tlen"].sum()) + cells + 4 * n_cig)   # SURVEY 8(d) B_B
    out = {"calls": n_calls, "cells": cells, "kernel_ms": ms, "gcups": cells / ms / 1e6, "e2e_gcups": cells / wall / 1e9,
           "alg_bytes": alg_bytes, "achieved_GBps": alg_bytes / ms / 1e6}
    # reference on one host core, bounded sample
    try:
        import ctypes as C
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from ksw_cases import ref_lib, ref_call
        R = ref_lib()
        if R is not None:
            ns = 300
            t1 = time.perf_counter()
            csum = 0
            for i in range(ns):
                qo, to = int(calls["q_off"][i]), int(calls["t_off"][i])
                qs, ts = seqs[qo:qo + int(calls["qlen"][i])].copy(), seqs[to:to + L].copy()
                ref_call(R, qs, ts, mat, 4, 2, 24, 1, 3001, 400, -1, 0x08)
                csum += cache[(len(qs), L)]
            dt = time.perf_counter() - t1
            out["cpu_reference_1core_gcups"] = csum / dt / 1e9
            out["cpu_sample"] = "%d calls of the same batch through ksw_extd2_sse (SSE4.1), 1 thread" % ns
    except Exception as ex:  # noqa: BLE001
        out["cpu_reference_1core_gcups"] = None
        out["cpu_sample"] = "failed: %r" % ex
    return out


def main_reference(a, rank, world):
    if rank != 0:
        return
    threads = usable_cpus()
    reads, desc = cpu_sample(a.cpu_sample_reads)
    rec = sq = 0
    tsum = 0.0
    for i in range(a.warmup + a.steps):
        r, s, tm, _ = run_reference(reads, threads)
        if i >= a.warmup:
            rec += r; sq += s; tsum += tm
    v = rec / tsum
    line = {"metric": "overlaps/sec", "value": v, "unit": "overlaps/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * tsum / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "impl": "reference",
            "aligned_gbp_per_s": sq / tsum / 1e9,
            "config": {"workload": "configs[1] sample: " + desc, "options": "--step 1 -x ava-ont -k 17 -w 10"},
            "cpu_baseline": {"value": v, "unit": "overlaps/s", "cores": threads, "kind": "reference", "sample": desc},
            "e2e": {"value": v, "unit": "overlaps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main_ours(a, rank, world, local_rank):
    import torch
    import nextdenovo_b200 as nd
    from nextdenovo_b200 import engine
    dist = None
    saved_stdout = os.dup(1)
    os.dup2(2, 1)                                       # NCCL / torchrun banners go to stderr; stdout carries the JSON line only
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank
    hbm_peak, peak_kind = peaks()

    # ---- workload ---------------------------------------------------------------------------
    genome_bp = int(a.reads * READ_LEN / DEPTH)
    t0 = time.time()
    reads = make_reads(a.reads, genome_bp)
    t_gen = time.time() - t0
    io, mo = nd.IdxOpt(), nd.MapOpt()
    nd.set_opt("ava-ont", io, mo)
    io.k, io.w = 17, 10
    mo.step = 1
    eng = engine.Engine(dev)
    eng.set_options(io, mo)
    sv, cv = [int(x) for x in a.variant.split(",")]
    eng.set_variant(sv, cv)

    # ---- index: each rank sketches 1/world of the targets, NCCL all-gather, identical table everywhere
    torch.cuda.synchronize(dev)
    t0 = time.time()
    if world == 1:
        stat = eng.build_index(reads)
    else:
        stat = engine.build_index_sharded(eng, reads, rank, world, dist, torch)
    torch.cuda.synchronize(dev)
    t_index = time.time() - t0

    # ---- query shards (no data-path collective).  Weak scaling needs every rank to do statistically identical
    # work; with MM_F_NO_DUAL the number of records a read yields depends on where its *name* sorts (strcmp on
    # decimal strings), so contiguous id ranges are not comparable.  Each rank therefore maps its own copy of the
    # same K batches (rank r starts r batches further into the list, so ranks never work on the same batch at the
    # same time).
    nb = a.warmup + a.steps
    B = a.batch_reads
    eng.set_limits(a.sub_batch_mbp << 20, 0)
    base_batches = [reads.slice((i * B) % max(reads.n - B, 1), (i * B) % max(reads.n - B, 1) + B) for i in range(nb)]
    batches = base_batches[:a.warmup] + [base_batches[a.warmup + (i + rank) % a.steps] for i in range(a.steps)]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- resident-input timing (value)
    for i in range(a.warmup):
        eng.map_upload(batches[i]); eng.map_resident()
    eng.prof_reset()
    launches0 = eng.kernel_launches()
    barrier()
    clk = ClockSampler(dev); clk.start()
    t_res = 0.0
    t_dev_ms = 0.0
    rec = sq = n_mz = n_a = n_hits = 0
    state = np.zeros(2, dtype=np.uint32)
    for i in range(a.warmup, nb):
        eng.map_upload(batches[i])                      # untimed: inputs resident before the timed region
        torch.cuda.synchronize(dev)
        e0 = time.perf_counter()
        r = eng.map_resident()                          # all kernels of the step; internally event-timed per kernel
        torch.cuda.synchronize(dev)
        t_res += time.perf_counter() - e0
        n_mz += r["n_minimizers"]; n_a += r["n_anchors"]; n_hits += r["n_hits"]; t_dev_ms += r["device_ms"]
        hits, _ = eng.map_download()
        out = bytearray()
        rr, ss = eng.emit_step1(hits, batches[i], state, out)
        rec += rr; sq += ss
    barrier()
    clocks = clk.stop()
    launches = eng.kernel_launches() - launches0
    prof = eng.prof()
    dev_ms = sum(v[0] for v in prof.values())           # sum of CUDA-event kernel durations over the K steps

    # ---- end-to-end timing (host buffers in page-locked memory, as the contract allows)
    pinned = [engine.PinnedReads(b) for b in batches]
    br = dict(upload=0.0, kernels=0.0, download=0.0, emit=0.0)
    for i in range(a.warmup, nb):                       # untimed diagnostic pass: where the e2e time goes
        b = pinned[i].reads
        t1 = time.perf_counter(); eng.map_upload(b); t2 = time.perf_counter(); eng.map_resident(); t3 = time.perf_counter()
        hits, _ = eng.map_download(); t4 = time.perf_counter()
        out = bytearray(); eng.emit_step1(hits, b, state, out); t5 = time.perf_counter()
        br["upload"] += t2 - t1; br["kernels"] += t3 - t2; br["download"] += t4 - t3; br["emit"] += t5 - t4
    for i in range(a.warmup):                           # warm-up of the end-to-end path itself
        eng.map_batch_ovl(pinned[i].reads, state)
    barrier()
    t_e2e = 0.0
    rec_e = 0
    h2d = d2h = 0
    e2e_steps = []
    for i in range(a.warmup, nb):
        b = pinned[i].reads
        t1 = time.perf_counter()
        data, st_e = eng.map_batch_ovl(b, state)          # H2D, all kernels incl. the output stage, D2H of the .ovl bytes
        t_e2e += time.perf_counter() - t1
        e2e_steps.append(round(1e3 * (time.perf_counter() - t1), 3))
        rec_e += st_e["records"]
        h2d += b.words.nbytes + b.word_off.nbytes + b.length.nbytes + b.name_key.nbytes + b.name_hash.nbytes + b.ids.nbytes
        d2h += len(data)
    barrier()

    # ---- reduce over ranks: total units / max time
    t_res_dev = t_dev_ms / 1e3                      # one CUDA-event pair around each step on the launching stream (gaps included)
    if dist is not None:
        tt = torch.tensor([t_res_dev, t_e2e, t_res], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        cc = torch.tensor([rec, sq, rec_e, launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        t_res_dev, t_e2e, t_res = [float(x) for x in tt.tolist()]
        rec, sq, rec_e, launches = [int(x) for x in cc.tolist()]

    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    if rank == 0:
        # dominant kernel + roofline (algorithmic bytes: DESIGN.md "Kernels and rooflines")
        top = max(prof.items(), key=lambda kv: kv[1][0])
        bytes_per = {"chain_fill_warp": 32 * n_a, "chain_fill_seq": 32 * n_a, "sort_level": 32 * n_a, "sort_seq": 32 * n_a,
                     "sort_perm": 32 * n_a, "chain_backtrack": 16 * n_a + 56 * n_hits, "sketch": READ_LEN / 4 * B * a.steps + 16 * n_mz,
                     "probe_count": 32 * n_mz + 8 * n_a, "fill_anchors": 16 * n_mz + 8 * n_a + 16 * n_a}
        alg = float(bytes_per.get(top[0], 32 * n_a))
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch at the default workload, from the ncu --set full capture
        # profiles/r01d_chain_fill_warp_ncu_full_raw.csv (65.36 MB + 21.61 MB); only quoted for that kernel and that workload
        ncu_traffic = {"chain_fill_warp": 86.97e6} if (a.reads == 100000 and B == 3000) else {}
        achieved = alg / (top[1][0] / 1e3) / 1e9 if top[1][0] > 0 else 0.0
        cpu = None
        if not a.no_cpu_baseline:
            try:
                sample, desc = cpu_sample(a.cpu_sample_reads)
                threads = usable_cpus()
                r_, s_, tm, tot = run_reference(sample, threads)
                cpu = {"value": r_ / tm, "unit": "overlaps/s", "cores": threads, "kind": "reference", "sample": desc,
                       "aligned_gbp_per_s": s_ / tm / 1e9, "mapping_s": tm, "process_s": tot}
            except Exception as ex:  # noqa: BLE001
                cpu = {"value": None, "unit": "overlaps/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % ex}
        ksw = None
        if not a.no_ksw and world == 1:
            try:
                ksw = ksw_bench(eng)
                ksw["roofline_frac_hbm"] = ksw["achieved_GBps"] / hbm_peak
            except Exception as ex:  # noqa: BLE001
                ksw = {"error": repr(ex)}
        rb = None
        if not a.no_regime_b and world == 1:
            try:
                rb = regime_b_bench(eng, engine, with_reference=not a.no_cpu_baseline, full=a.regime_b_full)
            except Exception as ex:  # noqa: BLE001
                rb = {"error": repr(ex)}
        s2 = None
        if not a.no_regime_b and world == 1:
            try:
                s2 = step2_bench(eng, engine, with_reference=not a.no_cpu_baseline)
            except Exception as ex:  # noqa: BLE001
                s2 = {"error": repr(ex)}
        line = {
            "metric": "overlaps/sec", "value": rec / t_res_dev, "unit": "overlaps/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * t_res_dev / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "aligned_gbp_per_s": sq / t_res_dev / 1e9,
            "config": {"workload": "configs[1]: %d synthetic %d bp ONT reads (sub/ins/del %.0f/%.0f/%.0f %%, %dx depth), -x ava-ont -k 17 -w 10, self-overlap step 1"
                                   % (a.reads, READ_LEN, 100 * ERR["sub"], 100 * ERR["ins"], 100 * ERR["dele"], DEPTH),
                       "batch_reads_per_step": B, "parallelism": "query-sharded x%d, index replicated (all-gather of sketches)" % world,
                       "l2": "per-step working set (index %.1f GB + anchors) exceeds L2; each step maps a different batch" % (stat["n_minimizers"] * 8 / 1e9),
                       "kernel_variant": a.variant},
            "e2e": {"value": rec_e / t_e2e, "unit": "overlaps/s", "h2d_bytes_per_step": h2d // a.steps, "d2h_bytes_per_step": d2h // a.steps,
                    "ms_per_step": 1e3 * t_e2e / a.steps, "step_ms": e2e_steps, "breakdown_ms_per_step": {k: round(1e3 * v / a.steps, 3) for k, v in br.items()}},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": ncu_traffic.get(top[0]), "algorithmic_bytes_per_launch": alg / a.steps, "peak_kind": peak_kind, "kernel_ms_per_step": top[1][0] / a.steps,
                         "note": "integer DP / hashing kernels: issue- and latency-bound, not HBM-bound (ncu: profiles/r01_summary.md)"},
            "cpu_baseline": cpu,
            "kernel_sum_ms_per_step": dev_ms / a.steps,
            "work": {"records": rec, "sum_qspan": sq, "minimizers": n_mz, "anchors": n_a, "hits": n_hits, "host_wall_resident_s": t_res},
            "kernels_ms_per_step": {k: round(v[0] / a.steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
            "index": {"build_s": t_index, "n_minimizers": stat["n_minimizers"], "n_distinct": stat["n_distinct"], "mid_occ": stat["mid_occ"]},
            "gen_s": t_gen,
            "ksw2_extd2": ksw,
            "regime_b": rb,
            "step2": s2,
        }
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("NCCL_DEBUG", "WARN")            # keep stdout to the one JSON line
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // max(world, 1)))   # read generator
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        main_reference(a, rank, world)
    else:
        main_ours(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
