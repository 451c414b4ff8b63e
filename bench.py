#!/usr/bin/env python
"""bench.py -- overlaps/s (+ aligned Gbp/s) of the read-overlap hot path on synthetic 20 kb ONT reads.

Workload = BASELINE.json configs[1]: 100k synthetic 20 kb ONT reads (15 % error: 5 % sub, 4 % ins, 6 % del, 40x depth),
`--step 1 -x ava-ont -k 17 -w 10`, self-overlap, .ovl records.  The index over all reads is built once (reported separately);
a *step* maps one query batch of `--batch-reads` reads (3000 reads = 60 Mbp, the reference's `-K 60M` mini-batch) against it.
The batches are strided samples of the set (batch i = reads i, i+S, i+2S, ... with S = n // B), so that every step carries the
set's average work: the number of records a read yields depends on where its NAME sorts (MM_F_NO_DUAL keeps a pair only on the
read whose decimal name is smaller in strcmp order), and contiguous name ranges are not comparable.  Both arms time the SAME batches.

  value : whole-job overlaps/s with the step's query batch already resident in HBM (all kernels of the step, one CUDA-event pair
          per step on the launching stream, max over ranks)
  e2e   : the same through ndg_map_batch_ovl() with HOST (pinned) buffers: H2D of the packed reads, all kernels incl. the output
          stage, D2H of the .ovl bytes inside the timed region
  roofline : dominant kernel of the step (largest accumulated device time): algorithmic bytes per DESIGN.md / SURVEY.md 8(d)
          divided by its event-timed duration, against MEASURED_PEAKS.json; traffic = ncu dram bytes from profiles/ (same workload)
  cpu_baseline : the unmodified reference (oracle/_ref/minimap2-nd) on the box's host cores mapping the first batches of the same
          set against the same 100k-read index (mapping phase, from its own stderr timestamps)

`--impl reference` times that reference binary as the measured arm: one process, index over the same reads, the same W+K batches
as queries with `-K 60M` so that its mini-batches are the steps; per-step times are the differences of its
`[M::worker_pipeline] mapped` timestamps (its 3-stage pipeline included).
N > 1 (torchrun, one process per GPU): WEAK scaling -- the read set and the genome grow with N (N x 100k reads, depth fixed), each
rank sketches its own 100k reads, the index is built sharded (all-to-all by minimizer range, 1/N of the sort per rank, slices
all-gathered: engine.build_index_sharded) and replicated, then every rank maps batches of its OWN reads against the full index
with no data-path collective.  Read names are interleaved over ranks (rank r owns ids r, r+N, ...) so that the name-order rule of
MM_F_NO_DUAL gives every rank statistically identical work.
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
    ap.add_argument("--cpu-batches", type=int, default=4, help="batches the cpu_baseline leg maps (same index, same batches as the GPU arm)")
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
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
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


def make_reads(n, genome_bp, first=0, rank=0, world=1):
    """reads first..first+n of the seeded set; with world > 1 the j-th read of rank r is NAMED j*world + r (names interleave over
    ranks; file order = rank blocks), world = 1 gives names first.."""
    from nextdenovo_b200 import synth
    from nextdenovo_b200.formats import name_keys, x31_hash
    g = synth.genome(genome_bp, SEED)
    r = synth.reads(g, n, READ_LEN, seed=SEED, first=first, **ERR)
    if world > 1:
        r.ids = (np.arange(n, dtype=np.uint64) * world + rank).astype(np.uint32)
        r.name_key, r.name_hash = name_keys(r.ids), x31_hash(r.ids)
    return r


def step_batches(reads, B, n):
    """the first n strided batches of B reads (see the module docstring)"""
    from nextdenovo_b200.formats import take_reads
    S = reads.n // B
    if n > S:
        raise SystemExit("warmup+steps = %d batches of %d reads exceed the %d-read set (%d batches)" % (n, B, reads.n, S))
    return [take_reads(reads, np.arange(B, dtype=np.int64) * S + i) for i in range(n)]


def concat_reads(parts):
    from nextdenovo_b200.formats import Reads
    off = [np.zeros(1, dtype=np.uint64)]
    base = 0
    for p in parts:
        off.append(p.word_off[1:].astype(np.uint64) + np.uint64(base)); base += int(p.word_off[-1])
    return Reads(np.concatenate([p.words for p in parts]), np.concatenate(off), np.concatenate([p.length for p in parts]), np.concatenate([p.ids for p in parts]),
                 np.concatenate([p.name_key for p in parts]), np.concatenate([p.name_hash for p in parts]))


def run_reference(reads, queries, n_batches, B, threads, keep=None):
    """The unmodified reference binary: index over `reads`, queries = `queries` (the GPU arm's batches, concatenated) with `-K` such
    that its mini-batches are exactly the GPU arm's steps.  Returns per-batch (records, sum_qspan) arrays, the timestamps of its
    `mapped` lines (one per batch, printed by the ordered output stage), the timestamp of the index line, and the process time."""
    import nextdenovo_b200 as nd
    exe = os.path.join(ROOT, "oracle", "_ref", "minimap2-nd")
    if not os.path.exists(exe):
        raise RuntimeError("oracle/_ref/minimap2-nd missing (build() makes it when /root/reference is present)")
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    with tempfile.TemporaryDirectory(dir=base) as td:
        tp, qp, out = os.path.join(td, "t.2bit"), os.path.join(td, "q.2bit"), os.path.join(td, "o.ovl")
        nd.write_2bit(tp, reads)
        nd.write_2bit(qp, queries)
        t0 = time.time()
        r = subprocess.run([exe, "--step", "1", "-t", str(threads), "-I", "100G", "-K", str(B * READ_LEN), "-x", "ava-ont", "-k", "17", "-w", "10", tp, qp, "-o", out],
                           stderr=subprocess.PIPE, text=True, check=True)
        total = time.time() - t0
        t_idx, t_map = None, []
        for line in r.stderr.splitlines():
            m = re.match(r"\[M::mm_idx_stat::([0-9.]+)\*", line)
            if m and t_idx is None:
                t_idx = float(m.group(1))
            m = re.match(r"\[M::worker_pipeline::([0-9.]+)\*[0-9.]+\] mapped (\d+) sequences", line)
            if m:
                t_map.append(float(m.group(1)))
        raw = open(out, "rb").read()
        if keep is not None:
            keep.append(raw)
        d = nd.decode_ovl8(raw)
    bidx = np.full(int(reads.ids.max()) + 1, n_batches, dtype=np.int64)     # query name -> batch
    bidx[queries.ids.astype(np.int64)] = np.arange(queries.n, dtype=np.int64) // B
    bi = bidx[d[:, 0].astype(np.int64)]
    rec = np.bincount(bi, minlength=n_batches)[:n_batches]
    sq = np.bincount(bi, weights=(d[:, 3].astype(np.int64) - d[:, 2].astype(np.int64)), minlength=n_batches)[:n_batches]
    return rec, sq, t_map, t_idx, total


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
    if os.environ.get("ND_NCU_RANGE"):
        import torch
        torch.cuda.cudart().cudaProfilerStart()
    t0 = time.perf_counter()
    res, cigs = eng.ksw_extd2_batch(calls, seqs, mat, 4, 2, 24, 1)
    wall = time.perf_counter() - t0
    if os.environ.get("ND_NCU_RANGE"):
        torch.cuda.cudart().cudaProfilerStop()
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


def workload_desc(a, world=1):
    return "configs[1]: %d synthetic %d bp ONT reads (sub/ins/del %.0f/%.0f/%.0f %%, %dx depth), -x ava-ont -k 17 -w 10, self-overlap step 1" \
        % (a.reads * world, READ_LEN, 100 * ERR["sub"], 100 * ERR["ins"], 100 * ERR["dele"], DEPTH)


def main_reference(a, rank, world):
    """Reference arm: same reads, same index, same batches as the GPU arm at N=1 (rank 0 only; other ranks exit)."""
    if rank != 0:
        return
    threads = usable_cpus()
    B = a.batch_reads
    nb = a.warmup + a.steps
    reads = make_reads(a.reads, int(a.reads * READ_LEN / DEPTH))
    rec, sq, t_map, t_idx, total = run_reference(reads, concat_reads(step_batches(reads, B, nb)), nb, B, threads)
    if len(t_map) != nb:
        raise SystemExit("reference printed %d batch lines, expected %d" % (len(t_map), nb))
    t_prev = t_map[a.warmup - 1] if a.warmup > 0 else t_idx
    tsum = t_map[nb - 1] - t_prev
    nrec, nsq = int(rec[a.warmup:nb].sum()), int(sq[a.warmup:nb].sum())
    v = nrec / tsum
    sample = "%d batches x %d reads of the same set against the same %d-read index, -t %d, mapping phase (differences of the binary's own " \
             "`mapped` timestamps, 3-stage pipeline included); process %.1f s incl. index %.1f s" % (a.steps, B, a.reads, threads, total, t_idx)
    line = {"metric": "overlaps/sec", "value": v, "unit": "overlaps/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * tsum / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "impl": "reference",
            "aligned_gbp_per_s": nsq / tsum / 1e9, "reads_per_s": a.steps * B / tsum, "query_gbp_per_s": a.steps * B * READ_LEN / tsum / 1e9,
            "config": dict({"workload": workload_desc(a), "batch_reads_per_step": B, "options": "--step 1 -x ava-ont -k 17 -w 10 -K %d" % (B * READ_LEN)},
                           **({"note": "host baseline on the N=1 workload (%d-read index): the box's host cores do not grow with --gpus, and the reference needs minutes "
                                       "to index the %d reads the GPU arm maps against at N=%d" % (a.reads, a.reads * world, world)} if world > 1 else {})),
            "work": {"records": nrec, "sum_qspan": nsq},
            "cpu_baseline": {"value": v, "unit": "overlaps/s", "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "overlaps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def ncu_traffic(a, B):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the newest profiles/r*_traffic.json whose workload matches
    (written from an `ncu --set full` capture of this same command; see profiles/)."""
    import glob
    best = {}
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json"))):
        try:
            t = json.load(open(p))
            if t.get("reads") == a.reads and t.get("batch_reads") == B:
                best = {k: (float(v), os.path.basename(p)) for k, v in t.get("kernels", {}).items()}
        except Exception:  # noqa: BLE001
            pass
    return best


def main_ours(a, rank, world, local_rank):
    import hashlib
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
    B = a.batch_reads
    nb = a.warmup + a.steps

    # ---- workload: N x 100k reads over a genome that grows with N (depth fixed); rank r owns reads [r*100k, (r+1)*100k) of the
    # seeded set, named r, r+N, r+2N, ... (see the module docstring)
    genome_bp = int(a.reads * world * READ_LEN / DEPTH)
    t0 = time.time()
    reads = make_reads(a.reads, genome_bp, first=rank * a.reads, rank=rank, world=world)
    t_gen = time.time() - t0
    io, mo = nd.IdxOpt(), nd.MapOpt()
    nd.set_opt("ava-ont", io, mo)
    io.k, io.w = 17, 10
    mo.step = 1
    eng = engine.Engine(dev)
    eng.set_options(io, mo)
    sv, cv = [int(x) for x in a.variant.split(",")]
    eng.set_variant(sv, cv)

    # ---- index
    torch.cuda.synchronize(dev)
    idx_t = {}
    if dist is not None:                                   # communicator set-up (NCCL's lazy init) is not part of the index build
        w_ = torch.zeros(world * 8, dtype=torch.int64, device="cuda")
        dist.all_reduce(w_); dist.all_to_all_single(torch.empty_like(w_), w_); dist.all_gather_into_tensor(torch.empty(world * world * 8, dtype=torch.int64, device="cuda"), w_)
        dist.broadcast(w_, src=0)
        dist.barrier()
    t0 = time.time()
    if world == 1:
        stat = eng.build_index(reads)
    else:
        meta = [((np.arange(a.reads, dtype=np.uint64) * world + r).astype(np.uint32), np.full(a.reads, READ_LEN, dtype=np.uint32)) for r in range(world)]
        stat = engine.build_index_sharded(eng, reads, rank, world, dist, torch, meta=meta, timings=idx_t)
    torch.cuda.synchronize(dev)
    t_index = time.time() - t0

    eng.set_limits(a.sub_batch_mbp << 20, 0)
    batches = step_batches(reads, B, nb)

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
    ncu_range = bool(os.environ.get("ND_NCU_RANGE"))    # `ncu --profile-from-start off`: profile exactly the timed resident steps
    if ncu_range:
        torch.cuda.cudart().cudaProfilerStart()
    clk = ClockSampler(dev); clk.start()
    t_res = 0.0
    t_dev_ms = 0.0
    rec = sq = n_mz = n_a = n_hits = 0
    state = np.zeros(2, dtype=np.uint32)
    h_res = hashlib.md5()
    for i in range(a.warmup, nb):
        eng.map_upload(batches[i])                      # untimed: inputs resident before the timed region
        torch.cuda.synchronize(dev)
        e0 = time.perf_counter()
        r = eng.map_resident()                          # all kernels of the step; one CUDA-event pair around them on the launching stream
        torch.cuda.synchronize(dev)
        t_res += time.perf_counter() - e0
        n_mz += r["n_minimizers"]; n_a += r["n_anchors"]; n_hits += r["n_hits"]; t_dev_ms += r["device_ms"]
    barrier()
    if ncu_range:
        torch.cuda.cudart().cudaProfilerStop()
    launches = eng.kernel_launches() - launches0
    prof = eng.prof()
    dev_ms = sum(v[0] for v in prof.values())           # sum of CUDA-event kernel durations over the K steps

    # ---- end-to-end timing (host buffers in page-locked memory, as the contract allows)
    pinned = [engine.PinnedReads(b) for b in batches]
    for i in range(a.warmup):                           # warm-up of the end-to-end path itself (incl. the copy stream and its buffers)
        if i == 0:
            eng.map_prefetch(pinned[0].reads)
        if i + 1 < a.warmup:
            eng.map_prefetch(pinned[i + 1].reads)
        eng.map_batch_ovl(pinned[i].reads, state)
    state[:] = 0
    barrier()
    t_e2e = 0.0
    rec_e = sq_e = 0
    h2d = d2h = 0
    e2e_steps = []
    h_e2e = hashlib.md5()
    for i in range(a.warmup, nb):
        b = pinned[i].reads
        t1 = time.perf_counter()
        if i == a.warmup:
            eng.map_prefetch(b)                           # the first batch's copy starts inside the timed region as well
        if i + 1 < nb:
            eng.map_prefetch(pinned[i + 1].reads)         # the next batch's H2D copy overlaps this batch's kernels (ndg_map_prefetch)
        data, st_e = eng.map_batch_ovl(b, state)          # (H2D of this batch: prefetched,) all kernels incl. the output stage, D2H of the .ovl bytes
        dt = time.perf_counter() - t1
        t_e2e += dt
        e2e_steps.append(round(1e3 * dt, 3))
        rec_e += st_e["records"]; sq_e += st_e["sum_qspan"]
        h2d += b.words.nbytes + b.word_off.nbytes + b.length.nbytes + b.name_key.nbytes + b.name_hash.nbytes + b.ids.nbytes
        d2h += len(data)
        h_e2e.update(data)
    barrier()
    rec, sq = rec_e, sq_e                               # the records of the timed batches (the resident pass maps the same batches)

    # ---- the same K batches through a pool of two contexts sharing the index (what bin/minimap2-nd does): upload / kernels / download
    # of neighbouring batches overlap on two streams; every batch is encoded from state (0,0) and the stream is stitched afterwards
    eng2 = engine.Engine(dev)
    eng2.set_options(io, mo)
    eng2.set_variant(sv, cv)
    eng2.set_limits(a.sub_batch_mbp << 20, 0)
    eng2.share_index_of(eng)
    pool = [eng, eng2]
    engine.map_batches_pool(pool, [pinned[i].reads for i in range(a.warmup)] * 2)        # warm-up (both contexts allocate their workspace)
    barrier()
    t1 = time.perf_counter()
    chunks, states, _ = engine.map_batches_pool(pool, [pinned[i].reads for i in range(a.warmup, nb)])
    st2 = np.zeros(2, dtype=np.uint32)
    stitched = engine.stitch_ovl(chunks, states, st2)
    t_pool = time.perf_counter() - t1
    barrier()
    clocks = clk.stop()                                 # sampled across the three timed regions (resident, one context, pool)
    pool_md5 = hashlib.md5(stitched).hexdigest()
    eng2.close()

    # ---- the whole read set of this rank, end to end, in file order (untimed region; reported beside the headline)
    state[:] = 0
    h_all = hashlib.md5()
    all_rec = all_sq = 0
    first_bytes = []
    t1 = time.perf_counter()
    for i in range(0, reads.n, B):
        data, st_e = eng.map_batch_ovl(reads.slice(i, min(i + B, reads.n)), state)
        all_rec += st_e["records"]; all_sq += st_e["sum_qspan"]
        h_all.update(data)
    t_all = time.perf_counter() - t1
    if rank == 0 and not a.no_cpu_baseline and world == 1:               # the bytes the cpu_baseline leg is checked against
        state[:] = 0
        for i in range(a.cpu_batches):
            first_bytes.append(eng.map_batch_ovl(batches[i], state)[0])

    # ---- reduce over ranks: total units / max time
    t_res_dev = t_dev_ms / 1e3                      # one CUDA-event pair around each step on the launching stream (gaps included)
    if dist is not None:
        tt = torch.tensor([t_res_dev, t_e2e, t_res, t_index, t_all, t_pool], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        cc = torch.tensor([rec, sq, launches, all_rec, n_mz, n_a, n_hits], dtype=torch.float64, device="cuda")
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        t_res_dev, t_e2e, t_res, t_index, t_all, t_pool = [float(x) for x in tt.tolist()]
        rec, sq, launches, all_rec, n_mz, n_a, n_hits = [int(x) for x in cc.tolist()]

    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    if rank == 0:
        # dominant kernel + roofline.  Algorithmic bytes per kernel family, for the K timed steps of all ranks:
        L_bytes = READ_LEN / 4 * B * a.steps * world
        n_ovl = rec
        bytes_per = {"sketch_fast": L_bytes + 16 * n_mz, "sketch": L_bytes + 16 * n_mz, "sketch_compact": 32 * n_mz,
                     "probe_count": 32 * n_mz + 8 * n_a, "fill_anchors": 16 * n_mz + 8 * n_a + 16 * n_a,
                     "sort_level": 32 * n_a, "sort_seq": 32 * n_a, "sort_init": 16 * n_a,
                     "chain_fill_warp": 32 * n_a, "chain_fill_seq": 32 * n_a, "chain_backtrack": 16 * n_a + 56 * n_hits,
                     "hit_order": 2 * 56 * n_hits, "emit": 56 * n_hits + 13 * n_ovl}
        step_bytes = L_bytes + 32 * n_mz + 88 * n_a + 13 * n_ovl          # B_A of SURVEY 8(d), whole step
        top = max(((k, v) for k, v in prof.items() if k in bytes_per), key=lambda kv: kv[1][0])
        alg = float(bytes_per[top[0]]) / world          # per rank: prof[] holds this rank's kernel times
        tr = ncu_traffic(a, B)
        achieved = alg / (top[1][0] / 1e3) / 1e9 if top[1][0] > 0 else 0.0
        per_kernel = {k: {"ms_per_step": round(v[0] / a.steps, 4), "launches_per_step": v[1] / a.steps,
                          "alg_GBps": round(bytes_per[k] / world / (v[0] / 1e3) / 1e9, 1) if k in bytes_per and v[0] > 0 else None,
                          "frac": round(bytes_per[k] / world / (v[0] / 1e3) / 1e9 / hbm_peak, 4) if k in bytes_per and v[0] > 0 else None,
                          "ncu_dram_bytes_per_launch": tr.get(k, (None,))[0]}
                      for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]) if v[0] > 0}
        cpu = None
        if not a.no_cpu_baseline and world == 1:
            try:
                threads = usable_cpus()
                keep = []
                r_rec, r_sq, t_map, t_idx, tot = run_reference(reads, concat_reads(batches[:a.cpu_batches]), a.cpu_batches, B, threads, keep=keep)
                tm = t_map[-1] - t_idx
                mine_bytes = b"".join(first_bytes)
                cpu = {"value": float(r_rec.sum()) / tm, "unit": "overlaps/s", "cores": threads, "kind": "reference",
                       "sample": "first %d batches x %d reads of the same set against the same %d-read index, mapping phase from the binary's own timestamps"
                                 % (a.cpu_batches, B, a.reads),
                       "aligned_gbp_per_s": float(r_sq.sum()) / tm / 1e9, "mapping_s": tm, "process_s": tot, "index_s": t_idx,
                       "ovl_bytes_identical_to_gpu": keep[0] == mine_bytes, "ovl_bytes": len(keep[0])}
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
            "aligned_gbp_per_s": sq / t_res_dev / 1e9, "reads_per_s": a.steps * B * world / t_res_dev, "query_gbp_per_s": a.steps * B * world * READ_LEN / t_res_dev / 1e9,
            "config": {"workload": workload_desc(a, world), "batch_reads_per_step": B,
                       "parallelism": "query-sharded x%d (each rank maps its own reads), index replicated, built sharded (all-to-all by minimizer range + all-gather of slices)" % world,
                       "l2": "per-step working set (index table %.1f GB + anchors) exceeds L2; every step maps a different batch" % (stat["n_distinct"] / 0.7 * 16 / 1e9),
                       "kernel_variant": a.variant},
            "e2e": {"value": rec / t_e2e, "unit": "overlaps/s", "h2d_bytes_per_step": h2d // a.steps, "d2h_bytes_per_step": d2h // a.steps,
                    "ms_per_step": 1e3 * t_e2e / a.steps, "step_ms": e2e_steps, "ovl_md5_rank0": h_e2e.hexdigest(),
                    "how": "the K timed batches from pinned host buffers through ndg_map_prefetch + ndg_map_batch_ovl (every batch's H2D copy, all kernels incl. the output stage and the D2H of the .ovl bytes lie inside the timed region; the copy of batch i+1 runs on a copy stream under the kernels of batch i), one call per step",
                    "two_context_pool": {"value": rec / t_pool, "ms_per_step": 1e3 * t_pool / a.steps, "same_bytes": pool_md5 == h_e2e.hexdigest(),
                                         "how": "the same batches on two contexts sharing the index (two host threads, two streams), stitched with ndg_ovl_rebase_first: what bin/minimap2-nd runs"}},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": tr.get(top[0], (None,))[0], "traffic_source": tr.get(top[0], (None, None))[1],
                         "algorithmic_bytes_per_launch": alg / top[1][1] if top[1][1] else None, "launches_per_step": top[1][1] / a.steps,
                         "peak_kind": peak_kind, "kernel_ms_per_step": top[1][0] / a.steps,
                         "whole_step": {"algorithmic_bytes_per_step": step_bytes / world / a.steps, "achieved": step_bytes / world / t_res_dev / 1e9,
                                        "frac": step_bytes / world / t_res_dev / 1e9 / hbm_peak},
                         "note": "integer DP / hashing kernels: issue- and latency-bound, not HBM-bound (ncu: profiles/)"},
            "cpu_baseline": cpu,
            "kernel_sum_ms_per_step": dev_ms / a.steps,
            "work": {"records": rec, "sum_qspan": sq, "minimizers": n_mz, "anchors": n_a, "hits": n_hits, "host_wall_resident_s": t_res},
            "kernels": per_kernel,
            "whole_set": {"reads": a.reads * world, "records": all_rec, "seconds_e2e_max_rank": t_all, "overlaps_per_s": all_rec / t_all,
                          "reads_per_s": a.reads * world / t_all, "ovl_md5_rank0": h_all.hexdigest()},
            "index": {"build_s": t_index, "stages_s_rank0": idx_t, "n_minimizers": stat["n_minimizers"], "n_distinct": stat["n_distinct"], "mid_occ": stat["mid_occ"],
                      "table_GB": stat["n_distinct"] / 0.7 * 16 / 1e9, "positions_GB": stat["n_minimizers"] * 8 / 1e9},
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
