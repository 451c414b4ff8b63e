#!/usr/bin/env python
"""ovl_sort on a configs[1]-style block: N synthetic 20 kb reads -> bin/minimap2-nd --step 1 -> .ovl -> bin/ovl_sort (GPU) next to the
unmodified reference oracle/_ref/ovl_sort on the host cores; bytes compared, wall times printed as one JSON line.
usage: python tools/ovlsort_bench.py [n_reads=30000]"""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nextdenovo_b200 as nd
from nextdenovo_b200 import synth
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
base = "/dev/shm" if os.path.isdir("/dev/shm") else None
with tempfile.TemporaryDirectory(dir=base) as td:
    g = synth.genome(n * 20000 // 40, bench.SEED)
    reads = synth.reads(g, n, 20000, seed=bench.SEED, **bench.ERR)
    p = os.path.join(td, "r.2bit"); nd.write_2bit(p, reads)
    with open(os.path.join(td, "seeds.idx"), "w") as f:
        for i in range(reads.n):
            f.write("%d\t0\t%d\n" % (reads.ids[i], reads.length[i]))
    t0 = time.time()
    subprocess.run([os.path.join(ROOT, "bin", "minimap2-nd"), "--step", "1", "-x", "ava-ont", "-k", "17", "-w", "10", p, p, "-o", os.path.join(td, "a.ovl")], check=True, capture_output=True)
    t_map = time.time() - t0
    open(os.path.join(td, "in.fofn"), "w").write(os.path.join(td, "a.ovl") + "\n")
    t0 = time.time()
    r = subprocess.run([os.path.join(ROOT, "bin", "ovl_sort"), "-i", "seeds.idx", "-o", "gpu.ovl", "in.fofn"], cwd=td, capture_output=True, text=True)
    t_gpu = time.time() - t0
    assert r.returncode == 0, r.stderr
    thr = bench.usable_cpus()
    t0 = time.time()
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ovl_sort"), "-i", "seeds.idx", "-t", str(max(thr, 2)), "-m", "40G", "-o", "ref.ovl", "in.fofn"], cwd=td, check=True, capture_output=True)
    t_ref = time.time() - t0
    same = open(os.path.join(td, "gpu.ovl"), "rb").read() == open(os.path.join(td, "ref.ovl"), "rb").read() and \
        open(os.path.join(td, "gpu.ovl.bl"), "rb").read() == open(os.path.join(td, "ref.ovl.bl"), "rb").read()
    print(json.dumps({"reads": n, "ovl_in_bytes": os.path.getsize(os.path.join(td, "a.ovl")), "ovl_out_bytes": os.path.getsize(os.path.join(td, "gpu.ovl")),
                      "gpu_process_s": t_gpu, "gpu_stderr": r.stderr.strip().splitlines()[-1], "reference_s": t_ref, "reference_threads": thr, "same_bytes": same,
                      "speedup": t_ref / t_gpu, "minimap2_nd_s": t_map}))
