"""Development aid (GPU box): throughput of the DP kernels on extension-like calls (L x L, band 3001, exact maximum, z-drop 400,
KSW_EZ_EXTZ_ONLY | KSW_EZ_RIGHT | KSW_EZ_REV_CIGAR) and on long unbanded joins.  usage: python tools/ksw_wide_bench.py [n_calls] [L]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from ksw_cases import simple_mat
from nextdenovo_b200 import engine
from nextdenovo_b200.engine import KSW_CALL_DTYPE
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
L = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
rng = np.random.default_rng(7)
mat = simple_mat()
def mk(L, w, flag):
    calls = np.zeros(n, dtype=KSW_CALL_DTYPE); chunks = []; off = 0; cells = 0; cache = {}
    for i in range(n):
        t = rng.integers(0, 4, L).astype(np.uint8)
        u = rng.random(L)
        qv = np.where((u >= 0.06) & (u < 0.11), (t + 1 + rng.integers(0, 3, L)) % 4, t).astype(np.uint8)[u >= 0.06]
        ins = np.nonzero(rng.random(qv.shape[0]) < 0.04)[0]
        qv = np.insert(qv, ins, rng.integers(0, 4, ins.shape[0]).astype(np.uint8))
        calls[i] = (len(qv), L, off, off + len(qv), w, 400, -1, flag)
        chunks += [qv, t]; off += len(qv) + L
        key = (len(qv), L, w)
        if key not in cache: cache[key] = bench.ksw_cells(len(qv), L, w if w >= 0 else max(len(qv), L))
        cells += cache[key]
    return calls, np.concatenate(chunks), cells
eng = engine.Engine(0)
for name, (LL, w, flag) in {"extension %dx%d w3001 exact" % (L, L): (L, 3001, 0x40 | 0x02 | 0x80), "gapfill 413 approx": (413, 3001, 0x08), "fill 1500 approx": (1500, 3001, 0x08),
                           "second pass 413 exact": (413, 3001, 0)}.items():
    calls, seqs, cells = mk(LL, w, flag)
    if LL < 1000:
        calls = np.tile(calls, 8); 
        cells *= 8
    for v in (1, 2, 3):
        if v == 2 and LL < 1000: continue
        eng.set_ksw_variant(v)
        eng.ksw_extd2_batch(calls[:64], seqs, mat, 4, 2, 24, 1)
        eng.prof_reset()
        eng.ksw_extd2_batch(calls, seqs, mat, 4, 2, 24, 1)
        ms = eng.prof()["ksw_extd2"][0]
        print("%-34s variant %d: %8.2f ms  %7.1f GCUPS" % (name, v, ms, cells / ms / 1e6), flush=True)
eng.close()
