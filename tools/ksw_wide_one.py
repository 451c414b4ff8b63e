"""ncu target: one launch of the register-resident kernel on extension-like calls.  usage: python tools/ksw_wide_one.py [n_calls] [L] [flag] [w]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from ksw_cases import simple_mat
from nextdenovo_b200 import engine
from nextdenovo_b200.engine import KSW_CALL_DTYPE
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
L = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
flag = int(sys.argv[3], 0) if len(sys.argv) > 3 else 0xc2
w = int(sys.argv[4]) if len(sys.argv) > 4 else 3001
rng = np.random.default_rng(7)
mat = simple_mat()
calls = np.zeros(n, dtype=KSW_CALL_DTYPE); chunks = []; off = 0; cells = 0
for i in range(n):
    t = rng.integers(0, 4, L).astype(np.uint8)
    u = rng.random(L)
    qv = np.where((u >= 0.06) & (u < 0.11), (t + 1 + rng.integers(0, 3, L)) % 4, t).astype(np.uint8)[u >= 0.06]
    ins = np.nonzero(rng.random(qv.shape[0]) < 0.04)[0]
    qv = np.insert(qv, ins, rng.integers(0, 4, ins.shape[0]).astype(np.uint8))
    calls[i] = (len(qv), L, off, off + len(qv), w, 400, -1, flag)
    chunks += [qv, t]; off += len(qv) + L
    cells += bench.ksw_cells(len(qv), L, w)
eng = engine.Engine(0)
eng.set_ksw_variant(3)
eng.ksw_extd2_batch(calls, np.concatenate(chunks), mat, 4, 2, 24, 1)
eng.prof_reset()
eng.ksw_extd2_batch(calls, np.concatenate(chunks), mat, 4, 2, 24, 1)
ms = eng.prof()["ksw_extd2"][0]
print("cells", cells, "ms", ms, "GCUPS", cells / ms / 1e6)
eng.close()
