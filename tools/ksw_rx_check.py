"""Development aid (GPU box): register-resident ksw kernel vs the reference function on many calls, with a breakdown of what differs.
usage: python tools/ksw_rx_check.py [n_random] [n_wide] > gpurun_out/ksw_rx_check.log"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ksw_cases import simple_mat, oracle_lib, oracle_call, ref_lib, ref_call, random_cases, wide_cases
from nextdenovo_b200 import engine
from nextdenovo_b200.engine import KSW_CALL_DTYPE

nr = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 64
mat = simple_mat()
cases = random_cases(nr, seed=23) + wide_cases(nw) + random_cases(nr // 2, seed=99, max_len=500)
calls = np.zeros(len(cases), dtype=KSW_CALL_DTYPE)
chunks, off = [], 0
for i, (qs, ts, w, zd, eb, flag) in enumerate(cases):
    calls[i] = (len(qs), len(ts), off, off + len(qs), w, zd, eb, flag)
    chunks += [qs, ts]; off += len(qs) + len(ts)
seqs = np.concatenate(chunks)
eng = engine.Engine(0)
eng.set_ksw_variant(3)
t0 = time.time()
res, cigs = eng.ksw_extd2_batch(calls, seqs, mat, 4, 2, 24, 1)
print("rx batch of %d calls: %.2f s" % (len(cases), time.time() - t0))
L, R = oracle_lib(), ref_lib()
names = ("max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar", "reach_end")
RXCAP = [512, 1024, 2048, 3072, 4096]
def cls(q, t, w):
    T = (t + 15) // 16 * 16
    if w < 0: w = max(q, t)
    for i, c in enumerate(RXCAP):
        if T <= c or w + 66 <= c: return i
    return -1
bad, by = [], {}
for i, (qs, ts, w, zd, eb, flag) in enumerate(cases):
    want = ref_call(R, qs, ts, mat, 4, 2, 24, 1, w, zd, eb, flag) if R is not None else oracle_call(L, qs, ts, mat, 4, 2, 24, 1, w, zd, eb, flag)
    got = (tuple(int(res[n][i]) for n in names), [int(x) for x in cigs[i]])
    key = (cls(len(qs), len(ts), w), "exact" if not flag & 8 else "approx", hex(flag), "ring" if (len(ts) + 15) // 16 * 16 > (RXCAP[cls(len(qs), len(ts), w)] if cls(len(qs), len(ts), w) >= 0 else 1 << 30) else "flat")
    st = by.setdefault(key, [0, 0]); st[0] += 1
    if got != want:
        st[1] += 1
        fd = [n for n, a, b in zip(names, got[0], want[0]) if a != b]
        cd = next((k for k, (a, b) in enumerate(zip(got[1], want[1])) if a != b), None) if got[1] != want[1] else None
        bad.append((i, len(qs), len(ts), w, zd, eb, hex(flag), fd, cd, got[0], want[0], got[1][:6], want[1][:6], int((qs == 4).any() or (ts == 4).any())))
print("mismatching calls: %d / %d" % (len(bad), len(cases)))
for k in sorted(by): print("  class %s: %d calls, %d bad" % (k, by[k][0], by[k][1]))
for b in bad[:25]: print(b)
# speed on the synthetic gap-fill batch of bench.py
sys.path.insert(0, ROOT)
import bench
for v in (1, 3):
    eng.set_ksw_variant(v)
    print("variant", v, {k: (round(x, 2) if isinstance(x, float) else x) for k, x in bench.ksw_bench(eng).items() if k in ("gcups", "kernel_ms", "e2e_gcups")})
eng.close()
