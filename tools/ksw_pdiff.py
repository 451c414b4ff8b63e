import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ksw_cases import simple_mat, random_cases
from nextdenovo_b200 import engine
from nextdenovo_b200.engine import KSW_CALL_DTYPE
eng = engine.Engine(0)
mat = simple_mat()
cases = random_cases(40, seed=23)
for ci, (qs, ts, w, zd, eb, flag) in enumerate(cases):
    calls = np.zeros(1, dtype=KSW_CALL_DTYPE); calls[0] = (len(qs), len(ts), 0, len(qs), w, zd, eb, flag)
    seqs = np.concatenate([qs, ts])
    out = []
    for v in (0, 1):
        eng.set_ksw_variant(v)
        res, cig = eng.ksw_extd2_batch(calls, seqs, mat, 4, 2, 24, 1)
        out.append((res.copy(), [int(x) for x in cig[0]], eng.debug_fetch("ksw_p", np.uint8).copy()))
    same = out[0][1] == out[1][1]
    qlen, tlen = len(qs), len(ts)
    ww = max(qlen, tlen) if w < 0 else w
    ncol = ((min(min(qlen, tlen), ww + 1) + 15) // 16 + 1) * 16
    if not same and not (flag & 1):
        p0, p1 = out[0][2], out[1][2]
        n = (qlen + tlen - 1) * ncol
        d = np.nonzero(p0[:n] != p1[:n])[0]
        # only in-band cells matter
        print("case", ci, "q", qlen, "t", tlen, "w", w, "flag", hex(flag), "cigar differ; p diffs:", len(d))
        for k in d[:12]:
            r, off = divmod(int(k), ncol)
            st = max(0, r - qlen + 1, (r - ww + 1) >> 1); en = min(tlen - 1, r, (r + ww) >> 1)
            print("   r", r, "col", off, "t", (st // 16 * 16) + off, "st0", st, "en0", en, "v0", hex(p0[k]), "v1", hex(p1[k]))
        break
else:
    print("no differing case found")
