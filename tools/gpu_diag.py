#!/usr/bin/env python
"""Diagnostic run for the GPU box: stage-by-stage status with as much detail as fits in gpurun_out/.
usage: python tools/gpu_diag.py [n_reads]"""
import os, sys, time, json, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nextdenovo_b200 as nd
from nextdenovo_b200 import engine

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
G = os.path.join(ROOT, "tests", "golden")


def opts(preset, k=None, w=None):
    io, mo = nd.IdxOpt(), nd.MapOpt()
    nd.set_opt(preset, io, mo)
    if k: io.k = k
    if w: io.w = w
    mo.step = 1
    return io, mo


def diff(got, want):
    try:
        a, b = nd.decode_ovl8(got), nd.decode_ovl8(want)
    except Exception as ex:
        return "undecodable: %r (len %d vs %d)" % (ex, len(got), len(want))
    n = min(len(a), len(b))
    bad = np.nonzero((a[:n] != b[:n]).any(axis=1))[0]
    s = "records %d vs %d, differing rows %d" % (len(a), len(b), len(bad))
    for i in bad[:5]:
        s += "\n   row %d got %s want %s" % (i, a[i].tolist(), b[i].tolist())
    sa = set(map(tuple, a.tolist())); sb = set(map(tuple, b.tolist()))
    s += "\n   set diff: only_got %d only_want %d" % (len(sa - sb), len(sb - sa))
    return s


def main():
    eng = engine.Engine(0)
    sub = nd.read_2bit(os.path.join(G, "sub120.2bit"))
    full = nd.read_2bit(os.path.join(G, "reads_test.seed.2bit"))
    res = {}
    for name, T, Q, preset, k, w, gold in [
        ("S1", sub, sub, "ava-ont", None, None, "S1_sub120_ava-ont.ovl"),
        ("S3", sub, sub, "ava-pb", None, None, "S3_sub120_ava-pb.ovl"),
        ("G2", full, full, "ava-ont", 17, 10, "G2_ava-ont_k17w10.ovl"),
        ("G1", full, full, "ava-ont", None, None, "G1_ava-ont.ovl"),
        ("G3", full, full, "ava-pb", None, None, "G3_ava-pb.ovl"),
    ]:
        for variant in ((0, 0), (1, 0), (0, 1), (1, 1)):
            if name in ("G1", "G3") and variant != (1, 1):
                continue
            tag = "%s v%d%d" % (name, *variant)
            try:
                io, mo = opts(preset, k, w)
                eng.set_variant(*variant)
                eng.prof_reset()
                t0 = time.time()
                ovl, st = engine.run_step1(eng, T, Q, io, mo)
                dt = time.time() - t0
                want = open(os.path.join(G, gold), "rb").read()
                ok = ovl == want
                prof = {k2: round(v[0], 3) for k2, v in eng.prof().items() if v[0] > 0}
                print("[%s] %s wall=%.3fs stats=%s" % (tag, "OK" if ok else "MISMATCH", dt, st), flush=True)
                print("      prof(ms)=%s" % prof, flush=True)
                if not ok:
                    print("      " + diff(ovl, want), flush=True)
                    open(os.path.join(OUT, "bad_%s_v%d%d.ovl" % (name, *variant)), "wb").write(ovl)
                res[tag] = dict(ok=ok, wall=dt, stats={k2: int(v) for k2, v in st.items()}, prof=prof)
            except Exception as ex:
                print("[%s] EXCEPTION %r" % (tag, ex), flush=True)
                traceback.print_exc()
                res[tag] = dict(ok=False, error=repr(ex))
                eng.close()
                eng = engine.Engine(0)
    json.dump(res, open(os.path.join(OUT, "diag.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
