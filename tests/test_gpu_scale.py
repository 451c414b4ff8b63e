"""Parity at the sizes BASELINE.json names (run on the B200 box: pytest -m gpu).  Every case runs the GPU drop-in executable
(bin/minimap2-nd) and the unmodified reference (oracle/_ref/minimap2-nd, on the box's host cores) on the same seeded synthetic
reads and compares the WHOLE output files byte for byte; plus size-independent properties of the full output.
  configs[1]  100k x 20 kb ONT (15 % error), -k 17 -w 10: the full size
  configs[2]  ava-ont preset (k15 w5) through the `-I` index-part loop (ref: minimap2/main.c:477-509): reduced to 12k reads, 3 parts
  configs[3]  PacBio CLR 15 kb (12 % error), -x ava-pb -r 500 -c (ksw band 751, ref: minimap2/align.c:580): reduced to 1500 reads
  regime B / step 2 on synthetic sets of >= 5k reads (round-1 gap: both were only ever run on the 997 real reads)
Needs oracle/_ref (travels with the repo).  The reference dominates the run time (a few minutes in total)."""
import os
import subprocess
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "bin", "minimap2-nd")
REF = os.path.join(ROOT, "oracle", "_ref", "minimap2-nd")
SEED = 20240923


def _threads():
    n = len(os.sched_getaffinity(0))
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = min(n, max(1, -(-int(q) // int(per))))
    except Exception:  # noqa: BLE001
        pass
    return str(n)


def _synth(tmp_path, n, L, depth, sub, ins, dele, name="r.2bit"):
    import nextdenovo_b200 as nd
    from nextdenovo_b200 import synth
    g = synth.genome(max(n * L // depth, 4 * L), SEED)
    reads = synth.reads(g, n, L, sub=sub, ins=ins, dele=dele, seed=SEED)
    p = str(tmp_path / name)
    nd.write_2bit(p, reads)
    return p, reads


def _both(tmp_path, args, target, query, bl=False):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref not present")
    g, r = str(tmp_path / "gpu.ovl"), str(tmp_path / "ref.ovl")
    a = subprocess.run([EXE] + args + ["-t", _threads(), target, query, "-o", g], capture_output=True, text=True, timeout=1800)
    assert a.returncode == 0, a.stderr[-3000:]
    b = subprocess.run([REF] + args + ["-t", _threads(), target, query, "-o", r], capture_output=True, text=True, timeout=3000)
    assert b.returncode == 0, b.stderr[-3000:]
    got, want = open(g, "rb").read(), open(r, "rb").read()
    assert len(want) > 1000
    if got != want:
        import nextdenovo_b200 as nd
        try:
            x, y = nd.decode_ovl8(got), nd.decode_ovl8(want)
            n = min(len(x), len(y)); bad = np.nonzero((x[:n] != y[:n]).any(axis=1))[0]
            msg = "records %d vs %d, first difference at record %s: got %s want %s" % (len(x), len(y), bad[:1], x[bad[0]] if len(bad) else None, y[bad[0]] if len(bad) else None)
        except Exception as ex:  # noqa: BLE001
            msg = "undecodable (%s)" % ex
        raise AssertionError("GPU output differs from the reference: " + msg)
    if bl:
        assert open(g + ".bl", "rb").read() == open(r + ".bl", "rb").read()
    return got, a.stderr


def test_configs1_whole_output_equals_reference(built, tmp_path):
    """BASELINE configs[1] at full size: the whole .ovl of the 100k-read self-overlap, not a prefix"""
    import nextdenovo_b200 as nd
    n = int(os.environ.get("ND_SCALE_READS", "100000"))
    p, reads = _synth(tmp_path, n, 20000, 40, .05, .04, .06)
    got, _ = _both(tmp_path, ["--step", "1", "-I", "100G", "-x", "ava-ont", "-k", "17", "-w", "10"], p, p)
    rec = nd.decode_ovl8(got)
    q, t = rec[:, 0].astype(np.int64), rec[:, 4].astype(np.int64)
    assert rec.shape[0] > 10 * n
    assert (np.diff(q) >= 0).all()                                   # query order
    assert (rec[:, 3] - rec[:, 2] >= 500).all()                      # --minlen
    assert (rec[:, 3] <= 20000).all() and (rec[:, 6] <= 20000).all() and (rec[:, 2] < rec[:, 3]).all() and (rec[:, 5] < rec[:, 6]).all()
    assert (q != t).all()
    # NO_DUAL: a pair is reported once, from the query whose name sorts first (strcmp order of the decimal names)
    names_q = np.array([str(x) for x in q[:20000]]); names_t = np.array([str(x) for x in t[:20000]])
    assert (names_q < names_t).all()
    pairs = set(zip(q.tolist(), t.tolist(), rec[:, 1].tolist(), rec[:, 2].tolist()))
    assert len(pairs) == rec.shape[0]


def test_configs2_reduced_index_parts(built, tmp_path):
    """configs[2] (ava-ont k15 w5, index too big for one part) at reduced size: 12k reads, -I 100M => several index parts, each query
    re-mapped against every part, mid_occ from the first part (ref: minimap2/main.c:477-509, options.c:70-71)"""
    p, reads = _synth(tmp_path, 12000, 20000, 40, .05, .04, .06)
    got, err = _both(tmp_path, ["--step", "1", "-I", "100M", "-x", "ava-ont"], p, p)
    assert "index parts: 2" in err or "index parts: 3" in err, err[-500:]


def test_configs3_reduced_clr_cigar_band751(built, tmp_path):
    """configs[3] (PacBio CLR 15 kb, 12 % error ins-heavy, -x ava-pb, -r 500 => ksw band (int)(500*1.5+1) = 751) with -c, reduced to
    1500 reads: HPC sketch + regime B on a synthetic set"""
    p, reads = _synth(tmp_path, 1500, 15000, 30, .03, .06, .03)
    _both(tmp_path, ["--step", "1", "-c", "-x", "ava-pb", "-r", "500"], p, p)


def test_regime_b_5k_synthetic(built, tmp_path):
    """`--step 1 -c -x ava-ont` on 5000 synthetic 10 kb ONT reads (15 % error): bytes of the reference"""
    p, reads = _synth(tmp_path, 5000, 10000, 30, .05, .04, .06)
    _both(tmp_path, ["--step", "1", "-c", "-x", "ava-ont"], p, p)


def test_step2_5k_synthetic(built, tmp_path):
    """`--step 2` (cns_align options of the pipeline) on 5000 synthetic corrected-like reads (1 % error): .ovl and .bl bytes"""
    p, reads = _synth(tmp_path, 5000, 20000, 30, .004, .003, .003)
    _both(tmp_path, ["-I", "6G", "--step", "2", "-x", "ava-ont", "-k", "17", "-w", "17", "--minlen", "2000", "--maxhan1", "5000"], p, p, bl=True)
