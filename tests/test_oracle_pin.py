"""Pins the oracle (oracle/nd_oracle.c, a CPU restatement) against
  (1) golden .ovl fixtures minted from the unmodified reference (tests/golden/make_golden.py), and
  (2) the unmodified reference itself (oracle/_ref/libminimap2_ref.so) stage by stage, when built.
CPU only."""
import ctypes as C
import hashlib
import os
import numpy as np
import pytest

from oracle import oracle


def _md5(p):
    return hashlib.md5(open(p, "rb").read()).hexdigest()


@pytest.mark.parametrize("preset,gold", [("ava-ont", "S1_sub120_ava-ont.ovl"), ("ava-pb", "S3_sub120_ava-pb.ovl")])
def test_oracle_reproduces_reference_ovl_sub120(golden_dir, tmp_path, preset, gold):
    o = oracle.make_opt(preset)
    src = os.path.join(golden_dir, "sub120.2bit")
    out = str(tmp_path / "o.ovl")
    n, sq = oracle.step1_files(src, src, o, out)
    assert open(out, "rb").read() == open(os.path.join(golden_dir, gold), "rb").read()
    assert n > 1000 and sq > 0


def test_oracle_reproduces_reference_ovl_full_k17w10(golden_dir, tmp_path):
    """997 reads of test_data/reads_test.fa.gz, -x ava-ont -k17 -w10: md5 270d357d..."""
    o = oracle.make_opt("ava-ont", k=17, w=10)
    src = os.path.join(golden_dir, "reads_test.seed.2bit")
    out = str(tmp_path / "o.ovl")
    n, _ = oracle.step1_files(src, src, o, out)
    assert n == 128381
    assert _md5(out) == "270d357dfda1e890aa8330b316ffcf1f" == _md5(os.path.join(golden_dir, "G2_ava-ont_k17w10.ovl"))


def test_golden_md5s(golden_dir):
    want = {"G1_ava-ont.ovl": "9fce14e8841db2e46012c58ad8f0059c", "G3_ava-pb.ovl": "7fb430f9dbd8032c43e3613e6d35680e",
            "G4_ava-ont_c.ovl": "4c7e852512ee6fb463b37a0ce20926b7"}
    for f, m in want.items():
        assert _md5(os.path.join(golden_dir, f)) == m


# ---- stage probes against the compiled reference ---------------------------------------------------
class _mm128v(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.c_void_p)]


@pytest.fixture(scope="module")
def ref():
    p = os.path.join(oracle.REF_DIR, "libminimap2_ref.so")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    L = C.CDLL(p)
    L.mm_sketch.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.POINTER(_mm128v)]
    L.radix_sort_128x.argtypes = [C.c_void_p, C.c_void_p]
    return L


def _ref_sketch(L, seq, w, k, rid, hpc):
    v = _mm128v(0, 0, None)
    L.mm_sketch(None, seq, len(seq), w, k, rid, hpc, C.byref(v))
    out = np.ctypeslib.as_array(C.cast(v.a, C.POINTER(C.c_uint64)), shape=(v.n, 2)).copy() if v.n else np.zeros((0, 2), np.uint64)
    C.CDLL(None).free(C.c_void_p(v.a))
    return out


@pytest.mark.parametrize("w,k,hpc", [(5, 15, 0), (10, 17, 0), (5, 19, 1), (3, 16, 0), (1, 11, 0), (12, 28, 1),
                                     (51, 51, 1), (51, 51, 0), (7, 29, 1), (9, 32, 1), (5, 33, 0), (20, 64, 1), (11, 65, 0), (6, 97, 1), (4, 128, 0)])
def test_sketch_vs_reference(ref, w, k, hpc):
    """ndo_sketch == mm_sketch on random, low-complexity, N-containing and very short sequences"""
    rng = np.random.default_rng(7 + w * 31 + k)
    seqs = []
    for L in (1, 5, k, k + w - 2, k + w - 1, k + w, 100, 1500, 6000):
        seqs.append(bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), L)))
    seqs.append(b"A" * 400)                                     # homopolymer: every k-mer identical
    seqs.append(b"AC" * 300 + b"ACGT" * 100)                    # short-period repeats, symmetric k-mers for even k
    s = bytearray(rng.choice(np.frombuffer(b"ACGT", np.uint8), 3000).tobytes())
    for pos in (0, 17, 18, 500, 501, 502, 1400, 2999):          # ambiguous bases
        s[pos] = ord("N")
    seqs.append(bytes(s))
    s2 = bytearray(rng.choice(np.frombuffer(b"ACGT", np.uint8), 2000).tobytes())
    s2[300:340] = b"T" * 40; s2[900:1300] = b"G" * 400           # long homopolymers (HPC span >= 256)
    seqs.append(bytes(s2))
    for i, seq in enumerate(seqs):
        got = oracle.sketch(seq, w, k, rid=i, hpc=hpc)
        want = _ref_sketch(ref, seq, w, k, i, hpc)
        assert got.shape == want.shape and (got == want).all(), "seq %d len %d" % (i, len(seq))


def test_radix_sort_tie_order_vs_reference(ref):
    """the unstable in-place MSD radix sort: same final order of equal keys"""
    rng = np.random.default_rng(3)
    for n in (1, 2, 64, 65, 300, 5000, 40000):
        for distinct in (3, 50, 100000):
            a = np.zeros((n, 2), dtype=np.uint64)
            a[:, 0] = (rng.integers(0, 2, n).astype(np.uint64) << np.uint64(63)) | \
                      (rng.integers(0, 7, n).astype(np.uint64) << np.uint64(32)) | rng.integers(0, distinct, n).astype(np.uint64)
            a[:, 1] = np.arange(n, dtype=np.uint64)
            mine = oracle.rs_sort_128x(a)
            theirs = np.ascontiguousarray(a.copy())
            ref.radix_sort_128x(theirs.ctypes.data_as(C.c_void_p), C.c_void_p(theirs.ctypes.data + theirs.nbytes))
            assert (mine == theirs).all(), (n, distinct)
