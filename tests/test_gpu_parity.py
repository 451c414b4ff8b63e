"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every test goes through the C ABI
(libndgpu.so) and compares with the oracle or with golden fixtures minted from the unmodified
reference (tests/golden/make_golden.py).  Bar: bit-exact (integer/byte work)."""
import hashlib
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _opts(preset, k=None, w=None):
    import nextdenovo_b200 as nd
    io, mo = nd.IdxOpt(), nd.MapOpt()
    nd.set_opt(preset, io, mo)
    if k: io.k = k
    if w: io.w = w
    mo.step, mo.minlen = 1, 500
    return io, mo


@pytest.fixture(scope="module")
def eng(engine_mod):
    e = engine_mod.Engine(0)
    yield e
    e.close()


def _diff_msg(got, want):
    import nextdenovo_b200 as nd
    try:
        a, b = nd.decode_ovl8(got), nd.decode_ovl8(want)
    except Exception as ex:  # noqa: BLE001
        return "undecodable: %s" % ex
    n = min(len(a), len(b))
    bad = np.nonzero((a[:n] != b[:n]).any(axis=1))[0]
    msg = "records %d vs %d; first differing record %s" % (len(a), len(b), bad[:1])
    if len(bad):
        i = bad[0]
        msg += "\n got  %s\n want %s" % (a[i], b[i])
    return msg


@pytest.mark.parametrize("preset,hpc", [("ava-ont", 0), ("ava-pb", 1)])
def test_sketch_matches_oracle(eng, reads_sub, preset, hpc):
    """ndg_sketch == mm_sketch restatement for every read (ref: minimap2/sketch.c:77-143)."""
    import nextdenovo_b200 as nd
    from nextdenovo_b200.formats import unpack_bases
    from oracle import oracle
    io, mo = _opts(preset)
    eng.set_options(io, mo)
    sub = reads_sub.slice(0, 40)
    xy, off = eng.sketch(sub, rid0=7)
    for r in range(sub.n):
        w0, w1 = int(sub.word_off[r]), int(sub.word_off[r + 1])
        seq = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[unpack_bases(sub.words[w0:w1], int(sub.length[r]))])
        want = oracle.sketch(seq, io.w, io.k, rid=7 + r, hpc=hpc)
        got = xy[int(off[r]):int(off[r + 1])]
        assert got.shape == want.shape and (got == want).all(), "read %d" % r


@pytest.mark.parametrize("variant", [(0, 0), (1, 1)])
@pytest.mark.parametrize("preset,gold", [("ava-ont", "S1_sub120_ava-ont.ovl"), ("ava-pb", "S3_sub120_ava-pb.ovl")])
def test_step1_sub120(eng, engine_mod, reads_sub, golden_dir, preset, gold, variant):
    io, mo = _opts(preset)
    eng.set_variant(*variant)
    eng.set_limits(64 << 20, 0)
    ovl, st = engine_mod.run_step1(eng, reads_sub, reads_sub, io, mo)
    want = open(os.path.join(golden_dir, gold), "rb").read()
    assert ovl == want, _diff_msg(ovl, want)


def test_step1_small_subbatches(eng, engine_mod, reads_sub, golden_dir):
    """device sub-batching must not change the byte stream"""
    io, mo = _opts("ava-ont")
    eng.set_variant(1, 1)
    eng.set_limits(300000, 0)
    ovl, st = engine_mod.run_step1(eng, reads_sub, reads_sub, io, mo)
    eng.set_limits(64 << 20, 0)
    want = open(os.path.join(golden_dir, "S1_sub120_ava-ont.ovl"), "rb").read()
    assert ovl == want, _diff_msg(ovl, want)


def test_step1_two_index_parts_dual(eng, engine_mod, reads_full, reads_sub, golden_dir):
    """-I 20M --dual=yes: two index parts, mid_occ from the first part only (ref: minimap2/main.c:477-509,
    options.c:70-71), NO_DUAL cleared (main.c:328-329)."""
    from nextdenovo_b200.options import MM_F_NO_DUAL
    io, mo = _opts("ava-ont")
    mo.flag &= ~MM_F_NO_DUAL
    io.batch_size = 20000000
    eng.set_variant(1, 1)
    ovl, st = engine_mod.run_step1(eng, reads_full, reads_sub, io, mo)
    want = open(os.path.join(golden_dir, "S5_sub120_dual.ovl"), "rb").read()
    assert ovl == want, _diff_msg(ovl, want)


@pytest.mark.parametrize("preset,k,w,gold,md5", [
    ("ava-ont", None, None, "G1_ava-ont.ovl", "9fce14e8841db2e46012c58ad8f0059c"),
    ("ava-ont", 17, 10, "G2_ava-ont_k17w10.ovl", "270d357dfda1e890aa8330b316ffcf1f"),
    ("ava-pb", None, None, "G3_ava-pb.ovl", "7fb430f9dbd8032c43e3613e6d35680e"),
])
def test_step1_full_golden(eng, engine_mod, reads_full, golden_dir, preset, k, w, gold, md5):
    """test_data/reads_test.fa.gz self-overlap: the .ovl must be byte-identical to the reference's
    (BASELINE.json config 0 / north-star parity target)."""
    io, mo = _opts(preset, k, w)
    eng.set_variant(1, 1)
    ovl, st = engine_mod.run_step1(eng, reads_full, reads_full, io, mo)
    want = open(os.path.join(golden_dir, gold), "rb").read()
    assert hashlib.md5(want).hexdigest() == md5
    assert ovl == want, _diff_msg(ovl, want)


def test_step1_full_baseline_kernels(eng, engine_mod, reads_full, golden_dir):
    """the one-thread-per-read/segment kernel generation gives the same bytes"""
    io, mo = _opts("ava-ont", 17, 10)
    eng.set_variant(0, 0)
    ovl, st = engine_mod.run_step1(eng, reads_full, reads_full, io, mo)
    eng.set_variant(1, 1)
    want = open(os.path.join(golden_dir, "G2_ava-ont_k17w10.ovl"), "rb").read()
    assert ovl == want, _diff_msg(ovl, want)


def test_index_stats(eng, reads_full):
    """ref stderr: 'distinct minimizers: 6273819 (79.81% are singletons)', mid_occ = 74 for ava-ont"""
    io, mo = _opts("ava-ont")
    eng.set_options(io, mo)
    st = eng.build_index(reads_full)
    assert st["n_distinct"] == 6273819 and st["mid_occ"] == 74
    assert abs(100.0 * st["n_singletons"] / st["n_distinct"] - 79.81) < 0.01


def test_library_loaded_is_in_tree(engine_mod):
    p = engine_mod.lib_path()
    assert os.path.exists(p) and "nextdenovo_b200" in p


def test_cli_dropin_matches_reference_bytes(built, golden_dir, tmp_path):
    """bin/minimap2-nd (C++ host over the C ABI) with the pipeline's own command line
    (ref: nextDenovo:445-446) writes the reference's bytes; bad usage exits 1."""
    import subprocess
    root = os.path.dirname(golden_dir.rstrip("/")).rsplit("/tests", 1)[0]
    exe = os.path.join(root, "bin", "minimap2-nd")
    src = os.path.join(golden_dir, "sub120.2bit")
    out = str(tmp_path / "o.ovl")
    r = subprocess.run([exe, "--step", "1", "-t", "8", "-x", "ava-ont", src, src, "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out, "rb").read() == open(os.path.join(golden_dir, "S1_sub120_ava-ont.ovl"), "rb").read()
    assert "mid_occ" in r.stderr
    full = os.path.join(golden_dir, "reads_test.seed.2bit")
    r = subprocess.run([exe, "--step", "1", "--dual=yes", "-I", "20M", "-t", "8", "-x", "ava-ont", full, src, "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out, "rb").read() == open(os.path.join(golden_dir, "S5_sub120_dual.ovl"), "rb").read()
    assert subprocess.run([exe, "--step", "1", "-x", "nope", src, src, "-o", out], capture_output=True).returncode == 1
    assert subprocess.run([exe, "--step", "1", "-c", "-x", "ava-ont", src, src, "-o", out], capture_output=True).returncode == 1


@pytest.mark.parametrize("ksw_variant", [1, 0])
def test_ksw_extd2_batch_bit_exact(eng, ksw_variant):
    """ndg_ksw_extd2_batch == ksw_extd2_sse (via the oracle restatement, and the compiled reference when present):
    score, max, end points, z-drop flag, reach_end and CIGAR, for every call of a mixed batch."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from ksw_cases import simple_mat, oracle_lib, oracle_call, ref_lib, ref_call, random_cases
    from nextdenovo_b200.engine import KSW_CALL_DTYPE
    mat = simple_mat()
    cases = random_cases(1500, seed=23)
    calls = np.zeros(len(cases), dtype=KSW_CALL_DTYPE)
    chunks, off = [], 0
    for i, (qs, ts, w, zd, eb, flag) in enumerate(cases):
        calls[i] = (len(qs), len(ts), off, off + len(qs), w, zd, eb, flag)
        chunks += [qs, ts]; off += len(qs) + len(ts)
    seqs = np.concatenate(chunks)
    eng.set_ksw_variant(ksw_variant)
    res, cigs = eng.ksw_extd2_batch(calls, seqs, mat, 4, 2, 24, 1)
    eng.set_ksw_variant(1)
    L, R = oracle_lib(), ref_lib()
    names = ("max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar", "reach_end")
    bad = []
    for i, (qs, ts, w, zd, eb, flag) in enumerate(cases):
        want = oracle_call(L, qs, ts, mat, 4, 2, 24, 1, w, zd, eb, flag)
        if R is not None and i % 5 == 0:
            assert ref_call(R, qs, ts, mat, 4, 2, 24, 1, w, zd, eb, flag) == want
        got = (tuple(int(res[n][i]) for n in names), [int(x) for x in cigs[i]])
        if got != want:
            bad.append((i, len(qs), len(ts), w, hex(flag), got[0], want[0]))
    assert not bad, "%d/%d calls differ; first: %s" % (len(bad), len(cases), bad[:3])


@pytest.mark.parametrize("w,k,hpc", [(5, 15, 0), (10, 17, 0), (7, 16, 0), (1, 11, 0), (19, 19, 0), (5, 19, 1), (70, 21, 0)])
def test_sketch_adversarial_reads(eng, w, k, hpc):
    """fast closed-form path + state-machine path vs mm_sketch restatement on homopolymers, short-period
    repeats (ties, palindromic k-mers for even k), reads around the one-window length, duplicated segments"""
    import nextdenovo_b200 as nd
    from nextdenovo_b200.formats import pack_bases, make_reads
    from oracle import oracle
    rng = np.random.default_rng(10)
    seqs = []
    for i in range(600):
        L = int(rng.integers(max(k - 3, 1), k + w + 60))
        mode = i % 4
        if mode == 0: s = rng.integers(0, 2, L)
        elif mode == 1: s = np.tile(rng.integers(0, 4, int(rng.integers(1, 6))), L)[:L]
        elif mode == 2: s = rng.integers(0, 4, L)
        else:
            s = rng.integers(0, 4, L); a = int(rng.integers(0, max(L - 5, 1))); s[a:a + int(rng.integers(3, 40))] = int(rng.integers(0, 4))
        seqs.append(s.astype(np.uint8))
    seqs += [np.zeros(3000, np.uint8), np.tile(np.array([0, 3], np.uint8), 1500), np.tile(np.array([0, 1, 2, 3], np.uint8), 700)]
    base = rng.integers(0, 4, 700).astype(np.uint8)
    seqs.append(np.concatenate([base, base, rng.integers(0, 4, 50).astype(np.uint8), base]))
    seqs += [rng.integers(0, 4, int(rng.integers(400, 4000))).astype(np.uint8) for _ in range(30)]
    reads = make_reads([pack_bases(s) for s in seqs], [len(s) for s in seqs], list(range(len(seqs))))
    io, mo = _opts("ava-ont", k, w)
    if hpc: io.flag |= 1
    eng.set_options(io, mo)
    xy, off = eng.sketch(reads, rid0=3)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    for r, s in enumerate(seqs):
        want = oracle.sketch(bytes(lut[s]), w, k, rid=3 + r, hpc=hpc)
        got = xy[int(off[r]):int(off[r + 1])]
        assert got.shape == want.shape and (got == want).all(), "read %d len %d" % (r, len(s))
