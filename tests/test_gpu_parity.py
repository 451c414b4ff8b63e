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


@pytest.mark.parametrize("variant", [(0, 0), (1, 1), (2, 1)])
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


@pytest.mark.parametrize("variant", [(0, 0), (2, 1)])
def test_step1_full_baseline_kernels(eng, engine_mod, reads_full, golden_dir, variant):
    """the one-thread-per-read/segment kernel generation, and the exact sort replay for every read (no block-sort fast path), give the
    same bytes"""
    io, mo = _opts("ava-ont", 17, 10)
    eng.set_variant(*variant)
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
    assert subprocess.run([exe, "--step", "3", "-x", "ava-ont", src, src, "-o", out], capture_output=True).returncode == 1
    assert subprocess.run([exe, "--step", "1", "-c", "-O", "4,2", "-x", "ava-ont", src, src, "-o", out], capture_output=True).returncode == 1


@pytest.mark.parametrize("preset,gold", [("ava-ont", "S6_sub120_ava-ont_c.ovl"), ("ava-pb", "S7_sub120_ava-pb_c.ovl")])
def test_step1_cigar_sub120(eng, engine_mod, reads_sub, golden_dir, preset, gold):
    """regime B (`-c`): chains -> ksw_extd2 on the GPU -> stitched alignments; the .ovl bytes (coordinates and match counts of the
    gapped alignments, hit order by dp_max) equal the reference's (ref: minimap2/align.c:565-913, map.c:484-505)."""
    from nextdenovo_b200.options import MM_F_CIGAR
    io, mo = _opts(preset)
    mo.flag |= MM_F_CIGAR
    eng.set_variant(1, 1)
    eng.set_limits(64 << 20, 0)
    ovl, st = engine_mod.run_step1(eng, reads_sub, reads_sub, io, mo)
    mo.flag &= ~MM_F_CIGAR
    eng.set_options(io, mo)                      # switch regime B off again for the tests that follow
    want = open(os.path.join(golden_dir, gold), "rb").read()
    assert ovl == want, _diff_msg(ovl, want)
    s = eng.aln_stats()
    assert s["n_chains"] > 7000 and s["n_calls"] > 300000 and s["n_launches"] >= s["n_rounds"]


def test_g4_full_cigar_cli(built, golden_dir, tmp_path):
    """The north star's parity target: `--step 1 -c -x ava-ont` on all 997 reads of test_data/reads_test.fa.gz must give the reference's
    133,100 records byte for byte (ref: minimap2/align.c:565-913 + ksw2_extd2_sse.c on 5.6 M DP calls)."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bin", "minimap2-nd")
    src = os.path.join(golden_dir, "reads_test.seed.2bit")
    out = str(tmp_path / "g4.ovl")
    r = subprocess.run([exe, "--step", "1", "-c", "-x", "ava-ont", src, src, "-o", out], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    got = open(out, "rb").read()
    want = open(os.path.join(golden_dir, "G4_ava-ont_c.ovl"), "rb").read()
    assert hashlib.md5(want).hexdigest() == "4c7e852512ee6fb463b37a0ce20926b7"
    assert got == want, _diff_msg(got, want)


def test_cli_cigar_matches_reference_bytes(built, golden_dir, tmp_path):
    """bin/minimap2-nd --step 1 -c: same bytes as the reference binary (golden S6)"""
    import subprocess
    root = os.path.dirname(golden_dir.rstrip("/")).rsplit("/tests", 1)[0]
    exe = os.path.join(root, "bin", "minimap2-nd")
    src = os.path.join(golden_dir, "sub120.2bit")
    out = str(tmp_path / "o.ovl")
    r = subprocess.run([exe, "--step", "1", "-t", "8", "-c", "-x", "ava-ont", src, src, "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out, "rb").read() == open(os.path.join(golden_dir, "S6_sub120_ava-ont_c.ovl"), "rb").read()
    assert "DP calls" in r.stderr


@pytest.mark.parametrize("ksw_variant", [1, 2, 3])
def test_ksw_extd2_wide_bands_bit_exact(eng, ksw_variant):
    """long calls with wide bands through the warp-per-call (1), the block-per-call (2) and the register-resident (3: rings of 3072/4096 columns,
    sliding bands, fallback for bands wider than every ring) kernels == the reference function"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from ksw_cases import simple_mat, oracle_lib, oracle_call, ref_lib, ref_call, wide_cases
    from nextdenovo_b200.engine import KSW_CALL_DTYPE
    mat = simple_mat()
    cases = wide_cases(48)
    calls = np.zeros(len(cases), dtype=KSW_CALL_DTYPE)
    chunks, off = [], 0
    for i, (qs, ts, w, zd, eb, flag) in enumerate(cases):
        calls[i] = (len(qs), len(ts), off, off + len(qs), w, zd, eb, flag)
        chunks += [qs, ts]; off += len(qs) + len(ts)
    eng.set_ksw_variant(ksw_variant)
    res, cigs = eng.ksw_extd2_batch(calls, np.concatenate(chunks), mat, 4, 2, 24, 1)
    eng.set_ksw_variant(3)
    L, R = oracle_lib(), ref_lib()
    names = ("max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar", "reach_end")
    bad = []
    for i, (qs, ts, w, zd, eb, flag) in enumerate(cases):
        want = ref_call(R, qs, ts, mat, 4, 2, 24, 1, w, zd, eb, flag) if R is not None else oracle_call(L, qs, ts, mat, 4, 2, 24, 1, w, zd, eb, flag)
        got = (tuple(int(res[n][i]) for n in names), [int(x) for x in cigs[i]])
        if got != want:
            bad.append((i, len(qs), len(ts), w, hex(flag), got[0], want[0]))
    assert not bad, "%d/%d calls differ; first: %s" % (len(bad), len(cases), bad[:3])


@pytest.mark.parametrize("ksw_variant", [1, 0, 2, 3])
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
    eng.set_ksw_variant(3)
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


def _oracle_ovl(tmp_path, T, Q, preset, **kw):
    import nextdenovo_b200 as nd
    from oracle import oracle
    pt, pq = str(tmp_path / "t.2bit"), str(tmp_path / "q.2bit")
    nd.write_2bit(pt, T)
    nd.write_2bit(pq, Q)
    o = oracle.make_opt(preset, **kw)
    out = str(tmp_path / "o.ovl")
    oracle.step1_files(pt, pt if Q is T else pq, o, out)
    return open(out, "rb").read()


def test_edge_cases_vs_oracle(eng, engine_mod, tmp_path):
    """ragged and degenerate inputs: reads shorter than k / than one window, a zero-length read, exact duplicates
    (every anchor tied), a 300 kb read, low-complexity reads, queries without any hit -- byte-equal to the oracle"""
    import nextdenovo_b200 as nd
    from nextdenovo_b200 import synth
    from nextdenovo_b200.formats import pack_bases, make_reads, unpack_bases
    rng = np.random.default_rng(77)
    g = synth.genome(400000, 5)
    base = synth.reads(g, 60, 20000, seed=5)
    seqs = []
    for r in range(base.n):
        w0, w1 = int(base.word_off[r]), int(base.word_off[r + 1])
        seqs.append(unpack_bases(base.words[w0:w1], 20000))
    seqs += [seqs[3].copy(), seqs[3].copy(), seqs[10][:9000].copy()]              # duplicates and a contained read
    seqs += [rng.integers(0, 4, n).astype(np.uint8) for n in (1, 5, 14, 15, 18, 19, 20, 33, 600)]   # too short / unrelated
    seqs.append(np.zeros(0, dtype=np.uint8))                                         # empty read
    seqs.append(np.concatenate([g[1000:151000], g[160000:310000]]).astype(np.uint8))  # 300 kb read spanning many others
    seqs.append(np.zeros(5000, dtype=np.uint8)); seqs.append(np.tile(np.array([0, 1], np.uint8), 3000))
    order = rng.permutation(len(seqs))
    seqs = [seqs[i] for i in order]
    ids = list(range(100, 100 + len(seqs)))                                         # names "100".."1xx": strcmp order != numeric for mixed widths
    ids[5], ids[7] = 9, 1000
    reads = make_reads([pack_bases(s) if len(s) else np.zeros(0, np.uint32) for s in seqs], [len(s) for s in seqs], ids)
    for preset, kw in (("ava-ont", {}), ("ava-pb", {}), ("ava-ont", dict(k=17, w=10))):
        io, mo = _opts(preset, kw.get("k"), kw.get("w"))
        for variant in ((1, 1), (0, 0)):
            eng.set_variant(*variant)
            ovl, st = engine_mod.run_step1(eng, reads, reads, io, mo)
            want = _oracle_ovl(tmp_path, reads, reads, preset, **kw)
            assert ovl == want, "%s %s %s: %s" % (preset, kw, variant, _diff_msg(ovl, want))
            assert st["records"] > 50
    eng.set_variant(1, 1)
    # queries that hit nothing
    lonely = make_reads([pack_bases(rng.integers(0, 4, 3000).astype(np.uint8)) for _ in range(4)], [3000] * 4, [5000, 5001, 5002, 5003])
    io, mo = _opts("ava-ont")
    ovl, st = engine_mod.run_step1(eng, reads, lonely, io, mo)
    assert ovl == b"" and st["records"] == 0


def test_fastx_input_with_ambiguous_bases(built, engine_mod, eng, reads_sub, tmp_path):
    """FASTA.gz input (decimal names, N runs, lower case): the drop-in executable and the Python path both equal the
    unmodified reference run on the same file (ambiguous bases reset the sketch window, ref: minimap2/sketch.c:112)."""
    import gzip
    import subprocess
    import nextdenovo_b200 as nd
    from nextdenovo_b200.formats import unpack_bases
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = os.path.join(root, "oracle", "_ref", "minimap2-nd")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not present")
    rng = np.random.default_rng(4)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    fa = str(tmp_path / "r.fa.gz")
    with gzip.open(fa, "wb") as fh:
        for r in range(70):
            w0, w1 = int(reads_sub.word_off[r]), int(reads_sub.word_off[r + 1])
            s = bytearray(lut[unpack_bases(reads_sub.words[w0:w1], int(reads_sub.length[r]))].tobytes())
            for _ in range(int(rng.integers(0, 6))):
                p, ln = int(rng.integers(0, len(s))), int(rng.integers(1, 30))
                s[p:p + ln] = b"N" * len(s[p:p + ln])
            if r % 3 == 0:
                s = bytearray(bytes(s).lower())
            fh.write(b">%d some comment\n" % int(reads_sub.ids[r]))
            for i in range(0, len(s), 70):
                fh.write(bytes(s[i:i + 70]) + b"\n")
    want_p = str(tmp_path / "ref.ovl"); got_p = str(tmp_path / "gpu.ovl")
    subprocess.run([ref, "--step", "1", "-t", "8", "-x", "ava-ont", fa, fa, "-o", want_p], check=True, capture_output=True)
    exe = os.path.join(root, "bin", "minimap2-nd")
    r = subprocess.run([exe, "--step", "1", "-t", "8", "-x", "ava-ont", fa, fa, "-o", got_p], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = open(want_p, "rb").read()
    assert len(want) > 10000 and open(got_p, "rb").read() == want
    reads = nd.read_seqfile(fa)
    assert reads.nmask is not None and reads.n == 70
    io, mo = _opts("ava-ont")
    ovl, st = engine_mod.run_step1(eng, reads, reads, io, mo)
    assert ovl == want, _diff_msg(ovl, want)
    # a name that is not a decimal integer is an error, not garbage ids
    bad = str(tmp_path / "bad.fa")
    open(bad, "wb").write(b">read_x\nACGTACGTACGT\n")
    assert subprocess.run([exe, "--step", "1", "-x", "ava-ont", bad, bad, "-o", got_p], capture_output=True).returncode == 1


def _step2_opts():
    from nextdenovo_b200.options import parse_args
    io, mo, out, pos = parse_args(["--step", "2", "-t", "8", "-x", "ava-ont", "-k", "17", "-w", "17", "--minlen", "2000", "--maxhan1", "5000", "a", "b", "-o", "x"])
    return io, mo


@pytest.mark.parametrize("src,gold", [("sub120.2bit", "S9_sub120_step2.ovl"), ("reads_test.seed.2bit", "G5_step2.ovl")])
def test_step2_matches_reference(eng, engine_mod, golden_dir, src, gold):
    """--step 2 (cns_align): first pass, re-mapping of every dovetail candidate against a (17,10) mini index of the query (all mini
    indexes of a batch share one tagged GPU index), 10-field records and the .bl text == the reference's bytes
    (ref: minimap2/map.c:988-1128,1305-1328; lib/ovl.c:339-361,449-563)."""
    import nextdenovo_b200 as nd
    reads = nd.read_2bit(os.path.join(golden_dir, src))
    io, mo = _step2_opts()
    eng.set_variant(1, 1)
    eng.set_limits(64 << 20, 0)
    ovl, bl, st = engine_mod.run_step2(eng, reads, reads, io, mo)
    want = open(os.path.join(golden_dir, gold), "rb").read()
    assert ovl == want, "records %d; %d vs %d bytes" % (st["records"], len(ovl), len(want))
    assert bl == open(os.path.join(golden_dir, gold + ".bl"), "rb").read()
    io1, mo1 = _opts("ava-ont")
    eng.set_options(io1, mo1)


def test_cli_step2_matches_reference_bytes(built, golden_dir, tmp_path):
    import subprocess
    root = os.path.dirname(golden_dir.rstrip("/")).rsplit("/tests", 1)[0]
    exe = os.path.join(root, "bin", "minimap2-nd")
    src = os.path.join(golden_dir, "sub120.2bit")
    out = str(tmp_path / "o.ovl")
    r = subprocess.run([exe, "-I", "6G", "--step", "2", "-t", "8", "-x", "ava-ont", "-k", "17", "-w", "17", "--minlen", "2000", "--maxhan1", "5000", src, src, "-o", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out, "rb").read() == open(os.path.join(golden_dir, "S9_sub120_step2.ovl"), "rb").read()
    assert open(out + ".bl", "rb").read() == open(os.path.join(golden_dir, "S9_sub120_step2.ovl.bl"), "rb").read()
    assert subprocess.run([exe, "--step", "2", "--mode", "1", "-x", "ava-ont", src, src, "-o", out], capture_output=True).returncode == 1


@pytest.mark.parametrize("args", [
    ["--step", "1", "-c", "-x", "ava-ont"],
    ["--step", "1", "-c", "--dual=yes", "-I", "300k", "-x", "ava-ont"],
    ["--step", "1", "-c", "-x", "ava-pb"],
    ["--step", "1", "-c", "-z", "200,100", "-s", "40", "--end-bonus", "5", "-x", "ava-ont"],
    ["-I", "6G", "--step", "2", "-x", "ava-ont", "-k", "17", "-w", "17", "--minlen", "2000", "--maxhan1", "5000"],
    ["-I", "400k", "--step", "2", "--dual=yes", "-x", "ava-ont", "-k", "17", "-w", "17", "--minlen", "1000", "--maxhan1", "3000", "--minide", "0.03"],
    ["--step", "2", "--mode", "0", "-x", "ava-ont", "-k", "17", "-w", "17"],
    ["--step", "2", "--outctn", "--cn", "7", "-K", "200k", "-x", "ava-ont", "-k", "15", "-w", "10", "--minlen", "500"],
], ids=["c", "c-dual-parts", "c-hpc", "c-scoring", "step2", "step2-parts-dual", "step2-mode0", "step2-outctn-batches"])
def test_cli_equals_reference_on_fastx_with_n(built, reads_sub, tmp_path, args):
    """the drop-in executable and the unmodified reference run side by side on a FASTA file with N runs and lower case (40 reads):
    regime B and step 2, index parts, query batches, --dual, non-default scoring: same .ovl (and .bl) bytes"""
    import subprocess
    from nextdenovo_b200.formats import unpack_bases
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = os.path.join(root, "oracle", "_ref", "minimap2-nd")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not present")
    rng = np.random.default_rng(9)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    fa = str(tmp_path / "r.fa")
    with open(fa, "wb") as fh:
        for r in range(40):
            w0, w1 = int(reads_sub.word_off[r]), int(reads_sub.word_off[r + 1])
            s = bytearray(lut[unpack_bases(reads_sub.words[w0:w1], int(reads_sub.length[r]))].tobytes())
            for _ in range(int(rng.integers(0, 5))):
                p, ln = int(rng.integers(0, len(s))), int(rng.integers(1, 40))
                s[p:p + ln] = b"N" * len(s[p:p + ln])
            if r % 4 == 0:
                s = bytearray(bytes(s).lower())
            fh.write(b">%d\n" % int(reads_sub.ids[r]) + bytes(s) + b"\n")
    want_p, got_p = str(tmp_path / "ref.ovl"), str(tmp_path / "gpu.ovl")
    subprocess.run([ref] + args + ["-t", "8", fa, fa, "-o", want_p], check=True, capture_output=True)
    exe = os.path.join(root, "bin", "minimap2-nd")
    r = subprocess.run([exe] + args + ["-t", "8", fa, fa, "-o", got_p], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = open(want_p, "rb").read()
    assert len(want) > 100 and open(got_p, "rb").read() == want
    if "2" in args[:4] and "--step" in args:
        if os.path.exists(want_p + ".bl"):
            assert open(got_p + ".bl", "rb").read() == open(want_p + ".bl", "rb").read()
