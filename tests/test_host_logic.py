"""Host-side logic: formats, option parsing, batching.  CPU only."""
import os
import numpy as np
import pytest
import nextdenovo_b200 as nd
from nextdenovo_b200 import options
from nextdenovo_b200.formats import pack_bases, unpack_bases, make_reads


def test_2bit_roundtrip(tmp_path, reads_sub, golden_dir):
    p = str(tmp_path / "x.2bit")
    nd.write_2bit(p, reads_sub, str(tmp_path / "x.idx"))
    assert open(p, "rb").read() == open(os.path.join(golden_dir, "sub120.2bit"), "rb").read()
    r2 = nd.read_2bit(p)
    assert (r2.words == reads_sub.words).all() and (r2.length == reads_sub.length).all() and (r2.ids == reads_sub.ids).all()
    first = open(str(tmp_path / "x.idx")).readline().split("\t")
    assert first[0] == "0" and first[1] == "10" and int(first[2]) == int(reads_sub.length[0])   # util/seq_dump.c:36-41


def test_pack_unpack():
    rng = np.random.default_rng(1)
    for n in (1, 15, 16, 17, 1000):
        c = rng.integers(0, 4, n).astype(np.uint8)
        assert (unpack_bases(pack_bases(c), n) == c).all()
    assert pack_bases(np.array([3, 0, 1, 2], dtype=np.uint8))[0] == (3 << 30 | 0 << 28 | 1 << 26 | 2 << 24)


def test_ovl_codec_against_golden(golden_dir):
    buf = open(os.path.join(golden_dir, "S1_sub120_ava-ont.ovl"), "rb").read()
    rec = nd.decode_ovl8(buf)
    assert rec.shape == (7272, 8)
    assert nd.encode_ovl8_py([tuple(int(x) for x in r) for r in rec]) == buf
    assert (rec[:, 3] > rec[:, 2]).all() and (rec[:, 6] > rec[:, 5]).all() and set(np.unique(rec[:, 1])) <= {0, 1}


def test_varint_edges():
    recs = [(0, 0, 0, 127, 0, 0, 127, 0), (128, 1, 128, 16383 + 128, 5, 16384, 16384 + 2097152, 268435456),
            (4294967295, 0, 99999999, 100000001, 4294967295, 7, 4294967290, 100000000), (3, 1, 1, 2, 1, 2, 100, 4294967295)]
    got = nd.decode_ovl8(nd.encode_ovl8_py(recs))
    assert [tuple(int(x) for x in r) for r in got] == recs


def test_name_keys_follow_strcmp():
    ids = np.array([0, 1, 9, 10, 11, 99, 100, 101, 1000, 123456789, 4294967295, 2, 20, 200], dtype=np.uint32)
    keys = nd.name_keys(ids)
    names = [str(int(i)) for i in ids]
    for i in range(len(ids)):
        for j in range(len(ids)):
            assert (names[i] < names[j]) == (int(keys[i]) < int(keys[j]))
            assert (names[i] == names[j]) == (int(keys[i]) == int(keys[j]))


def test_x31_hash():
    ids = np.array([0, 7, 10, 99, 12345, 4294967295], dtype=np.uint32)
    def x31(s):
        h = ord(s[0])
        for ch in s[1:]:
            h = (h * 31 + ord(ch)) & 0xffffffff
        return h
    assert [int(h) for h in nd.x31_hash(ids)] == [x31(str(int(i))) for i in ids]


def test_parse_num():
    assert nd.parse_num("3G") == 3000000000 and nd.parse_num("500m") == 500000000 and nd.parse_num("2.5k") == 2500 and nd.parse_num("17") == 17


def test_parse_args_pipeline_commands():
    """command lines emitted by nextDenovo:445-464 (raw_align)"""
    io, mo, out, pos = options.parse_args("--step 1 --dual=yes -t 8 -x ava-ont seed.2bit part.2bit -o x.ovl".split())
    assert (io.k, io.w, io.flag & options.MM_I_HPC) == (15, 5, 0)
    assert mo.step == 1 and mo.minlen == 500 and mo.outraw == 0 and out == "x.ovl" and pos == ["seed.2bit", "part.2bit"]
    assert not (mo.flag & options.MM_F_NO_DUAL) and (mo.flag & options.MM_F_NO_DIAG)
    assert (mo.bw, mo.max_gap, mo.min_chain_score) == (2000, 10000, 100)
    io, mo, out, pos = options.parse_args("-k 17 --step 1 -I 3G -t 8 -x ava-pb a b -o y".split())     # -x applies first
    assert (io.k, io.w) == (17, 5) and io.flag & options.MM_I_HPC and io.batch_size == 3000000000
    assert mo.flag & options.MM_F_NO_DUAL and mo.bw == 500
    with pytest.raises(ValueError):
        options.parse_args("--step 1 -x nonsense a b".split())
    with pytest.raises(ValueError):
        options.parse_args("--step 7 a b".split())


def test_index_parts_and_query_batches():
    from nextdenovo_b200.engine import index_parts, query_batches, shard_bounds
    lens = [30] * 10
    r = make_reads([np.zeros(2, np.uint32)] * 10, lens, list(range(10)))
    io = nd.IdxOpt()
    io.batch_size, io.mini_batch_size = 100, 50                  # mini-batches of >=50 bases (2 reads), part ends once >100
    assert index_parts(r, io) == [(0, 4), (4, 8), (8, 10)]
    io.batch_size = 10 ** 9
    assert index_parts(r, io) == [(0, 10)]
    assert query_batches(r, 100) == [(0, 4), (4, 8), (8, 10)]
    assert shard_bounds(10, 4) == [0, 2, 5, 7, 10]


def test_fastx_reader_and_mask(tmp_path):
    """FASTA / FASTQ(.gz) parsing, ambiguity mask layout (bit i&31 of nmask[word_off/2 + i>>5]), format sniffing"""
    import gzip
    from nextdenovo_b200.formats import unpack_bases
    fa = tmp_path / "t.fa"
    fa.write_bytes(b">3 c\nACGTNN\nacgu\n>4\n\n>5\n" + b"N" * 40 + b"\n" + b"ACGT" * 20 + b"\n")
    a = nd.read_seqfile(str(fa))
    assert a.n == 3 and list(a.length) == [10, 0, 120] and list(a.ids) == [3, 4, 5]
    assert (a.word_off % 2 == 0).all()
    assert int(a.nmask[0]) == 0b110000 and int(a.nmask[int(a.word_off[2]) // 2]) == 0xffffffff and int(a.nmask[int(a.word_off[2]) // 2 + 1]) == 0xff
    assert list(unpack_bases(a.words[:2], 10)) == [0, 1, 2, 3, 0, 0, 0, 1, 2, 3]
    fq = tmp_path / "t.fq.gz"
    with gzip.open(str(fq), "wb") as f:
        f.write(b"@7\nACGTN\n+\nIIII@\n@8 x\nGGG\n+\n@@@\n")
    b = nd.read_seqfile(str(fq))
    assert b.n == 2 and list(b.length) == [5, 3] and list(b.ids) == [7, 8]
    bad = tmp_path / "bad.fa"
    bad.write_bytes(b">chr1\nACGT\n")
    with pytest.raises(ValueError):
        nd.read_seqfile(str(bad))
    s = a.slice(2, 3)
    assert s.n == 1 and int(s.nmask[0]) == 0xffffffff


def test_decode_ovl10_matches_ovl_cvt(golden_dir):
    """10-field records of --step 2 (ref: lib/ovl.c:150-253, util/ovl_cvt.c:33-50) decoded by formats.decode_ovl10 == ovl_cvt -m 1"""
    import os
    import subprocess
    import numpy as np
    import nextdenovo_b200 as nd
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cvt = os.path.join(root, "oracle", "_ref", "ovl_cvt")
    f = os.path.join(golden_dir, "G5_step2.ovl")
    a = nd.decode_ovl10(open(f, "rb").read())
    assert a.shape == (7635, 10)
    assert a[0].tolist() == [0, 0, 1929, 67698, 1, 31, 67777, 0, 85492, 2178]   # qlen 0: read 0 equals the initial "previous" name
    if os.path.exists(cvt):
        txt = subprocess.run([cvt, "-m", "1", f], capture_output=True, text=True).stdout
        b = np.array([[int(x) for x in line.split("\t")] for line in txt.strip().split("\n")], dtype=np.uint32)
        assert (a == b).all()


def test_option_grammar_is_the_references():
    """unknown options are an error (ref: minimap2/main.c:186-188), options that are not built are refused by name, attached and
    bundled short forms parse as with ketopt"""
    from nextdenovo_b200 import options
    import pytest as _pt
    with _pt.raises(ValueError, match="unknown option"):
        options.parse_args("--step 1 -x ava-ont --bogus a b -o x".split())
    with _pt.raises(ValueError, match="unknown option"):
        options.parse_args("--step 1 -x ava-ont -J a b -o x".split())
    for bad in ("-a", "--cs", "--splice", "--sr"):
        with _pt.raises(ValueError, match="not built"):
            options.parse_args(("--step 1 -x ava-ont %s a b -o x" % bad).split())
    io, mo, out, pos = options.parse_args("--step=1 -cxava-ont -k17 -w 10 --min-count 4 --mb-size 60M a b -ox.ovl".split())
    assert (io.k, io.w, mo.min_cnt, mo.mini_batch_size, out, pos) == (17, 10, 4, 60000000, "x.ovl", ["a", "b"])
    assert mo.flag & options.MM_F_CIGAR and mo.flag & options.MM_F_ALL_CHAINS
    io, mo, out, pos = options.parse_args("--step 1 -k 15 a b -o x".split())       # no preset: not all-chain
    assert not mo.flag & options.MM_F_ALL_CHAINS


def test_cli_rejects_what_it_does_not_implement(tmp_path):
    """bin/minimap2-nd fails with status 1 before touching a GPU: unknown option, missing all-chain mode, PAF/SAM output"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "bin", "minimap2-nd")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    for args, msg in ((["--step", "1", "-x", "ava-ont", "--bogus", "a", "b", "-o", "x"], "unknown option"),
                      (["--step", "1", "-k", "15", "a", "b", "-o", str(tmp_path / "x")], "all-chain"),
                      (["--step", "1", "-x", "ava-ont", "-a", "a", "b", "-o", "x"], "not built"),
                      (["--step", "1", "-x", "ava-ont", "a", "b", "-o"], "missing option argument")):
        r = subprocess.run([exe] + args, capture_output=True, text=True)
        assert r.returncode == 1 and msg in r.stderr, (args, r.stderr)
