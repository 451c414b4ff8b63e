"""bin/seq_dump (nextdenovo_b200/csrc/seq_dump_main.cc) against the unmodified reference tool (oracle/_ref/seq_dump, built from
util/seq_dump.c + lib/bseq.c): every output file byte for byte, on inputs that exercise the reader's corners (multi-line FASTA, FASTQ,
gzip, CRLF, blank lines, lower case, N and other letters, reads on every side of the -f / -s / 1 Mb cut-offs, block roll-over, comment
and relative lines in the list, a truncated FASTQ).  The fixed md5s below were minted from the reference tool by this same test
(SEQ_DUMP_MINT=1) so that the check keeps its teeth where oracle/_ref is not built.  CPU only."""
import gzip
import hashlib
import os
import random
import subprocess
import pytest

from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "bin", "seq_dump")
SRC = os.path.join(ROOT, "nextdenovo_b200", "csrc", "seq_dump_main.cc")


@pytest.fixture(scope="module")
def ours():
    if not os.path.exists(OURS) or os.path.getmtime(OURS) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(OURS), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", OURS, SRC, "-lz"])
    return OURS


def _seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def _make_inputs(d):
    rng = random.Random(20240923)
    os.makedirs(os.path.join(d, "sub"), exist_ok=True)
    # 1. multi-line FASTA, mixed case, N runs, IUPAC letters, blank lines, comments on the header
    with open(os.path.join(d, "a.fa"), "w") as f:
        for i in range(60):
            L = rng.choice([10, 499, 500, 501, 1200, 1999, 2000, 2001, 3500, 7000])
            s = list(_seq(rng, L, "ACGTacgt"))
            for _ in range(rng.randrange(0, 6)):
                p = rng.randrange(L); s[p] = rng.choice("NnRYKMUu-* ")
            if i % 7 == 0 and L > 40:
                s[16] = "N"; s[31] = "N"; s[32] = "N"            # first / last base of a word
            s = "".join(s)
            f.write(">read%d some comment\n" % i)
            w = rng.choice([60, 70, 80, 10 ** 6])
            for j in range(0, L, w):
                f.write(s[j:j + w] + "\n")
            if i % 5 == 0:
                f.write("\n")
    # 2. FASTQ, gzip, with '@' and '+' inside quality strings
    with gzip.open(os.path.join(d, "sub", "b.fq.gz"), "wt") as f:
        for i in range(40):
            L = rng.choice([300, 800, 2500, 5000])
            f.write("@q%d\n%s\n+\n%s\n" % (i, _seq(rng, L), "".join(rng.choice("@+I5#>") for _ in range(L))))
    # 3. CRLF FASTA, header without comment, empty sequence, a read over the 1 Mb limit and one just under it
    with open(os.path.join(d, "c.fa"), "wb") as f:
        f.write(b">x\r\nACGTACGTAC\r\nGGGTTT\r\n>empty\r\n>y\r\n" + _seq(rng, 2600).encode() + b"\r\n")
        f.write(b">big\n" + _seq(rng, 1000000).encode() + b"\n>almost\n" + _seq(rng, 999999).encode() + b"\n>tail\n" + _seq(rng, 900).encode())
    # 4. FASTQ whose third record has a short quality string: reading of the file ends there
    with open(os.path.join(d, "d.fq"), "w") as f:
        for i in range(5):
            L = 1500
            q = "I" * (L if i != 2 else L - 7)
            f.write("@t%d\n%s\n+t%d\n%s\n" % (i, _seq(rng, L), i, q))
    with open(os.path.join(d, "in.fofn"), "w") as f:
        f.write("# comment line\n\na.fa\nsub/b.fq.gz\n%s\nd.fq\n" % os.path.join(d, "c.fa"))
    return os.path.join(d, "in.fofn")


def _run(exe, fofn, out, args):
    subprocess.check_call([exe] + args + ["-d", out, fofn], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return {n: open(os.path.join(out, n), "rb").read() for n in sorted(os.listdir(out))}


CASES = {
    "one_seed_file": ["-f", "500", "-s", "2000", "-b", "10g", "-n", "1"],
    "three_seed_files_small_blocks": ["-f", "1k", "-s", "2.5k", "-b", "20k", "-n", "3"],
    "no_block_limit_seed_above_limit": ["-f", "0", "-s", "2000000", "-b", "0", "-n", "2"],
}
GOLD = {"one_seed_file": "cb238e91a409ee026e2e75225d0554c0", "three_seed_files_small_blocks": "10a5ded1c1e0e593b86d99cb11139b34",
        "no_block_limit_seed_above_limit": "e4fa009f134cd5ab2622e9130652baee"}


@pytest.mark.parametrize("case", list(CASES))
def test_seq_dump_equals_reference(ours, tmp_path, case):
    fofn = _make_inputs(str(tmp_path / "in"))
    got = _run(ours, fofn, str(tmp_path / "ours"), CASES[case])
    assert any(n.endswith(".2bit") and len(b) > 2 for n, b in got.items())
    digest = hashlib.md5(b"".join(n.encode() + b"\0" + hashlib.md5(b).digest() for n, b in got.items())).hexdigest()
    ref = oracle.ref_binary("seq_dump")
    if ref:
        want = _run(ref, fofn, str(tmp_path / "ref"), CASES[case])
        assert sorted(got) == sorted(want)
        for n in want:
            assert got[n] == want[n], "%s differs (%d vs %d bytes)" % (n, len(got[n]), len(want[n]))
        if os.environ.get("SEQ_DUMP_MINT"):
            print("GOLD", case, digest)
    if case in GOLD:
        assert digest == GOLD[case]
    elif not ref:
        pytest.skip("no reference binary and no minted digest")


def test_seq_dump_output_is_what_minimap2_nd_reads(ours, tmp_path):
    """the container round trip: the reads come back from the .2bit + .idx pair as they went in (A for every other letter)"""
    import numpy as np
    import nextdenovo_b200 as nd
    fofn = _make_inputs(str(tmp_path / "in"))
    out = str(tmp_path / "o")
    _run(ours, fofn, out, CASES["one_seed_file"])
    reads = nd.read_2bit(os.path.join(out, "input.seed.001.2bit"))
    idx = [l.split("\t") for l in open(os.path.join(out, ".input.seed.001.idx")).read().splitlines()]
    assert len(idx) == reads.n > 10
    for (name, off, ln), rn, rl in zip(idx, reads.ids, reads.length):
        assert int(name) == int(rn) and int(ln) == int(rl) and int(ln) >= 2000
    # the .idx offset is the byte position of the read's first word (ref: util/seq_dump.c:39)
    raw = open(os.path.join(out, "input.seed.001.2bit"), "rb").read()
    off0, w0 = int(idx[0][1]), int(reads.word_off[0])
    assert np.frombuffer(raw[off0:off0 + 8], dtype="<u4").tolist() == reads.words[w0:w0 + 2].tolist()


def test_seq_dump_errors(ours, tmp_path):
    d = str(tmp_path)
    open(os.path.join(d, "l.fofn"), "w").write("missing.fa\n")
    assert subprocess.run([ours, "-f", "1k", "-s", "500", "-b", "1g", "-n", "1", "-d", d + "/o", d + "/l.fofn"], capture_output=True).returncode == 1   # -s <= -f
    assert subprocess.run([ours, "-f", "1k", "-s", "2k", "-b", "1g", "-n", "1", "-d", d + "/o", d + "/l.fofn"], capture_output=True).returncode == 1    # unreadable input
    assert subprocess.run([ours, "-f", "1k", "-s", "2k", "-b", "1g", "-n", "1", "-d", d + "/o", d + "/none.fofn"], capture_output=True).returncode == 1  # no list
    assert subprocess.run([ours], capture_output=True).returncode == 1
