"""`-x ava-hifi` (the preset NextDenovo uses for read_type = hifi, ref: lib/config_parser.py:46-47,205-212): long k-mers (k = w = 51,
homopolymer-compressed, ref: minimap2/sketch.c:283-356 incl. the wrapped 32-entry run queue), own scoring; the executable against the
unmodified reference binary, byte for byte (run on the B200 box)."""
import os
import subprocess
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "bin", "minimap2-nd")
REF = os.path.join(ROOT, "oracle", "_ref", "minimap2-nd")


SEED = 20240923


@pytest.fixture(scope="module")
def hifi_reads(tmp_path_factory, built):
    """1500 synthetic HiFi-like reads: 15 kb, 1 % error, 30x"""
    import nextdenovo_b200 as nd
    from nextdenovo_b200 import synth
    d = tmp_path_factory.mktemp("hifi")
    g = synth.genome(1500 * 15000 // 30, SEED)
    reads = synth.reads(g, 1500, 15000, sub=.004, ins=.003, dele=.003, seed=SEED)
    p = str(d / "hifi.2bit")
    nd.write_2bit(p, reads)
    return p


def _both(args, tmp_path, files=("o.ovl",), min_bytes=20000):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref not present")
    outs = []
    for exe, tag in ((EXE, "g"), (REF, "r")):
        d = tmp_path / tag; d.mkdir()
        r = subprocess.run([exe] + args + ["-o", str(d / "o.ovl")], capture_output=True, text=True, timeout=1800)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append([open(d / f, "rb").read() for f in files])
    assert len(outs[1][0]) > min_bytes
    assert outs[0] == outs[1]
    return outs[0]


@pytest.mark.parametrize("extra", [[], ["-f", "200"], ["-k", "40", "-w", "20"], ["-k", "64", "-w", "30", "-H"], ["--dual=yes", "-I", "8M"]],
                         ids=["preset", "f200", "k40w20", "k64w30", "dual-parts"])
def test_step1_ava_hifi_equals_reference(built, hifi_reads, tmp_path, extra):
    _both(["--step", "1", "-t", "8", "-x", "ava-hifi"] + extra + [hifi_reads, hifi_reads], tmp_path)


def test_step1_ava_hifi_on_noisy_reads(built, golden_dir, tmp_path):
    """ONT reads under the HiFi preset: few 51-mers survive 15 % error -- a sparse-seed corner (and real homopolymer runs for the HPC queue)"""
    src = os.path.join(golden_dir, "reads_test.seed.2bit")
    _both(["--step", "1", "-t", "8", "-x", "ava-hifi", src, src], tmp_path, min_bytes=1000)


def test_step1_ava_hifi_cigar_equals_reference(built, golden_dir, tmp_path):
    """-c under the HiFi scoring (a=1 b=4 q=6,26 e=2,1).  On clean HiFi-like reads the reference itself aborts in mm_fix_cigar (align.c:125) with
    this preset, so the comparison runs on the noisy reads, where it completes."""
    src = os.path.join(golden_dir, "sub120.2bit")
    _both(["--step", "1", "-c", "-t", "16", "-x", "ava-hifi", src, src], tmp_path, min_bytes=500)


def test_step2_ava_hifi_equals_reference(built, hifi_reads, tmp_path):
    """cns_align as the pipeline runs it for HiFi reads (ref: lib/config_parser.py:60-64): --minide 0.1 --maxhan1 N"""
    _both(["-I", "6G", "--step", "2", "-t", "8", "-x", "ava-hifi", "--minide", "0.1", "--maxhan1", "1000", hifi_reads, hifi_reads], tmp_path, files=("o.ovl", "o.ovl.bl"), min_bytes=5000)


def test_paf_ava_hifi_equals_reference(built, hifi_reads, tmp_path):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref not present")
    src = hifi_reads
    a = subprocess.run([EXE, "-x", "ava-hifi", "-t", "8", src, src], capture_output=True, timeout=900)
    b = subprocess.run([REF, "-x", "ava-hifi", "-t", "8", src, src], capture_output=True, timeout=900)
    assert a.returncode == 0 and b.returncode == 0, a.stderr.decode()[-1000:]
    assert len(b.stdout) > 1000 and a.stdout == b.stdout
