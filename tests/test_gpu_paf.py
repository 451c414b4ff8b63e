"""PAF text mode of the drop-in executable against the unmodified reference binary, byte for byte (run on the B200 box: -m gpu).
PAF is the only place the reference shows chain statistics (cm/s1), the divergence estimate of mm_est_err (dv:f, ref:
minimap2/esterr.c:30-64), rep_len (rl:i, ref: minimap2/map.c:106-119), mapq (ref: minimap2/hit.c:446-491) and, with -c, the ksw2
scores (AS:i = dp_score, ms:i = dp_max) and the CIGAR (cg:Z) -- so this is where BASELINE.json's "within +-1 on ksw2 alignment
scores" is checked (they are equal)."""
import os
import re
import subprocess
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "bin", "minimap2-nd")
REF = os.path.join(ROOT, "oracle", "_ref", "minimap2-nd")


def _run(exe, args, cwd):
    r = subprocess.run([exe] + args, capture_output=True, cwd=cwd, timeout=1800)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    return r.stdout


def _pair(args, tmp_path):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref not present")
    got = _run(EXE, args, str(tmp_path))
    want = _run(REF, args, str(tmp_path))
    assert len(want) > 1000
    if got != want:
        g, w = got.split(b"\n"), want.split(b"\n")
        n = min(len(g), len(w))
        bad = next((i for i in range(n) if g[i] != w[i]), n)
        raise AssertionError("PAF differs: %d vs %d lines; first difference at line %d:\n got %r\nwant %r" % (len(g), len(w), bad, g[bad][:400] if bad < len(g) else None, w[bad][:400] if bad < len(w) else None))
    return got


@pytest.mark.parametrize("args", [
    ["-x", "ava-ont"],                                   # plain mapping: PAF (no --step)
    ["--step", "1", "-x", "ava-ont"],                    # --step 1 without -o: outraw stays on (ref: minimap2/main.c:191,274)
    ["-x", "ava-pb"],                                    # HPC minimizers: spans differ, avg_k is a real average
    ["-x", "ava-ont", "-k", "17", "-w", "10", "-K", "2M"],
], ids=["ava-ont", "step1-outraw", "ava-pb-hpc", "k17w10-batches"])
def test_paf_chain_mode_equals_reference(built, golden_dir, tmp_path, args):
    src = os.path.join(golden_dir, "sub120.2bit")
    out = _pair(args + ["-t", "8", src, src], tmp_path)
    line = out.split(b"\n")[0].split(b"\t")
    assert line[12] == b"tp:A:S" and line[15].startswith(b"dv:f:") and line[16].startswith(b"rl:i:")


def test_paf_full_set_divergence_and_rep_len(built, golden_dir, tmp_path):
    """all 997 reads (130k lines): dv:f / rl:i / cm:i / s1:i of every hit"""
    src = os.path.join(golden_dir, "reads_test.seed.2bit")
    _pair(["-x", "ava-ont", "-t", "16", src, src], tmp_path)


def test_paf_outraw_into_file(built, golden_dir, tmp_path):
    src = os.path.join(golden_dir, "sub120.2bit")
    for exe, name in ((EXE, "g.paf"), (REF, "r.paf")):
        if not os.path.exists(exe):
            pytest.skip("oracle/_ref not present")
        r = subprocess.run([exe, "--step", "1", "-x", "ava-ont", "-t", "8", src, src, "-o", str(tmp_path / name), "--outraw"], capture_output=True)
        assert r.returncode == 0, r.stderr.decode()[-1000:]
    assert open(tmp_path / "g.paf", "rb").read() == open(tmp_path / "r.paf", "rb").read()


@pytest.mark.parametrize("preset", ["ava-ont", "ava-pb"])
def test_paf_cigar_mode_scores_and_cigars(built, golden_dir, tmp_path, preset):
    """-c: NM/ms/AS/nn, tp (the split-off halves are 'P' with a real mapq), de:f, zd:i, cg:Z -- and the +-1 score clause of the north star"""
    src = os.path.join(golden_dir, "sub120.2bit")
    args = ["-c", "-x", preset, "-t", "16", src, src]
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref not present")
    got = _run(EXE, args, str(tmp_path)).split(b"\n")
    want = _run(REF, args, str(tmp_path)).split(b"\n")
    assert len(got) == len(want) and len(want) > 5000
    pat = re.compile(rb"\tms:i:(-?\d+)\tAS:i:(-?\d+)")
    worst = 0
    for g, w in zip(got, want):
        if not w:
            continue
        assert g.split(b"\t")[:12] == w.split(b"\t")[:12]                # coordinates, mlen, blen, mapq: bit-exact
        mg, mw = pat.search(g), pat.search(w)
        worst = max(worst, abs(int(mg.group(1)) - int(mw.group(1))), abs(int(mg.group(2)) - int(mw.group(2))))
    assert worst <= 1                                                    # BASELINE.json north_star tolerance on ksw2 scores
    assert got == want                                                   # and in fact every byte, CIGARs included
    assert any(b"tp:A:P" in x and b"zd:i:" in x for x in want)


def test_paf_rejects_what_is_not_built(built, golden_dir, tmp_path):
    src = os.path.join(golden_dir, "sub120.2bit")
    assert subprocess.run([EXE, "--step", "2", "-x", "ava-ont", src, src], capture_output=True).returncode == 1      # step-2 text mode
    assert subprocess.run([EXE, "-x", "map-ont", src, src], capture_output=True).returncode == 1
    assert subprocess.run([EXE, src, src], capture_output=True).returncode == 1                                        # not all-chain
