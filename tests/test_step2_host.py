"""Host side of `--step 2` (nextdenovo_b200/csrc/nd_step2.hpp: candidate marking, re-mapping plan and update rules, 10-field writer with
its order-dependent contained/dovetail bookkeeping, `.bl` text in klib bucket order) with the oracle as the mapping backend, against the
bytes the unmodified reference writes (tests/golden/S9_*, minted by make_golden.py).  CPU only."""
import os
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("s2") / "step2_harness")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "libnd_oracle.so"])
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-o", exe, os.path.join(ROOT, "tests", "step2_harness.cc"),
                           "-L" + os.path.join(ROOT, "oracle"), "-lnd_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    return exe


def test_step2_logic_sub120_matches_reference(harness, golden_dir, tmp_path):
    src = os.path.join(golden_dir, "sub120.2bit")
    out = str(tmp_path / "o.ovl")
    r = subprocess.run([harness, "ava-ont", "17", "17", "2000", "5000", src, src, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out, "rb").read() == open(os.path.join(golden_dir, "S9_sub120_step2.ovl"), "rb").read()
    assert open(out + ".bl", "rb").read() == open(os.path.join(golden_dir, "S9_sub120_step2.ovl.bl"), "rb").read()
    assert "records=148" in r.stderr


def test_step2_logic_with_ambiguous_bases_matches_reference_run(harness, tmp_path):
    """40 reads with N runs and lower case in FASTA: the reference binary (when oracle/_ref is present) and the host logic + oracle
    mapping agree on .ovl and .bl; pins the N -> A reading of re-mapped targets (ref: minimap2/map.c:1041-1045)"""
    ref = os.path.join(ROOT, "oracle", "_ref", "minimap2-nd")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not present")
    fa = str(tmp_path / "r.fa")
    subprocess.check_call(["python", os.path.join(ROOT, "tools", "mkfa_n.py"), fa], cwd=ROOT)
    want, got = str(tmp_path / "ref.ovl"), str(tmp_path / "mine.ovl")
    subprocess.run([ref, "-I", "6G", "--step", "2", "-t", "4", "-x", "ava-ont", "-k", "17", "-w", "17", "--minlen", "2000", "--maxhan1", "5000", fa, fa, "-o", want],
                   check=True, capture_output=True)
    r = subprocess.run([harness, "ava-ont", "17", "17", "2000", "5000", fa, fa, got], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(got, "rb").read() == open(want, "rb").read()
    assert open(got + ".bl", "rb").read() == open(want + ".bl", "rb").read()


@pytest.mark.skipif(not os.path.exists("/root/reference/util/khash.h"), reason="reference tree not present")
def test_ktable_bucket_layout_equals_klib(tmp_path):
    """the table behind the .bl line order against the reference's own khash.h: same buckets after the same insertions
    (66 cases: sequential ids, colliding low bits, random keys, repeats; 1 to 40000 keys, i.e. through 14 growth steps)"""
    exe = str(tmp_path / "ktable_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I/root/reference/util", "-o", exe, os.path.join(ROOT, "tests", "ktable_check.cc")])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "bad=0" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("preset,k,w,minlen,maxhan1,extra,env", [
    ("ava-ont", 15, 10, 500, 3000, [], {}),
    ("ava-ont", 15, 5, 500, 5000, ["--outctn", "--minide", "0.02"], {"S2_OUTCTN": "1", "S2_MINIDE": "20"}),
    ("ava-pb", 19, 10, 1000, 5000, [], {}),
    ("ava-ont", 15, 5, 300, 5000, ["--cn", "5", "--minide", "0.01", "--minmatch", "30"], {"S2_CN": "5", "S2_MINIDE": "10", "S2_MINMATCH": "30"}),
], ids=["k15w10", "outctn", "hpc", "cn5"])
def test_step2_logic_option_sweep_against_reference_run(harness, golden_dir, tmp_path, preset, k, w, minlen, maxhan1, extra, env):
    ref = os.path.join(ROOT, "oracle", "_ref", "minimap2-nd")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not present")
    src = os.path.join(golden_dir, "sub120.2bit")
    want, got = str(tmp_path / "ref.ovl"), str(tmp_path / "mine.ovl")
    subprocess.run([ref, "-I", "6G", "--step", "2", "-t", "4", "-x", preset, "-k", str(k), "-w", str(w), "--minlen", str(minlen), "--maxhan1", str(maxhan1)] + extra +
                   [src, src, "-o", want], check=True, capture_output=True)
    r = subprocess.run([harness, preset, str(k), str(w), str(minlen), str(maxhan1), src, src, got], capture_output=True, text=True, env=dict(os.environ, **env))
    assert r.returncode == 0, r.stderr
    assert open(got, "rb").read() == open(want, "rb").read()
    assert open(got + ".bl", "rb").read() == open(want + ".bl", "rb").read()
