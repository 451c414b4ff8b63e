"""Host side of regime B (`-c`): nextdenovo_b200/csrc/nd_align.hpp decides which DP calls a chain needs and
stitches their results (replaces mm_align_skeleton & co, ref: minimap2/align.c).  Here it runs with a CPU
checker as the DP backend (the unmodified reference's ksw_extd2_sse from oracle/_ref) on chains from the
oracle, and the .ovl bytes are compared with what the reference writes for `--step 1 -c`.  CPU only; the
product's only DP backend is the sm_100a kernel (tests/test_gpu_parity.py covers that leg)."""
import os
import subprocess
import pytest
import nextdenovo_b200 as nd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libminimap2_ref.so")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("aln") / "aln_harness")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "libnd_oracle.so"])
    subprocess.check_call(["g++", "-O3", "-std=c++17", "-fopenmp", "-o", exe, os.path.join(ROOT, "tests", "aln_harness.cc"),
                           "-L" + os.path.join(ROOT, "oracle"), "-lnd_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-ldl"])
    return exe


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built (make -C oracle ref)")
def test_align_logic_sub30_matches_reference(harness, golden_dir, tmp_path):
    sub30 = str(tmp_path / "sub30.2bit")
    nd.write_2bit(sub30, nd.read_2bit(os.path.join(golden_dir, "sub120.2bit"), max_reads=30))
    out = str(tmp_path / "o.ovl")
    r = subprocess.run([harness, "ref", "ava-ont", "0", "0", sub30, sub30, out], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    assert open(out, "rb").read() == open(os.path.join(golden_dir, "S8_sub30_ava-ont_c.ovl"), "rb").read()
    assert "records=439" in r.stderr


def test_align_logic_tiny_with_oracle_dp(harness, golden_dir, tmp_path):
    """same logic with the oracle's scalar ksw restatement as the DP backend (8 reads: seconds)"""
    sub = str(tmp_path / "sub8.2bit")
    nd.write_2bit(sub, nd.read_2bit(os.path.join(golden_dir, "sub120.2bit"), max_reads=8))
    a, b = str(tmp_path / "a.ovl"), str(tmp_path / "b.ovl")
    r = subprocess.run([harness, "oracle", "ava-ont", "0", "0", sub, sub, a], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    if os.path.exists(REF_SO):
        r2 = subprocess.run([harness, "ref", "ava-ont", "0", "0", sub, sub, b], capture_output=True, text=True, cwd=ROOT)
        assert r2.returncode == 0, r2.stderr
        assert open(a, "rb").read() == open(b, "rb").read()


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built (make -C oracle ref)")
def test_align_logic_with_ambiguous_bases_matches_reference_run(harness, tmp_path):
    """40 reads with N runs and lower case in FASTA: a live run of the reference with -c and the host logic (+ reference ksw as the
    DP checker, oracle chains) write the same bytes"""
    ref = os.path.join(ROOT, "oracle", "_ref", "minimap2-nd")
    fa = str(tmp_path / "r.fa")
    subprocess.check_call(["python", os.path.join(ROOT, "tools", "mkfa_n.py"), fa], cwd=ROOT)
    want, got = str(tmp_path / "ref.ovl"), str(tmp_path / "mine.ovl")
    subprocess.run([ref, "--step", "1", "-c", "-t", "8", "-x", "ava-ont", fa, fa, "-o", want], check=True, capture_output=True)
    r = subprocess.run([harness, "ref", "ava-ont", "0", "0", fa, fa, got], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    assert open(got, "rb").read() == open(want, "rb").read()
