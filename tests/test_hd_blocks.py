"""The per-lane building blocks the CUDA kernels execute (nextdenovo_b200/csrc/nd_core.cuh) are
__host__ __device__: compile them for the host and check them stage by stage against the oracle
(chunked sketch with mid-read warm-up, exact flag-sort emulation, per-segment chain DP + backtrack).
CPU only; the product never runs these on the CPU."""
import os
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("hd") / "hd_harness")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-o", exe, os.path.join(ROOT, "tests", "hd_harness.cc"),
                           os.path.join(ROOT, "oracle", "nd_oracle.c"), "-I", ROOT])
    return exe


@pytest.mark.parametrize("preset,n,chunk", [("ava-ont", 30, 61), ("ava-pb", 30, 47), ("k17w10", 40, 512), ("k16w7", 25, 100)])
def test_hd_blocks_match_oracle(harness, golden_dir, preset, n, chunk):
    out = subprocess.run([harness, os.path.join(golden_dir, "sub120.2bit"), preset, str(n), str(chunk)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "bad_sketch=0 bad_sort=0 bad_hits=0" in out.stdout


def test_ambiguous_bases_chunked_sketch(harness):
    """1500 random sequences with N runs x 6 (w,k,hpc) settings: chunked state machine with the 1-bit mask == oracle"""
    out = subprocess.run([harness, "selftest_n"], capture_output=True, text=True)
    assert out.returncode == 0 and "bad=0" in out.stdout, out.stdout + out.stderr


def test_ordered_index_table(harness):
    """placement rule of k_index_place + forward scan of nd_idx_lookup replayed on the host: hits, misses, slices, pad"""
    out = subprocess.run([harness, "selftest_idx"], capture_output=True, text=True)
    assert out.returncode == 0 and "idx bad=0" in out.stdout, out.stdout + out.stderr
