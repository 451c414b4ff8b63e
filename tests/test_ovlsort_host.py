"""`ovl_sort`: the per-read stage of the GPU implementation (nextdenovo_b200/csrc/nd_ovlsort.cuh, __host__ __device__)
compiled for the host inside tests/ovlsort_harness.cc and compared, bytes of OUT and OUT.bl, with the unmodified reference binary
(oracle/_ref/ovl_sort, ref: util/ovl_sort.c) on the overlap files of the reference's test data.  CPU only; the product path runs the
same code on the GPU (tests/test_gpu_ovlsort.py)."""
import os
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ovl_sort")


@pytest.fixture(scope="module")
def env(tmp_path_factory, golden_dir):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/ovl_sort not present")
    import nextdenovo_b200 as nd
    d = tmp_path_factory.mktemp("os")
    exe = str(d / "ovlsort_harness")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-o", exe, os.path.join(ROOT, "tests", "ovlsort_harness.cc"), "-I", ROOT])
    r = nd.read_2bit(os.path.join(golden_dir, "reads_test.seed.2bit"))
    idx = str(d / "seeds.idx")
    with open(idx, "w") as f:
        for i in range(r.n):
            f.write("%d\t%d\t%d\n" % (r.ids[i], 0, r.length[i]))
    return d, exe, idx


def run_reference(d, idx, name, files, opts):
    fofn = str(d / (name + ".fofn"))
    open(fofn, "w").write("".join(f + "\n" for f in files))
    # -t 2: one reader thread, so the in-memory order is the file order (with more readers the reference's tie order depends on a race)
    subprocess.run([REF, "-i", idx, "-t", "2", "-m", "1G"] + opts + ["-o", "ref_%s.ovl" % name, fofn], cwd=str(d), check=True, capture_output=True)
    out = open(d / ("ref_%s.ovl" % name), "rb").read()
    blp = d / ("ref_%s.ovl.bl" % name)
    return out, open(blp, "rb").read() if os.path.exists(blp) else None


CASES = [("g1", ["G1_ava-ont.ovl"], []), ("g1H", ["G1_ava-ont.ovl"], ["-H"]), ("g3", ["G3_ava-pb.ovl"], []), ("g4c", ["G4_ava-ont_c.ovl"], []),
         ("g4cH", ["G4_ava-ont_c.ovl"], ["-H"]), ("g2k20", ["G2_ava-ont_k17w10.ovl"], ["-k", "20"]), ("g1l0", ["G1_ava-ont.ovl"], ["-l", "0"]),
         ("g1l100k10", ["G1_ava-ont.ovl"], ["-l", "100", "-k", "10"]), ("two", ["G1_ava-ont.ovl", "G3_ava-pb.ovl"], [])]


@pytest.mark.parametrize("name,files,opts", CASES, ids=[c[0] for c in CASES])
def test_per_read_stage_equals_reference(env, golden_dir, name, files, opts):
    d, exe, idx = env
    paths = [os.path.join(golden_dir, f) for f in files]
    want, want_bl = run_reference(d, idx, name, paths, opts)
    mine = str(d / ("mine_%s.ovl" % name))
    subprocess.run([exe, idx, mine] + opts + paths, check=True, capture_output=True)
    assert len(want) > 100000
    assert open(mine, "rb").read() == want
    if want_bl is not None:
        assert open(mine + ".bl", "rb").read() == want_bl
    else:
        assert not os.path.exists(mine + ".bl")
