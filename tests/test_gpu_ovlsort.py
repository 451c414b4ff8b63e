"""GPU `ovl_sort` (ndg_ovl_sort, bin/ovl_sort) against the unmodified reference binary: bytes of OUT and OUT.bl (run on the B200 box)."""
import os
import subprocess
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ovl_sort")
EXE = os.path.join(ROOT, "bin", "ovl_sort")


@pytest.fixture(scope="module")
def seeds(tmp_path_factory, golden_dir, built):
    import nextdenovo_b200 as nd
    d = tmp_path_factory.mktemp("gos")
    r = nd.read_2bit(os.path.join(golden_dir, "reads_test.seed.2bit"))
    idx = str(d / "seeds.idx")
    with open(idx, "w") as f:
        for i in range(r.n):
            f.write("%d\t%d\t%d\n" % (r.ids[i], 0, r.length[i]))
    return d, idx, r


def _ref(d, idx, name, files, opts):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/ovl_sort not present")
    fofn = str(d / (name + ".fofn"))
    open(fofn, "w").write("".join(f + "\n" for f in files))
    subprocess.run([REF, "-i", idx, "-t", "2", "-m", "1G"] + opts + ["-o", "ref_%s.ovl" % name, fofn], cwd=str(d), check=True, capture_output=True)
    blp = d / ("ref_%s.ovl.bl" % name)
    return fofn, open(d / ("ref_%s.ovl" % name), "rb").read(), (open(blp, "rb").read() if os.path.exists(blp) else None)


CASES = [("g1", ["G1_ava-ont.ovl"], []), ("g1H", ["G1_ava-ont.ovl"], ["-H"]), ("g3", ["G3_ava-pb.ovl"], []), ("g4cH", ["G4_ava-ont_c.ovl"], ["-H"]),
         ("g2k20", ["G2_ava-ont_k17w10.ovl"], ["-k", "20"]), ("g1l0", ["G1_ava-ont.ovl"], ["-l", "0"]), ("two", ["G1_ava-ont.ovl", "G3_ava-pb.ovl"], []),
         ("sub", ["S1_sub120_ava-ont.ovl", "S3_sub120_ava-pb.ovl", "S5_sub120_dual.ovl"], ["-l", "100", "-k", "10"])]


@pytest.mark.parametrize("name,files,opts", CASES, ids=[c[0] for c in CASES])
def test_cli_equals_reference(seeds, golden_dir, name, files, opts):
    d, idx, _ = seeds
    fofn, want, want_bl = _ref(d, idx, name, [os.path.join(golden_dir, f) for f in files], opts)
    out = str(d / ("gpu_%s.ovl" % name))
    r = subprocess.run([EXE, "-i", idx, "-t", "8"] + opts + ["-o", out, fofn], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(want) > 10000
    assert open(out, "rb").read() == want
    if want_bl is not None:
        assert open(out + ".bl", "rb").read() == want_bl
    else:
        assert not os.path.exists(out + ".bl")


def test_engine_api_and_properties(seeds, golden_dir, engine_mod):
    """through the C ABI (ctypes): same bytes as the executable's run; size-independent properties of the sorted stream"""
    import nextdenovo_b200 as nd
    d, idx, r = seeds
    data = open(os.path.join(golden_dir, "G1_ava-ont.ovl"), "rb").read()
    e = engine_mod.Engine(0)
    try:
        ovl, bl, n_in, n_out = e.ovl_sort([data], r.ids, r.length)
    finally:
        e.close()
    assert n_in == 130626 and n_out > 50000
    rec = nd.decode_ovl8(ovl)
    q = rec[:, 0].astype(np.int64)
    assert (np.diff(q) >= 0).all()                                   # sorted by query
    starts = np.r_[0, np.nonzero(np.diff(q))[0] + 1]
    assert (rec[starts, 0] == rec[starts, 4]).all() and (rec[starts, 2] == 0).all()          # every read opens with its own record
    lens = dict(zip(r.ids.tolist(), r.length.tolist()))
    assert all(int(rec[s, 3]) == lens[int(rec[s, 0])] - 1 for s in starts)
    inner = np.ones(len(q), bool); inner[starts] = False
    same = inner[1:] & inner[:-1] & (q[1:] == q[:-1])
    assert (rec[1:, 7][same] <= rec[:-1, 7][same]).all()             # match descending inside a read
    for line in bl.decode().splitlines():
        name, lab = line.split()
        assert lab in ("c", "k") and int(name) in lens


def test_rejects_missing_inputs(seeds, tmp_path):
    d, idx, _ = seeds
    fofn = str(tmp_path / "x.fofn"); open(fofn, "w").write("/nonexistent.ovl\n")
    assert subprocess.run([EXE, "-i", idx, "-o", str(tmp_path / "o.ovl"), fofn], capture_output=True).returncode == 1
    assert subprocess.run([EXE, "-o", str(tmp_path / "o.ovl"), fofn], capture_output=True).returncode == 1
