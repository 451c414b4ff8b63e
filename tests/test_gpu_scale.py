"""Parity at BASELINE.json configs[1] scale (100k synthetic 20 kb ONT reads, -k 17 -w 10): the GPU drop-in
executable maps the full read set against itself; the unmodified reference maps the first queries
against the same full index on the host cores; the reference's bytes must be a prefix of ours
(records come out in query order and the delta state starts at zero in both).  Also size-independent
properties of the full output.  Needs oracle/_ref (travels with the repo) and ~3 minutes."""
import os
import subprocess
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n_reads,n_check", [(int(os.environ.get("ND_SCALE_READS", "100000")), 600)])
def test_full_scale_prefix_equals_reference(built, tmp_path, n_reads, n_check):
    import nextdenovo_b200 as nd
    from nextdenovo_b200 import synth
    ref = os.path.join(ROOT, "oracle", "_ref", "minimap2-nd")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not present")
    g = synth.genome(n_reads * 20000 // 40, 20240923)
    reads = synth.reads(g, n_reads, 20000, seed=20240923)
    full = str(tmp_path / "full.2bit"); head = str(tmp_path / "head.2bit")
    nd.write_2bit(full, reads)
    nd.write_2bit(head, reads.slice(0, n_check))
    gpu_out = str(tmp_path / "gpu.ovl"); ref_out = str(tmp_path / "ref.ovl")
    exe = os.path.join(ROOT, "bin", "minimap2-nd")
    r = subprocess.run([exe, "--step", "1", "-t", "8", "-I", "100G", "-x", "ava-ont", "-k", "17", "-w", "10", full, full, "-o", gpu_out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    threads = str(os.cpu_count() or 8)
    r2 = subprocess.run([ref, "--step", "1", "-t", threads, "-I", "100G", "-x", "ava-ont", "-k", "17", "-w", "10", full, head, "-o", ref_out],
                        capture_output=True, text=True)
    assert r2.returncode == 0, r2.stderr[-2000:]
    want = open(ref_out, "rb").read()
    got = open(gpu_out, "rb").read()
    assert len(want) > 1000
    assert got[:len(want)] == want, "first %d queries differ from the reference at full index size" % n_check
    # size-independent properties of the whole job
    rec = nd.decode_ovl8(got)
    q, t = rec[:, 0].astype(np.int64), rec[:, 4].astype(np.int64)
    assert (np.diff(q) >= 0).all()                                   # query order
    assert (rec[:, 3] - rec[:, 2] >= 500).all()                      # --minlen
    assert (rec[:, 3] <= 20000).all() and (rec[:, 6] <= 20000).all() and (rec[:, 2] < rec[:, 3]).all() and (rec[:, 5] < rec[:, 6]).all()
    assert (q != t).all()
    # NO_DUAL: a pair is reported once, from the query whose name sorts first (strcmp order of the decimal names)
    names_q = np.array([str(x) for x in q[:20000]]); names_t = np.array([str(x) for x in t[:20000]])
    assert (names_q < names_t).all()
    pairs = set(zip(q.tolist(), t.tolist(), rec[:, 1].tolist(), rec[:, 2].tolist()))
    assert len(pairs) == rec.shape[0]
