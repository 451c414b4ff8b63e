"""Index and multi-GPU parity (run on the B200 box: pytest -m gpu), through the C ABI.
  * ndg_index_get == the reference's mm_idx_get semantics (count + ascending positions), checked against a dictionary built from the
    oracle's sketch of every target;
  * the N-way sharded build (ndg_index_shard_*), exercised in one process with two contexts on one GPU: mapping against the index
    built in two halves gives the golden bytes on BOTH contexts, and the index statistics equal the single-context build;
  * the same over NCCL with one process per GPU (torch.distributed) when the box has >= 2 GPUs;
  * advisor findings: a borrower of a shared index that calls set_options() keeps mapping correctly (mid_occ is re-derived);
  * the host emit path (ndg_map_batch + ndg_ovl_emit_step1) gives the same bytes as the device output stage."""
import os
import subprocess
import sys
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _opts(preset="ava-ont", k=None, w=None):
    import nextdenovo_b200 as nd
    io, mo = nd.IdxOpt(), nd.MapOpt()
    nd.set_opt(preset, io, mo)
    if k: io.k = k
    if w: io.w = w
    mo.step, mo.minlen = 1, 500
    return io, mo


def _seq(reads, r):
    from nextdenovo_b200.formats import unpack_bases
    w0, w1 = int(reads.word_off[r]), int(reads.word_off[r + 1])
    return bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[unpack_bases(reads.words[w0:w1], int(reads.length[r]))])


def test_index_get_matches_sketch_dictionary(engine_mod, reads_sub):
    """ref: minimap2/index.c:81-98 (mm_idx_get): n and the position list (ascending, index.c:230) of a minimizer"""
    from oracle import oracle
    io, mo = _opts()
    e = engine_mod.Engine(0)
    try:
        e.set_options(io, mo)
        sub = reads_sub.slice(0, 40)
        e.build_index(sub)
        d = {}
        for r in range(sub.n):
            for x, y in oracle.sketch(_seq(sub, r), io.w, io.k, rid=r):
                d.setdefault(int(x) >> 8, []).append(int(y))
        keys = sorted(d)
        rng = np.random.default_rng(7)
        pick = [keys[i] for i in rng.integers(0, len(keys), 300)] + [k for k in keys if len(d[k]) > 3][:50]
        for k in pick:
            pos, n = e.index_get(k)
            assert n == len(d[k]) and pos.tolist() == sorted(d[k]), "minimizer %x" % k
        for k in rng.integers(0, 1 << 30, 200):
            if int(k) not in d:
                assert e.index_get(int(k))[1] == 0
        assert e.stat["n_distinct"] == len(keys) and e.stat["n_minimizers"] == sum(len(v) for v in d.values())
    finally:
        e.close()


def test_host_emit_path_equals_device_emit(engine_mod, reads_sub, golden_dir):
    io, mo = _opts()
    e = engine_mod.Engine(0)
    try:
        ovl, st = engine_mod.run_step1(e, reads_sub, reads_sub, io, mo, device_emit=False)
        assert ovl == open(os.path.join(golden_dir, "S1_sub120_ava-ont.ovl"), "rb").read()
    finally:
        e.close()


def test_borrower_set_options_after_share_keeps_mid_occ(engine_mod, reads_sub, golden_dir):
    """advisor r1: ndg_set_opt() after ndg_index_share() used to zero mid_occ and drop every seed silently"""
    io, mo = _opts()
    a, b = engine_mod.Engine(0), engine_mod.Engine(0)
    try:
        a.set_options(io, mo); b.set_options(io, mo)
        a.build_index(reads_sub)
        b.share_index_of(a)
        b.set_options(io, mo)                               # after the share
        st = np.zeros(2, dtype=np.uint32)
        assert b.map_batch_ovl(reads_sub, st)[0] == open(os.path.join(golden_dir, "S1_sub120_ava-ont.ovl"), "rb").read()
        with pytest.raises(engine_mod.NdgError):             # a borrower must not build over the owner's buffers
            b._ck(b.L.ndg_index_finalize(b.h, None))
    finally:
        b.close(); a.close()


def test_mapping_without_cutoff_is_an_error(engine_mod, reads_sub):
    e = engine_mod.Engine(0)
    try:
        with pytest.raises(engine_mod.NdgError):
            e.map_batch(reads_sub)                           # no index: NDG_ERR_STATE, never an empty result
    finally:
        e.close()


@pytest.mark.parametrize("preset,gold,cut", [("ava-ont", "S1_sub120_ava-ont.ovl", 60), ("ava-pb", "S3_sub120_ava-pb.ovl", 35)])
def test_sharded_build_two_contexts_one_gpu(engine_mod, reads_sub, golden_dir, preset, gold, cut):
    """ndg_index_shard_split/bucket/recv_reserve/sort/place/table_slice/pos_reserve/finish with N = 2: each context sketches one part
    of the targets and builds one key range of the table; both end with the full index (same statistics as a one-context build) and
    map to the golden bytes."""
    import torch
    io, mo = _opts(preset)
    a, b, c = engine_mod.Engine(0), engine_mod.Engine(0), engine_mod.Engine(0)
    try:
        for e in (a, b, c):
            e.set_options(io, mo)
        want_stat = c.build_index(reads_sub)
        stats = engine_mod.build_index_sharded_local([a, b], [reads_sub.slice(0, cut), reads_sub.slice(cut, reads_sub.n)], torch)
        for s in stats:
            for k in ("n_seq", "n_minimizers", "n_distinct", "n_singletons", "mid_occ"):
                assert s[k] == want_stat[k], k
        want = open(os.path.join(golden_dir, gold), "rb").read()
        for e in (a, b):
            st = np.zeros(2, dtype=np.uint32)
            assert e.map_batch_ovl(reads_sub, st)[0] == want
    finally:
        for e in (a, b, c):
            e.close()


def test_sharded_build_nccl_two_gpus(built, golden_dir, tmp_path):
    """one process per GPU, torch.distributed over NCCL (engine.build_index_sharded): every rank maps all queries against the index it
    assembled and must produce the single-GPU golden bytes; bench.py --gpus N uses exactly this path"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    script = tmp_path / "w.py"
    script.write_text('''
import os, sys, hashlib
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import nextdenovo_b200 as nd
from nextdenovo_b200 import engine
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
reads = nd.read_2bit(os.path.join(%r, "reads_test.seed.2bit"))
io, mo = nd.IdxOpt(), nd.MapOpt(); nd.set_opt("ava-ont", io, mo); mo.step, mo.minlen = 1, 500
e = engine.Engine(rank); e.set_options(io, mo)
b = engine.shard_bounds(reads.n, world)
st = engine.build_index_sharded(e, reads.slice(b[rank], b[rank + 1]), rank, world, dist, torch)
state = np.zeros(2, dtype=np.uint32)
ovl = e.map_batch_ovl(reads, state)[0]
print("RESULT", rank, hashlib.md5(ovl).hexdigest(), st["n_distinct"], st["mid_occ"], flush=True)
e.close(); dist.destroy_process_group()
''' % (ROOT, golden_dir))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(29600 + os.getpid() % 300), str(script)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = [l.split() for l in r.stdout.splitlines() if l.startswith("RESULT")]
    assert len(res) == 2
    for x in res:
        assert x[2] == "9fce14e8841db2e46012c58ad8f0059c" and x[3] == "6273819" and x[4] == "74", x     # G1 of SURVEY 8(c)


def test_cli_two_gpus_same_bytes(built, golden_dir, tmp_path):
    """bin/minimap2-nd --gpus 2: index part built sharded over both GPUs (peer copies), query batches pulled by 2 x 2 contexts, output
    merged in input order with re-based batch heads: the bytes of the single-GPU run / the reference (G1, and a 3-part --dual run)"""
    import hashlib
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    exe = os.path.join(ROOT, "bin", "minimap2-nd")
    src = os.path.join(golden_dir, "reads_test.seed.2bit")
    out = str(tmp_path / "o.ovl")
    r = subprocess.run([exe, "--step", "1", "--gpus", "2", "-K", "5M", "-x", "ava-ont", src, src, "-o", out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert hashlib.md5(open(out, "rb").read()).hexdigest() == "9fce14e8841db2e46012c58ad8f0059c"
    one = str(tmp_path / "one.ovl")
    args = ["--step", "1", "--dual=yes", "-I", "12M", "-K", "3M", "-x", "ava-ont", "-k", "17", "-w", "10", src, src]
    assert subprocess.run([exe] + args + ["-o", one], capture_output=True).returncode == 0
    assert subprocess.run([exe, "--gpus", "2"] + args + ["-o", out], capture_output=True).returncode == 0
    assert open(one, "rb").read() == open(out, "rb").read() and os.path.getsize(one) > 100000


def test_cli_context_pool_small_batches(built, golden_dir, tmp_path):
    """one GPU, two contexts sharing the index, many small query batches (-K 1M): the stitched stream equals the golden bytes"""
    exe = os.path.join(ROOT, "bin", "minimap2-nd")
    src = os.path.join(golden_dir, "sub120.2bit")
    out = str(tmp_path / "o.ovl")
    r = subprocess.run([exe, "--step", "1", "-K", "1M", "-x", "ava-ont", src, src, "-o", out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert open(out, "rb").read() == open(os.path.join(golden_dir, "S1_sub120_ava-ont.ovl"), "rb").read()
    assert r.stderr.count("mapped") >= 3


def test_prefetched_batches_equal_plain_calls(engine_mod, reads_sub, golden_dir):
    """ndg_map_prefetch: batches uploaded ahead on the copy stream (immediately, or deferred until the previous one is taken) map to the
    same bytes as plain calls; a prefetch that does not match the next call, or exceeds a sub-batch, is simply not used"""
    io, mo = _opts()
    e = engine_mod.Engine(0)
    try:
        e.set_options(io, mo)
        e.build_index(reads_sub)
        parts = [engine_mod.PinnedReads(reads_sub.slice(a, b)) for a, b in ((0, 30), (30, 75), (75, 76), (76, reads_sub.n))]
        want = open(os.path.join(golden_dir, "S1_sub120_ava-ont.ovl"), "rb").read()
        st = np.zeros(2, dtype=np.uint32); got = b""
        e.map_prefetch(parts[0].reads)
        for i, p in enumerate(parts):
            if i + 1 < len(parts):
                e.map_prefetch(parts[i + 1].reads)               # deferred: the second buffer still holds batch i
            got += e.map_batch_ovl(p.reads, st)[0]
        assert got == want
        st[:] = 0; got = b""
        e.map_prefetch(parts[2].reads)                           # wrong batch prefetched: ignored
        for p in parts:
            got += e.map_batch_ovl(p.reads, st)[0]
        assert got == want
        e.set_limits(1 << 20, 0)                                 # 1 Mbp sub-batches: the prefetch declines, the call uploads piecewise
        st[:] = 0
        e.map_prefetch(parts[3].reads)
        assert e.map_batch_ovl(parts[0].reads, st)[0] + e.map_batch_ovl(parts[1].reads, st)[0] + e.map_batch_ovl(parts[2].reads, st)[0] + e.map_batch_ovl(parts[3].reads, st)[0] == want
    finally:
        e.close()
