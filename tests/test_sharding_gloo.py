"""N > 1 host logic of the sharded index build on CPU (gloo, world_size 2).  Product code under test: the key-range rule
(ndg_index_key_slice = nd_idx_slice of nd_core.cuh, the function the CUDA kernels partition with) and the all-to-all plan
(engine.shard_exchange_plan); the minimizers themselves come from the oracle (test infrastructure).  Each rank sketches its reads
with global read ids, buckets them by key range, exchanges the buckets with all_to_all_single in rank order and sorts what it
received stably: the concatenation over ranks must equal the stable sort of the single-process sketch of all reads (ascending
position inside a key, ref: minimap2/index.c:230) -- which is what engine.build_index_sharded does with NCCL and device buffers."""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K, W = 15, 5


def _mix(key):
    """nd_idx_mix (nd_core.cuh): odd multiplier mod 2^(2K)"""
    return (key * np.uint64(0x9E3779B97F4A7C15)) & np.uint64((1 << (2 * K)) - 1)


def _seqs(reads, r0, r1):
    from nextdenovo_b200.formats import unpack_bases
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = []
    for r in range(r0, r1):
        w0, w1 = int(reads.word_off[r]), int(reads.word_off[r + 1])
        out.append(bytes(lut[unpack_bases(reads.words[w0:w1], int(reads.length[r]))]))
    return out


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    import nextdenovo_b200 as nd
    from nextdenovo_b200.engine import shard_bounds, shard_exchange_plan, load_library
    from oracle import oracle
    L = load_library()
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    reads = nd.read_2bit(os.path.join(ROOT, "tests", "golden", "sub120.2bit"), max_reads=24)
    b = shard_bounds(reads.n, world)
    mine = np.concatenate([oracle.sketch(s, W, K, rid=b[rank] + i) for i, s in enumerate(_seqs(reads, b[rank], b[rank + 1]))])
    key, val = _mix(mine[:, 0] >> np.uint64(8)), mine[:, 1]    # the table is ordered by the mixed key (nd_idx_mix)
    order = np.argsort(key, kind="stable")
    key, val = key[order], val[order]
    raw = (mine[:, 0] >> np.uint64(8))[order]
    sl = np.array([L.ndg_index_key_slice(int(k), K, world) for k in raw], dtype=np.int64)
    assert (np.diff(sl) >= 0).all()                     # slices cut the sorted keys into contiguous ranges
    counts = np.bincount(sl, minlength=world).astype(np.int64)
    cm = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(cm, torch.from_numpy(counts))
    in_split, out_split, total = shard_exchange_plan(np.stack([c.numpy() for c in cm]), rank)
    got = []
    for arr in (key, val):
        out = torch.empty(total, dtype=torch.int64)
        dist.all_to_all_single(out, torch.from_numpy(arr.astype(np.int64)), output_split_sizes=out_split, input_split_sizes=in_split)
        got.append(out.numpy().astype(np.uint64))
    o2 = np.argsort(got[0], kind="stable")
    rk, rv = got[0][o2], got[1][o2]
    single = np.concatenate([oracle.sketch(s, W, K, rid=i) for i, s in enumerate(_seqs(reads, 0, reads.n))])
    sraw = single[:, 0] >> np.uint64(8)
    sk, sv = _mix(sraw), single[:, 1]
    so = np.argsort(sk, kind="stable")
    sk, sv, sraw = sk[so], sv[so], sraw[so]
    ssl = np.array([L.ndg_index_key_slice(int(k), K, world) for k in sraw])
    want_k, want_v = sk[ssl == rank], sv[ssl == rank]
    ok = rk.shape == want_k.shape and bool((rk == want_k).all() and (rv == want_v).all())
    q.put((rank, ok, int(rk.shape[0]), int(sk.shape[0])))
    dist.destroy_process_group()


def test_sharded_build_host_logic_equals_single():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps: p.start()
    res = sorted(q.get(timeout=180) for _ in ps)
    for p in ps: p.join(60)
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] + res[1][2] == res[0][3] > 1000 and min(res[0][2], res[1][2]) > 0.3 * res[0][3]    # both ranks own a real share


def test_exchange_plan():
    from nextdenovo_b200.engine import shard_exchange_plan
    cm = [[1, 2, 3], [4, 5, 6], [7, 8, 9]]
    assert shard_exchange_plan(cm, 1) == ([4, 5, 6], [2, 5, 8], 15)
