"""N > 1 host logic on CPU (gloo, world_size 2): rank r sketches its contiguous slice of the targets
with global read ids, the variable-size all-gather concatenates them in rank order, and the result is
the single-process sketch of all targets -- so every rank builds the identical index.  The GPU path
(engine.build_index_sharded) does exactly this with NCCL and device buffers."""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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
    from nextdenovo_b200.engine import shard_bounds
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    reads = nd.read_2bit(os.path.join(ROOT, "tests", "golden", "sub120.2bit"), max_reads=24)
    b = shard_bounds(reads.n, world)
    mine = [oracle.sketch(s, 5, 15, rid=b[rank] + i) for i, s in enumerate(_seqs(reads, b[rank], b[rank + 1]))]
    mine = np.concatenate(mine).astype(np.int64)
    cnt = torch.tensor([mine.shape[0]])
    cnts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    mx = max(int(c) for c in cnts)
    send = torch.zeros(mx * 2, dtype=torch.int64); send[:mine.size] = torch.from_numpy(mine.reshape(-1))
    recv = [torch.zeros(mx * 2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(recv, send)
    full = np.concatenate([recv[r].numpy()[:int(cnts[r]) * 2] for r in range(world)]).reshape(-1, 2).astype(np.uint64)
    single = np.concatenate([oracle.sketch(s, 5, 15, rid=i) for i, s in enumerate(_seqs(reads, 0, reads.n))])
    q.put((rank, bool(full.shape == single.shape and (full == single).all()), int(full.shape[0])))
    dist.destroy_process_group()


def test_sharded_sketch_allgather_equals_single():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps: p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps: p.join(60)
    assert [r[1] for r in res] == [True, True] and res[0][2] == res[1][2] > 1000
