"""The exactness argument of the sort's group filter (nextdenovo_b200/csrc/nd_warp.cuh: k_group_filter + k_sort_fast), checked on the
CPU against the oracle's literal mm_chain_dp / mm_gen_regs (oracle/nd_oracle.c; ref: minimap2/chain.c:22-162, hit.c:52-88).

The product drops, before the sort, every anchor whose (strand, target) group -- x >> 32 -- cannot reach min_cnt / min_chain_score,
sorts the kept anchors by any correct sort when their x are distinct, and writes placeholders behind them (x above every real
target, more than max_gap apart).  Claim: the chains and regions are those of the reference order (unstable radix sort of ALL
anchors).  Here both orders go through the same oracle chain code on real reads (tests/golden/sub120.2bit) and must give the same
chains (u), the same compacted anchors and the same regions.  The placeholders keep the dropped anchors' y so that the oracle's
average span (chain.c:41-42, a sum over all anchors of the read) is the same number on both sides, as it is in the product, which
computes it before the sort.  CPU only; the GPU kernels themselves are covered by the -m gpu parity suite."""
import ctypes as C
import os
import numpy as np
import pytest

from oracle import oracle


class _Reads(C.Structure):
    _fields_ = [("n", C.c_int), ("seq", C.POINTER(C.c_char_p)), ("len", C.POINTER(C.c_int)), ("name", C.POINTER(C.c_char_p)), ("id", C.POINTER(C.c_uint32))]


@pytest.fixture(scope="module")
def L():
    L = oracle.lib()
    L.ndo_read_2bit.restype = C.POINTER(_Reads); L.ndo_read_2bit.argtypes = [C.c_char_p]
    L.ndo_reads_free.argtypes = [C.POINTER(_Reads)]
    L.ndo_index_build.restype = C.c_void_p
    L.ndo_index_build.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int]
    L.ndo_index_free.argtypes = [C.c_void_p]
    L.ndo_index_mid_occ.restype = C.c_int; L.ndo_index_mid_occ.argtypes = [C.c_void_p, C.c_float]
    L.ndo_collect_seeds.restype = C.c_void_p
    L.ndo_collect_seeds.argtypes = [C.c_void_p, C.POINTER(oracle.Opt), C.c_int, C.c_char_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int64), C.c_int]
    L.ndo_chain_dp.restype = C.c_void_p
    L.ndo_chain_dp.argtypes = [C.c_int] * 7 + [C.c_int64, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]
    L.ndo_gen_regs.restype = C.c_void_p
    L.ndo_gen_regs.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.ndo_query_hash.restype = C.c_uint32; L.ndo_query_hash.argtypes = [C.c_char_p, C.c_int, C.c_int]
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p; libc.malloc.argtypes = [C.c_size_t]
    libc.free.argtypes = [C.c_void_p]
    L._libc = libc
    return L


def _chain(L, o, a, qname, qlen):
    """a: (n, 2) uint64 anchors in the order the chain DP shall see them.  Returns (u, chain anchors, regions) as bytes."""
    n = len(a)
    buf = L._libc.malloc(max(16 * n, 16))                 # ndo_chain_dp consumes (frees) its input
    C.memmove(buf, a.ctypes.data, 16 * n)
    n_u = C.c_int(0); u = C.c_void_p(None)
    b = L.ndo_chain_dp(o.max_gap, o.max_gap, o.bw, o.max_chain_skip, o.max_chain_iter, o.min_cnt, o.min_chain_score, n, buf, C.byref(n_u), C.byref(u))
    if n_u.value == 0:
        if b: L._libc.free(b)
        if u.value: L._libc.free(u)
        return b"", b"", b""
    ub = C.string_at(u.value, 8 * n_u.value)
    n_anch = int((np.frombuffer(ub, dtype=np.uint64) & np.uint64(0xffffffff)).sum())
    ab = C.string_at(b, 16 * n_anch)
    r = L.ndo_gen_regs(L.ndo_query_hash(qname, qlen, o.seed), qlen, n_u.value, u, b)
    rb = C.string_at(r, C.sizeof(oracle.Reg) * n_u.value)
    L._libc.free(r); L._libc.free(b); L._libc.free(u)
    return ub, ab, rb


def _filtered_order(a, o, max_span):
    """What the product's sort stage leaves in a read's anchor range (None: equal keys among the kept anchors -> exact replay of all)."""
    grp = a[:, 0] >> np.uint64(32)
    _, inv, cnt = np.unique(grp, return_inverse=True, return_counts=True)
    c = cnt[inv]
    keep = (c >= o.min_cnt) & (c.astype(np.int64) * max_span >= o.min_chain_score)
    kept, dropped = a[keep], a[~keep]
    if len(np.unique(kept[:, 0])) != len(kept):
        return None, int(keep.sum())
    kept = kept[np.argsort(kept[:, 0], kind="stable")]
    ph = dropped.copy()
    ph[:, 0] = (np.uint64(0xffffffff) << np.uint64(32)) | (np.arange(len(ph), dtype=np.uint64) * np.uint64(o.max_gap + 1))
    return np.concatenate([kept, ph]), int(keep.sum())


@pytest.mark.parametrize("preset,k,w", [("ava-ont", 17, 10), ("ava-ont", None, None), ("ava-pb", None, None)])
def test_chains_do_not_depend_on_stray_groups(L, golden_dir, preset, k, w):
    o = oracle.make_opt(preset, k=k, w=w)
    R = L.ndo_read_2bit(os.path.join(golden_dir, "sub120.2bit").encode())
    assert R and R.contents.n == 120
    r = R.contents
    idx = L.ndo_index_build(r.n, r.seq, r.len, r.name, o.w, o.k, o.hpc)
    mid_occ = L.ndo_index_mid_occ(idx, o.mid_occ_frac)
    max_span = 255 if o.hpc else o.k
    n_filtered = n_fallback = n_chains = n_dropped = 0
    try:
        for q in range(r.n):
            qlen = r.len[q]; qname = r.name[q]
            mv = np.zeros((qlen + 1, 2), dtype=np.uint64)
            n_mv = L.ndo_sketch(r.seq[q], qlen, o.w, o.k, 0, o.hpc, mv.ctypes.data_as(C.c_void_p), 0)
            n_a = C.c_int64(0)
            p = L.ndo_collect_seeds(idx, C.byref(o), mid_occ, qname, qlen, mv.ctypes.data_as(C.c_void_p), n_mv, C.byref(n_a), 0)
            if not p or n_a.value == 0:
                continue
            a = np.frombuffer(C.string_at(p, 16 * n_a.value), dtype=np.uint64).reshape(-1, 2).copy()
            L._libc.free(p)
            full = oracle.rs_sort_128x(a)                                  # the reference order (unstable radix sort, ties and all)
            want = _chain(L, o, full, qname, qlen)
            order, n_kept = _filtered_order(a, o, max_span)
            if order is None:
                n_fallback += 1
                continue                                                    # the product replays the exact sort over all anchors: nothing to show
            assert len(order) == len(a)
            got = _chain(L, o, order, qname, qlen)
            assert got[0] == want[0], "chains (score, count) differ for read %d" % q
            assert got[1] == want[1], "chain anchors differ for read %d" % q
            assert got[2] == want[2], "regions differ for read %d" % q
            n_filtered += 1; n_dropped += len(a) - n_kept
            n_chains += len(want[0]) // 8
    finally:
        L.ndo_index_free(idx); L.ndo_reads_free(R)
    # real reads of a repeat-rich genome: with the short default k most reads do hold equal keys and take the exact replay
    assert n_filtered >= 10 and n_chains > 100 and n_dropped > 0, (n_filtered, n_fallback, n_chains, n_dropped)
    print(preset, k, w, "filtered reads", n_filtered, "exact-replay reads", n_fallback, "chains", n_chains, "dropped anchors", n_dropped)
