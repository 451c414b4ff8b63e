"""ctypes binding of libndgpu.so (include/ndgpu.h) and the step-1 driver that mirrors
minimap2-nd's main loop (ref: minimap2/main.c:477-509, minimap2/index.c:280-372 index parts,
minimap2/map.c:1188-1354 query batches + ordered .ovl output).

No CPU fallback: a missing library or a missing sm_100 device raises NdgError.
"""
import ctypes as C
import os
import numpy as np
from .formats import Reads
from .options import IdxOpt, MapOpt, MM_F_NO_DIAG, MM_F_NO_DUAL, MM_F_CIGAR, MM_F_NO_END_FLT, MM_F_ALL_CHAINS, MM_I_HPC

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    return os.path.join(_HERE, "libndgpu.so")


class NdgError(RuntimeError):
    pass


class _Opt(C.Structure):
    _fields_ = [("k", C.c_int32), ("w", C.c_int32), ("hpc", C.c_int32),
                ("flag_no_diag", C.c_int32), ("flag_no_dual", C.c_int32),
                ("min_cnt", C.c_int32), ("min_chain_score", C.c_int32),
                ("bw", C.c_int32), ("max_gap", C.c_int32),
                ("max_chain_skip", C.c_int32), ("max_chain_iter", C.c_int32),
                ("seed", C.c_int32), ("mid_occ_frac", C.c_float),
                ("mid_occ", C.c_int32), ("min_mid_occ", C.c_int32)]


class _AlnOpt(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("a", "b", "q", "e", "q2", "e2", "sc_ambi", "zdrop", "zdrop_inv", "end_bonus", "min_dp_max",
                                         "min_ksw_len", "no_end_flt")] + [("max_clip_ratio", C.c_float)]


class _Step2Opt(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("minlen", "maxhan1", "maxhan2", "minmatch", "mode", "cn", "kn", "wn", "outctn")] + [("minide", C.c_float)]


class _AlnStat(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_chains", "n_calls", "n_rounds", "n_launches", "n_retries", "rect_cells", "p_bytes")] + \
               [("dp_ms", C.c_double), ("host_ms", C.c_double), ("dp_wall_ms", C.c_double)]


class _Reads(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("words", C.c_void_p), ("word_off", C.c_void_p), ("len", C.c_void_p),
                ("nmask", C.c_void_p), ("name_key", C.c_void_p), ("name_hash", C.c_void_p), ("name_id", C.c_void_p), ("tag", C.c_void_p)]


class _MapResult(C.Structure):
    _fields_ = [("n_hits", C.c_int64), ("hits", C.c_void_p),
                ("n_minimizers", C.c_uint64), ("n_anchors", C.c_uint64), ("n_segments", C.c_uint64), ("device_ms", C.c_double)]


class _IndexStat(C.Structure):
    _fields_ = [("n_seq", C.c_uint64), ("n_minimizers", C.c_uint64), ("n_distinct", C.c_uint64),
                ("n_singletons", C.c_uint64), ("sum_len", C.c_uint64), ("mid_occ", C.c_int32)]


HIT_DTYPE = np.dtype([("zkey", "<u8"), ("qidx", "<i4"), ("rid", "<i4"), ("qs", "<i4"), ("qe", "<i4"),
                      ("rs", "<i4"), ("re", "<i4"), ("mlen", "<i4"), ("blen", "<i4"), ("score", "<i4"),
                      ("cnt", "<i4"), ("rev", "<i4"), ("as_", "<i4")])
assert HIT_DTYPE.itemsize == 56

KSW_CALL_DTYPE = np.dtype([("qlen", "<i4"), ("tlen", "<i4"), ("q_off", "<u8"), ("t_off", "<u8"), ("w", "<i4"), ("zdrop", "<i4"),
                           ("end_bonus", "<i4"), ("flag", "<i4")])
KSW_RES_DTYPE = np.dtype([(n_, "<i4") for n_ in ("max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar",
                                                 "reach_end", "reserved")])
assert KSW_CALL_DTYPE.itemsize == 40 and KSW_RES_DTYPE.itemsize == 48

EXPORTS = ["ndg_init", "ndg_destroy", "ndg_last_error", "ndg_version", "ndg_set_opt", "ndg_set_limits",
           "ndg_set_variant", "ndg_index_begin", "ndg_index_add", "ndg_index_add_meta",
           "ndg_index_key_slice", "ndg_copy_between", "ndg_ovl_rebase_first", "ndg_index_shard_split", "ndg_index_shard_bucket", "ndg_index_shard_recv_reserve", "ndg_index_shard_sort", "ndg_index_shard_place",
           "ndg_index_shard_table_slice", "ndg_index_shard_pos_reserve", "ndg_index_shard_finish", "ndg_index_finalize", "ndg_index_get",
           "ndg_map_batch", "ndg_map_upload", "ndg_map_resident", "ndg_map_download", "ndg_sketch",
           "ndg_debug_fetch", "ndg_prof_get", "ndg_prof_reset", "ndg_kernel_launches", "ndg_ovl_encode8",
           "ndg_ovl_emit_step1", "ndg_ksw_extd2_batch", "ndg_host_alloc", "ndg_host_free", "ndg_map_batch_ovl", "ndg_set_ksw_variant", "ndg_set_aln_opt", "ndg_aln_stats", "ndg_set_threads", "ndg_step2_begin", "ndg_step2_map_batch", "ndg_step2_finish", "ndg_index_share",
           "ndg_set_paf", "ndg_map_extras", "ndg_ovl_sort", "ndg_map_prefetch"]

_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise NdgError("CUDA library %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)" % p)
    L = C.CDLL(p)
    for s in EXPORTS:
        if not hasattr(L, s):
            raise NdgError("libndgpu.so lacks symbol %s" % s)
    L.ndg_last_error.restype = C.c_char_p
    L.ndg_version.restype = C.c_char_p
    L.ndg_debug_fetch.restype = C.c_int64
    L.ndg_kernel_launches.restype = C.c_uint64
    L.ndg_ovl_emit_step1.restype = C.c_int64
    L.ndg_init.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.ndg_destroy.argtypes = [C.c_void_p]
    L.ndg_last_error.argtypes = [C.c_void_p]
    L.ndg_set_opt.argtypes = [C.c_void_p, C.POINTER(_Opt)]
    L.ndg_set_limits.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
    L.ndg_set_variant.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ndg_set_ksw_variant.argtypes = [C.c_void_p, C.c_int]
    L.ndg_set_aln_opt.argtypes = [C.c_void_p, C.c_void_p]
    L.ndg_aln_stats.argtypes = [C.c_void_p, C.c_void_p]
    L.ndg_set_threads.argtypes = [C.c_void_p, C.c_int]
    L.ndg_index_share.argtypes = [C.c_void_p, C.c_void_p]
    L.ndg_set_paf.argtypes = [C.c_void_p, C.c_int]
    L.ndg_map_prefetch.argtypes = [C.c_void_p, C.POINTER(_Reads)]
    L.ndg_ovl_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p] + [C.c_void_p] * 6
    L.ndg_step2_begin.argtypes = [C.c_void_p, C.c_void_p]
    L.ndg_step2_map_batch.argtypes = [C.c_void_p] + [C.c_void_p] * 6
    L.ndg_step2_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ndg_index_begin.argtypes = [C.c_void_p]
    L.ndg_index_add.argtypes = [C.c_void_p, C.POINTER(_Reads)]
    L.ndg_index_add_meta.argtypes = [C.c_void_p, C.POINTER(_Reads)]
    L.ndg_ovl_rebase_first.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ndg_copy_between.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.ndg_index_key_slice.argtypes = [C.c_uint64, C.c_int32, C.c_int32]
    L.ndg_index_key_slice.restype = C.c_int32
    L.ndg_index_shard_split.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    L.ndg_index_shard_bucket.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.ndg_index_shard_recv_reserve.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.ndg_index_shard_sort.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.ndg_index_shard_place.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.c_void_p]
    L.ndg_index_shard_table_slice.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.ndg_index_shard_pos_reserve.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.ndg_index_shard_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(_IndexStat)]
    L.ndg_index_finalize.argtypes = [C.c_void_p, C.POINTER(_IndexStat)]
    L.ndg_index_get.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    L.ndg_map_batch.argtypes = [C.c_void_p, C.POINTER(_Reads), C.POINTER(_MapResult)]
    L.ndg_map_upload.argtypes = [C.c_void_p, C.POINTER(_Reads)]
    L.ndg_map_resident.argtypes = [C.c_void_p, C.POINTER(_MapResult)]
    L.ndg_map_download.argtypes = [C.c_void_p, C.POINTER(_MapResult)]
    L.ndg_sketch.argtypes = [C.c_void_p, C.POINTER(_Reads), C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
    L.ndg_debug_fetch.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint64]
    L.ndg_prof_get.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.ndg_prof_reset.argtypes = [C.c_void_p]
    L.ndg_kernel_launches.argtypes = [C.c_void_p]
    L.ndg_ovl_encode8.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] + [C.c_uint32] * 7
    L.ndg_ovl_emit_step1.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 6 + [C.c_int32] * 4 + \
                                    [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    L.ndg_ksw_extd2_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_uint64, C.c_int32, C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p] * 3
    L.ndg_map_batch_ovl.argtypes = [C.c_void_p, C.POINTER(_Reads), C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
                                    C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(_MapResult)]
    L.ndg_host_alloc.restype = C.c_void_p
    L.ndg_host_alloc.argtypes = [C.c_uint64]
    L.ndg_host_free.argtypes = [C.c_void_p]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _reads_struct(r: Reads):
    # keep references alive on the struct object
    s = _Reads()
    s._keep = (np.ascontiguousarray(r.words, dtype=np.uint32), np.ascontiguousarray(r.word_off, dtype=np.uint64),
               np.ascontiguousarray(r.length, dtype=np.uint32), np.ascontiguousarray(r.name_key, dtype=np.uint64),
               np.ascontiguousarray(r.name_hash, dtype=np.uint32), np.ascontiguousarray(r.ids, dtype=np.uint32),
               None if getattr(r, "nmask", None) is None else np.ascontiguousarray(r.nmask, dtype=np.uint32))
    s.n_reads = r.n
    s.words, s.word_off, s.len = _ptr(s._keep[0]), _ptr(s._keep[1]), _ptr(s._keep[2])
    s.nmask = _ptr(s._keep[6])
    s.name_key, s.name_hash, s.name_id = _ptr(s._keep[3]), _ptr(s._keep[4]), _ptr(s._keep[5])
    return s


class Engine:
    """One GPU context.  Mirrors the life cycle of one `minimap2-nd --step 1` process."""

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.ndg_init(device, C.byref(h))
        if rc != 0:
            raise NdgError("ndg_init(device=%d) failed with %d: no usable sm_100 GPU (no CPU fallback exists)" % (device, rc))
        self.h = h
        self.device = device
        self.io, self.mo = IdxOpt(), MapOpt()
        self.targets = None
        self.stat = None

    def close(self):
        if self.h:
            self.L.ndg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise NdgError("libndgpu error %d: %s" % (rc, self.L.ndg_last_error(self.h).decode()))

    def set_options(self, io: IdxOpt, mo: MapOpt):
        if not mo.flag & MM_F_ALL_CHAINS:
            # the library keeps every chain; the reference would prune secondaries here (mm_set_parent / mm_select_sub with
            # best_n and pri_ratio, ref: minimap2/map.c:250-259) and its .ovl bytes would differ
            raise NdgError("the GPU path maps in all-chain mode only (MM_F_ALL_CHAINS): use the ava-ont / ava-pb presets, -X or -P")
        self.io, self.mo = io, mo
        if mo.flag & MM_F_CIGAR:                        # regime B (ref: minimap2/map.c:484-505)
            ao = _AlnOpt(a=mo.a, b=mo.b, q=mo.q, e=mo.e, q2=mo.q2, e2=mo.e2, sc_ambi=mo.sc_ambi, zdrop=mo.zdrop, zdrop_inv=mo.zdrop_inv,
                         end_bonus=mo.end_bonus, min_dp_max=mo.min_dp_max, min_ksw_len=mo.min_ksw_len,
                         no_end_flt=1 if mo.flag & MM_F_NO_END_FLT else 0, max_clip_ratio=mo.max_clip_ratio)
            self._ck(self.L.ndg_set_aln_opt(self.h, C.byref(ao)))
        else:
            self._ck(self.L.ndg_set_aln_opt(self.h, None))
        o = _Opt(k=io.k, w=io.w, hpc=1 if io.flag & MM_I_HPC else 0,
                 flag_no_diag=1 if mo.flag & MM_F_NO_DIAG else 0, flag_no_dual=1 if mo.flag & MM_F_NO_DUAL else 0,
                 min_cnt=mo.min_cnt, min_chain_score=mo.min_chain_score, bw=mo.bw, max_gap=mo.max_gap,
                 max_chain_skip=mo.max_chain_skip, max_chain_iter=mo.max_chain_iter, seed=mo.seed,
                 mid_occ_frac=mo.mid_occ_frac, mid_occ=mo.mid_occ, min_mid_occ=mo.min_mid_occ)
        self._ck(self.L.ndg_set_opt(self.h, C.byref(o)))

    def set_limits(self, sub_batch_bases=0, max_anchors=0):
        self._ck(self.L.ndg_set_limits(self.h, sub_batch_bases, max_anchors))

    def set_variant(self, sort_variant=1, chain_variant=1):
        self._ck(self.L.ndg_set_variant(self.h, sort_variant, chain_variant))

    # ---- index -------------------------------------------------------------------------------
    def build_index(self, targets: Reads):
        self._ck(self.L.ndg_index_begin(self.h))
        s = _reads_struct(targets)
        self._ck(self.L.ndg_index_add(self.h, C.byref(s)))
        return self.finalize_index(targets)

    def share_index_of(self, other):
        """borrow the finalized index of another Engine on the same device (ndg_index_share)"""
        self._ck(self.L.ndg_index_share(self.h, other.h))
        self.targets, self.stat = other.targets, other.stat

    def finalize_index(self, targets: Reads):
        st = _IndexStat()
        self._ck(self.L.ndg_index_finalize(self.h, C.byref(st)))
        self.targets = targets
        self.stat = dict(n_seq=st.n_seq, n_minimizers=st.n_minimizers, n_distinct=st.n_distinct,
                         n_singletons=st.n_singletons, sum_len=st.sum_len, mid_occ=st.mid_occ)
        return self.stat

    def index_get(self, minier, cap=4096):
        pos = np.zeros(cap, dtype=np.uint64)
        n = C.c_int()
        self._ck(self.L.ndg_index_get(self.h, int(minier), _ptr(pos), cap, C.byref(n)))
        return pos[:min(n.value, cap)].copy(), n.value

    # ---- sketch (tests) ----------------------------------------------------------------------
    def sketch(self, reads: Reads, rid0=0):
        cap = int(reads.n_bases) + reads.n + 16
        xy = np.zeros((cap, 2), dtype=np.uint64)
        off = np.zeros(reads.n + 1, dtype=np.uint64)
        s = _reads_struct(reads)
        self._ck(self.L.ndg_sketch(self.h, C.byref(s), rid0, _ptr(xy), cap, _ptr(off)))
        return xy[:int(off[-1])].copy(), off

    # ---- mapping -----------------------------------------------------------------------------
    def _result(self, res):
        n = res.n_hits
        if n and res.hits:
            buf = (C.c_char * (n * HIT_DTYPE.itemsize)).from_address(res.hits)
            hits = np.frombuffer(buf, dtype=HIT_DTYPE, count=n).copy()
        else:
            hits = np.zeros(0, dtype=HIT_DTYPE)
        return hits, dict(n_minimizers=res.n_minimizers, n_anchors=res.n_anchors, n_segments=res.n_segments)

    def map_batch(self, queries: Reads):
        """End-to-end call with host buffers: H2D, kernels, D2H (ndg_map_batch)."""
        s = _reads_struct(queries)
        res = _MapResult()
        self._ck(self.L.ndg_map_batch(self.h, C.byref(s), C.byref(res)))
        return self._result(res)

    def map_batch_ovl(self, queries: Reads, state):
        """Whole step 1 for a batch with the output stage on the GPU (ndg_map_batch_ovl): returns (.ovl bytes, stats)."""
        s = _reads_struct(queries)
        eo = np.array([self.mo.minlen, self.mo.dvt, self.mo.maxhan1, self.mo.maxhan2], dtype=np.int32)
        ptr, nb, nrec, sq = C.c_void_p(), C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        res = _MapResult()
        self._ck(self.L.ndg_map_batch_ovl(self.h, C.byref(s), _ptr(eo), _ptr(state), C.byref(ptr), C.byref(nb), C.byref(nrec), C.byref(sq), C.byref(res)))
        data = C.string_at(ptr.value, nb.value) if nb.value else b""
        return data, dict(records=nrec.value, sum_qspan=sq.value, n_hits=res.n_hits, n_minimizers=res.n_minimizers,
                          n_anchors=res.n_anchors, n_segments=res.n_segments)

    def map_prefetch(self, queries: Reads):
        """start the upload of the NEXT batch; the following map_batch_ovl(queries, ...) finds it resident (ndg_map_prefetch)"""
        s = _reads_struct(queries)
        keep = getattr(self, "_pf_keep", [])
        keep.append((s, queries)); self._pf_keep = keep[-3:]           # the library keeps the pointers until the batch has been mapped
        self._ck(self.L.ndg_map_prefetch(self.h, C.byref(s)))

    def map_upload(self, queries: Reads):
        s = _reads_struct(queries)
        self._ck(self.L.ndg_map_upload(self.h, C.byref(s)))

    def map_resident(self):
        res = _MapResult()
        self._ck(self.L.ndg_map_resident(self.h, C.byref(res)))
        return dict(n_hits=res.n_hits, n_minimizers=res.n_minimizers, n_anchors=res.n_anchors, n_segments=res.n_segments, device_ms=res.device_ms)

    def map_download(self):
        res = _MapResult()
        self._ck(self.L.ndg_map_download(self.h, C.byref(res)))
        return self._result(res)

    def debug_fetch(self, what, dtype):
        nb = self.L.ndg_debug_fetch(self.h, what.encode(), None, 0)
        if nb < 0:
            self._ck(int(nb))
        out = np.zeros(nb // np.dtype(dtype).itemsize, dtype=dtype)
        if nb:
            rc = self.L.ndg_debug_fetch(self.h, what.encode(), _ptr(out), nb)
            if rc < 0:
                self._ck(int(rc))
        return out

    # ---- base-level extension (ksw_extd2) ----------------------------------------------------
    # ---- --step 2 -----------------------------------------------------------------------------
    def step2_begin(self, mo: MapOpt):
        o = _Step2Opt(minlen=mo.minlen, maxhan1=mo.maxhan1, maxhan2=mo.maxhan2, minmatch=mo.minmatch, mode=mo.mode, cn=mo.cn, kn=mo.kn, wn=mo.wn,
                      outctn=mo.outctn, minide=mo.minide)
        self._ck(self.L.ndg_step2_begin(self.h, C.byref(o)))

    def step2_end(self):
        self._ck(self.L.ndg_step2_begin(self.h, None))

    def step2_map_batch(self, queries: Reads):
        s = _reads_struct(queries)
        ptr, nb, nrec, sq = C.c_void_p(), C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        res = _MapResult()
        self._ck(self.L.ndg_step2_map_batch(self.h, C.byref(s), C.byref(ptr), C.byref(nb), C.byref(nrec), C.byref(sq), C.byref(res)))
        data = C.string_at(ptr.value, nb.value) if nb.value else b""
        return data, dict(records=nrec.value, sum_qspan=sq.value, n_hits=res.n_hits, n_minimizers=res.n_minimizers, n_anchors=res.n_anchors)

    def step2_finish(self):
        ptr, nb = C.c_char_p(), C.c_uint64(0)
        self._ck(self.L.ndg_step2_finish(self.h, C.byref(ptr), C.byref(nb)))
        return C.string_at(ptr, nb.value) if nb.value else b""

    def set_threads(self, n=0):
        return self.L.ndg_set_threads(self.h, n)

    def aln_stats(self):
        st = _AlnStat()
        self._ck(self.L.ndg_aln_stats(self.h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in _AlnStat._fields_}

    def set_ksw_variant(self, v):
        self._ck(self.L.ndg_set_ksw_variant(self.h, v))

    def ksw_extd2_batch(self, calls, seqs, mat, q, e, q2, e2, m=5):
        """calls: structured array KSW_CALL_DTYPE; seqs: uint8 codes.  Returns (results KSW_RES_DTYPE, cigars list)."""
        n = calls.shape[0]
        calls = np.ascontiguousarray(calls)
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        mat = np.ascontiguousarray(mat, dtype=np.int8)
        coff = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(calls["qlen"].astype(np.uint64) + calls["tlen"].astype(np.uint64) + np.uint64(2), out=coff[1:])
        res = np.zeros(n, dtype=KSW_RES_DTYPE)
        cig = np.zeros(int(coff[-1]) + 4, dtype=np.uint32)
        self._ck(self.L.ndg_ksw_extd2_batch(self.h, _ptr(calls), n, _ptr(seqs), seqs.shape[0], m, _ptr(mat), q, e, q2, e2,
                                            _ptr(res), _ptr(cig), _ptr(coff)))
        return res, [cig[int(coff[i]):int(coff[i]) + int(res["n_cigar"][i])] for i in range(n)]

    def ovl_sort(self, streams, seed_names, seed_lens, max_bin_cov=40, max_flank_len=300, is_hq=False):
        """`ovl_sort` on the GPU (ref: util/ovl_sort.c; ndg_ovl_sort): streams = the bytes of the input 8-field .ovl files in fofn order,
        seeds = the (name, length) list of the `-i` index.  Returns (sorted .ovl bytes, .bl text, records read, records written)."""
        n = len(streams)
        bufs = [np.frombuffer(b, dtype=np.uint8) if len(b) else np.zeros(1, dtype=np.uint8) for b in streams]
        ptrs = (C.c_void_p * max(n, 1))(*[x.ctypes.data for x in bufs])
        sizes = np.array([len(b) for b in streams], dtype=np.uint64) if n else np.zeros(1, dtype=np.uint64)
        sn = np.ascontiguousarray(seed_names, dtype=np.uint32); sl = np.ascontiguousarray(seed_lens, dtype=np.uint32)
        opt = (C.c_int32 * 3)(int(max_bin_cov), int(max_flank_len), 1 if is_hq else 0)
        ovl, nb, bl, nbl, n_in, n_out = C.c_void_p(), C.c_uint64(), C.c_char_p(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(self.L.ndg_ovl_sort(self.h, ptrs, _ptr(sizes), n, _ptr(sn), _ptr(sl), C.c_uint64(sn.shape[0]), opt,
                                     C.byref(ovl), C.byref(nb), C.byref(bl), C.byref(nbl), C.byref(n_in), C.byref(n_out)))
        data = C.string_at(ovl.value, nb.value) if nb.value else b""
        text = C.string_at(bl, nbl.value) if nbl.value else b""
        return data, text, int(n_in.value), int(n_out.value)

    def prof(self):
        names = (C.c_char_p * 64)()
        ms = (C.c_double * 64)()
        ln = (C.c_uint64 * 64)()
        n = self.L.ndg_prof_get(self.h, 64, names, ms, ln)
        return {names[i].decode(): (ms[i], ln[i]) for i in range(n)}

    def prof_reset(self):
        self.L.ndg_prof_reset(self.h)

    def kernel_launches(self):
        return int(self.L.ndg_kernel_launches(self.h))

    # ---- step 1 ------------------------------------------------------------------------------
    def emit_step1(self, hits, queries: Reads, state, out: bytearray):
        """Apply the step-1 filter and append 8-field records (ndg_ovl_emit_step1)."""
        t = self.targets
        cap = hits.shape[0] * 40 + 64
        dst = np.zeros(cap, dtype=np.uint8)
        n_rec, sq = C.c_uint64(0), C.c_uint64(0)
        hits = np.ascontiguousarray(hits)
        nb = self.L.ndg_ovl_emit_step1(_ptr(hits), hits.shape[0], _ptr(queries.ids), _ptr(queries.name_key), _ptr(queries.length),
                                       _ptr(t.ids), _ptr(t.name_key), _ptr(t.length), self.mo.minlen, self.mo.dvt,
                                       self.mo.maxhan1, self.mo.maxhan2, _ptr(state), _ptr(dst), cap, C.byref(n_rec), C.byref(sq))
        if nb < 0:
            self._ck(int(nb))
        out += dst[:nb].tobytes()
        return n_rec.value, sq.value


class PinnedReads:
    """A copy of a read batch whose arrays live in page-locked host memory (ndg_host_alloc)."""

    def __init__(self, reads: Reads):
        self.L = load_library()
        self._ptrs = []
        def pin(a):
            a = np.ascontiguousarray(a)
            p = self.L.ndg_host_alloc(max(a.nbytes, 16))
            if not p:
                raise NdgError("ndg_host_alloc failed")
            self._ptrs.append(p)
            out = np.frombuffer((C.c_char * a.nbytes).from_address(p), dtype=a.dtype, count=a.size)
            out[...] = a.reshape(-1)
            return out
        self.reads = Reads(pin(reads.words), pin(reads.word_off), pin(reads.length), pin(reads.ids), pin(reads.name_key), pin(reads.name_hash))

    def free(self):
        self.reads = None
        for p in self._ptrs:
            self.L.ndg_host_free(p)
        self._ptrs = []


def stitch_ovl(chunks, states, state):
    """Concatenate .ovl batches that were each encoded from state (0, 0): the first record of every batch is re-based onto the names the
    stream really ended with (ndg_ovl_rebase_first).  chunks[i] = bytes of batch i, states[i] = its final (qname, tname); `state`
    (uint32[2]) is the running state of the stream, updated in place.  Returns the stitched bytes."""
    L = load_library()
    out = bytearray()
    first = np.zeros(48, dtype=np.uint8)
    skip = C.c_uint32(0)
    for data, st in zip(chunks, states):
        if not data:
            continue
        buf = np.frombuffer(data, dtype=np.uint8)
        n = L.ndg_ovl_rebase_first(_ptr(buf), len(data), _ptr(state), _ptr(first), C.byref(skip))
        if n < 0:
            raise NdgError("malformed .ovl batch")
        out += first[:n].tobytes()
        out += data[skip.value:]
        state[0], state[1] = st[0], st[1]
    return bytes(out)


def map_batches_pool(engs, batches):
    """Map query batches with a pool of contexts (one host thread each; they share an index through ndg_index_share or hold replicas on
    several GPUs), each batch encoded from state (0, 0).  Upload, kernels and download of neighbouring batches overlap on the contexts'
    streams.  Returns (chunks, states, stats) in batch order; stitch with stitch_ovl()."""
    import threading
    n = len(batches)
    chunks, states, stats = [None] * n, [None] * n, [None] * n
    nxt = [0]
    lock = threading.Lock()
    err = []

    def work(e):
        while True:
            with lock:
                i = nxt[0]; nxt[0] += 1
            if i >= n or err:
                return
            try:
                st = np.zeros(2, dtype=np.uint32)
                data, s = e.map_batch_ovl(batches[i], st)
                chunks[i], states[i], stats[i] = data, (int(st[0]), int(st[1])), s
            except Exception as ex:  # noqa: BLE001
                err.append(ex)
    ts = [threading.Thread(target=work, args=(e,)) for e in engs]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if err:
        raise err[0]
    return chunks, states, stats


def shard_bounds(n, world):
    """contiguous, near-equal read ranges per rank"""
    return [(n * r) // world for r in range(world + 1)]


class _DevView:
    """zero-copy torch view of library-owned device memory (CUDA array interface v2)"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


def _meta_struct(ids, lengths):
    """ndg_reads_t carrying only what ndg_index_add_meta() reads (names and lengths of targets sketched on another rank)"""
    from .formats import name_keys, x31_hash
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    s = _Reads()
    s._keep = (np.ascontiguousarray(lengths, dtype=np.uint32), name_keys(ids), x31_hash(ids), ids)
    s.n_reads = ids.shape[0]
    s.len, s.name_key, s.name_hash, s.name_id = _ptr(s._keep[0]), _ptr(s._keep[1]), _ptr(s._keep[2]), _ptr(s._keep[3])
    return s


def shard_exchange_plan(counts_matrix, rank):
    """counts_matrix[r][j] = pairs rank r holds for rank j.  Returns (input_split_sizes, output_split_sizes, total_received) of
    rank `rank` for the all-to-all of the sharded index build (pure host logic; also exercised by the gloo test)."""
    m = np.asarray(counts_matrix, dtype=np.int64)
    return [int(x) for x in m[rank]], [int(x) for x in m[:, rank]], int(m[:, rank].sum())


def build_index_sharded(eng: Engine, mine: Reads, rank, world, dist, torch, meta=None, timings=None):
    """N-GPU index build (ndg_index_shard_*): rank r sketches its own targets, the minimizers are exchanged by key
    range (one all-to-all), every rank sorts 1/N of them and builds its slice of the ordered table, and the finished slices and
    position segments are gathered.  Target order = rank order; read ids (rid) are global, so mapping results equal those of a
    single-GPU index over the concatenated target set.  meta[r] = (ids, lengths) of rank r's targets (None entries are gathered)."""
    L, h = eng.L, eng.h
    import time as _t
    t0 = _t.perf_counter()
    dev = torch.device("cuda", torch.cuda.current_device())
    if meta is None:                                    # names/lengths of everybody's targets
        meta = [None] * world
        cnt = torch.tensor([mine.n], dtype=torch.int64, device=dev)
        cnts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(cnts, cnt)
        for r in range(world):
            n_r = int(cnts[r].item())
            buf = torch.empty(2 * n_r, dtype=torch.int64, device=dev)
            if r == rank:
                buf[:n_r] = torch.from_numpy(mine.ids.astype(np.int64)).to(dev); buf[n_r:] = torch.from_numpy(mine.length.astype(np.int64)).to(dev)
            dist.broadcast(buf, src=r)
            b = buf.cpu().numpy()
            meta[r] = (b[:n_r].astype(np.uint32), b[n_r:].astype(np.uint32))
    eng._ck(L.ndg_index_begin(h))
    keep = []
    for r in range(world):
        if r == rank:
            s = _reads_struct(mine); keep.append(s)
            eng._ck(L.ndg_index_add(h, C.byref(s)))
        else:
            s = _meta_struct(*meta[r]); keep.append(s)
            if s.n_reads:
                eng._ck(L.ndg_index_add_meta(h, C.byref(s)))
    t1 = _t.perf_counter()
    counts = np.zeros(world, dtype=np.uint64)
    eng._ck(L.ndg_index_shard_split(h, world, _ptr(counts)))
    cm = torch.zeros(world * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(cm, torch.from_numpy(counts.astype(np.int64)).to(dev))
    cm = cm.cpu().numpy().reshape(world, world)
    in_split, out_split, total = shard_exchange_plan(cm, rank)
    kp, vp, n0 = C.c_void_p(), C.c_void_p(), C.c_uint64()
    eng._ck(L.ndg_index_shard_bucket(h, 0, C.byref(kp), C.byref(vp), C.byref(n0)))     # bucket 0 starts the sorted local arrays
    rk, rv = C.c_void_p(), C.c_void_p()
    eng._ck(L.ndg_index_shard_recv_reserve(h, total, C.byref(rk), C.byref(rv)))
    n_local = int(sum(in_split))
    empty = torch.empty(0, dtype=torch.int64, device=dev)
    for src_p, dst_p in ((kp, rk), (vp, rv)):
        inp = torch.as_tensor(_DevView(src_p.value, n_local), device=dev) if n_local else empty
        out = torch.as_tensor(_DevView(dst_p.value, total), device=dev) if total else empty
        dist.all_to_all_single(out, inp, output_split_sizes=out_split, input_split_sizes=in_split)
    torch.cuda.synchronize()
    t2 = _t.perf_counter()
    nd = C.c_uint64()
    eng._ck(L.ndg_index_shard_sort(h, C.byref(nd)))
    st = torch.zeros(2 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(st, torch.tensor([nd.value, total], dtype=torch.int64, device=dev))
    st = st.cpu().numpy().reshape(world, 2)
    hist = np.zeros(65536, dtype=np.uint32)
    eng._ck(L.ndg_index_shard_place(h, rank, int(st[:, 0].max()), _ptr(hist)))
    t3 = _t.perf_counter()
    sp, sb = C.c_void_p(), C.c_uint64()
    eng._ck(L.ndg_index_shard_table_slice(h, 0, C.byref(sp), C.byref(sb)))
    words = sb.value // 8
    full = torch.as_tensor(_DevView(sp.value, words * world), device=dev)
    dist.all_gather_into_tensor(full, full[rank * words:(rank + 1) * words])             # in place: slice r already sits at its offset
    npos = np.ascontiguousarray(st[:, 1].astype(np.uint64))
    segs = (C.c_void_p * world)()
    eng._ck(L.ndg_index_shard_pos_reserve(h, rank, _ptr(npos), segs))
    for r in range(world):
        if int(npos[r]):
            dist.broadcast(torch.as_tensor(_DevView(segs[r], int(npos[r])), device=dev), src=r)
    hs = torch.from_numpy(hist.astype(np.int64)).to(dev)
    dist.all_reduce(hs)
    torch.cuda.synchronize()
    hist = np.ascontiguousarray(hs.cpu().numpy().astype(np.uint32))
    stt = _IndexStat()
    eng._ck(L.ndg_index_shard_finish(h, _ptr(hist), int(st[:, 0].sum()), C.byref(stt)))
    t4 = _t.perf_counter()
    eng.targets = None
    eng.stat = dict(n_seq=stt.n_seq, n_minimizers=stt.n_minimizers, n_distinct=stt.n_distinct, n_singletons=stt.n_singletons, sum_len=stt.sum_len, mid_occ=stt.mid_occ)
    if timings is not None:
        timings.update(sketch_s=t1 - t0, exchange_s=t2 - t1, sort_place_s=t3 - t2, gather_s=t4 - t3)
    return eng.stat


def build_index_sharded_local(engs, parts, torch):
    """The same N-way build with the N contexts living in ONE process (any mix of devices): the collectives become device-to-device
    copies between the contexts' buffers.  This is what `bin/minimap2-nd --gpus N` does in C++ (peer copies); here it lets the whole
    ndg_index_shard_* sequence be exercised on a single GPU (two contexts on device 0).  parts[r] = the reads of "rank" r."""
    world = len(engs)
    L = engs[0].L

    def view(ptr, n, e):
        return torch.as_tensor(_DevView(ptr, n), device=torch.device("cuda", e.device))

    keep = []
    for r, e in enumerate(engs):
        e._ck(L.ndg_index_begin(e.h))
        for j in range(world):
            if j == r:
                s = _reads_struct(parts[j]); keep.append(s)
                e._ck(L.ndg_index_add(e.h, C.byref(s)))
            else:
                s = _meta_struct(parts[j].ids, parts[j].length); keep.append(s)
                e._ck(L.ndg_index_add_meta(e.h, C.byref(s)))
    cm = np.zeros((world, world), dtype=np.uint64)
    for r, e in enumerate(engs):
        row = np.zeros(world, dtype=np.uint64)
        e._ck(L.ndg_index_shard_split(e.h, world, _ptr(row)))
        cm[r] = row
    for j, e in enumerate(engs):                         # "all-to-all": rank j receives bucket j of every rank, in rank order
        _, out_split, total = shard_exchange_plan(cm, j)
        rk, rv = C.c_void_p(), C.c_void_p()
        e._ck(L.ndg_index_shard_recv_reserve(e.h, total, C.byref(rk), C.byref(rv)))
        at = 0
        for r, src in enumerate(engs):
            kp, vp, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
            src._ck(L.ndg_index_shard_bucket(src.h, j, C.byref(kp), C.byref(vp), C.byref(n)))
            if n.value:
                view(rk.value + 8 * at, n.value, e).copy_(view(kp.value, n.value, src))
                view(rv.value + 8 * at, n.value, e).copy_(view(vp.value, n.value, src))
            at += n.value
        assert at == total
    torch.cuda.synchronize()
    nds = []
    for e in engs:
        nd_ = C.c_uint64()
        e._ck(L.ndg_index_shard_sort(e.h, C.byref(nd_)))
        nds.append(nd_.value)
    hist = np.zeros(65536, dtype=np.uint64)
    for r, e in enumerate(engs):
        hr = np.zeros(65536, dtype=np.uint32)
        e._ck(L.ndg_index_shard_place(e.h, r, max(nds), _ptr(hr)))
        hist += hr
    npos = np.ascontiguousarray(cm.sum(axis=0).astype(np.uint64))
    slices, segs = [], []
    for r, e in enumerate(engs):
        sp, sb = C.c_void_p(), C.c_uint64()
        e._ck(L.ndg_index_shard_table_slice(e.h, 0, C.byref(sp), C.byref(sb)))
        slices.append((sp.value, sb.value))
        sg = (C.c_void_p * world)()
        e._ck(L.ndg_index_shard_pos_reserve(e.h, r, _ptr(npos), sg))
        segs.append([sg[j] for j in range(world)])
    for r, e in enumerate(engs):                         # "all-gather": everybody fetches slice j / segment j from rank j
        for j, src in enumerate(engs):
            if j == r:
                continue
            w = slices[j][1] // 8
            view(slices[r][0] + j * slices[r][1], w, e).copy_(view(slices[j][0] + j * slices[j][1], w, src))
            if int(npos[j]):
                view(segs[r][j], int(npos[j]), e).copy_(view(segs[j][j], int(npos[j]), src))
    torch.cuda.synchronize()
    h32 = np.ascontiguousarray(hist.astype(np.uint32))
    stats = []
    for e in engs:
        stt = _IndexStat()
        e._ck(L.ndg_index_shard_finish(e.h, _ptr(h32), int(sum(nds)), C.byref(stt)))
        e.targets = None
        e.stat = dict(n_seq=stt.n_seq, n_minimizers=stt.n_minimizers, n_distinct=stt.n_distinct, n_singletons=stt.n_singletons, sum_len=stt.sum_len, mid_occ=stt.mid_occ)
        stats.append(e.stat)
    return stats


def index_parts(targets: Reads, io: IdxOpt):
    """Split the target file into index parts exactly as mm_idx_gen does
    (ref: minimap2/index.c:284-288,353-358: read mini-batches of >= mini_batch_size bases until the
    running total exceeds batch_size)."""
    mini = min(io.mini_batch_size, io.batch_size)
    parts, r, n = [], 0, targets.n
    lens = targets.length
    while r < n:
        start, total = r, 0
        while r < n and not total > io.batch_size:
            size = 0
            while r < n:
                size += int(lens[r]); r += 1
                if size >= mini:
                    break
            total += size
        parts.append((start, r))
    return parts


def query_batches(queries: Reads, mini_batch_size):
    """ref: minimap2/bseq.c:86-123 (mm_bseq_read3: stop once >= chunk_size bases were read)."""
    out, r, n = [], 0, queries.n
    while r < n:
        start, size = r, 0
        while r < n:
            size += int(queries.length[r]); r += 1
            if size >= mini_batch_size:
                break
        out.append((start, r))
    return out


def run_step2(engine: Engine, targets: Reads, queries: Reads, io: IdxOpt, mo: MapOpt):
    """`minimap2-nd --step 2 ... TARGET QUERY -o OUT` (ref: minimap2/main.c:452-526).  Returns (ovl_bytes, bl_text, stats)."""
    engine.set_options(io, mo)
    engine.step2_begin(mo)
    out = bytearray()
    n_rec = 0
    try:
        for (t0, t1) in index_parts(targets, io):
            engine.build_index(targets.slice(t0, t1))
            for (q0, q1) in query_batches(queries, mo.mini_batch_size):
                data, st = engine.step2_map_batch(queries.slice(q0, q1))
                out += data
                n_rec += st["records"]
        bl = engine.step2_finish()
    finally:
        engine.step2_end()
    return bytes(out), bl, dict(records=n_rec)


def run_step1(engine: Engine, targets: Reads, queries: Reads, io: IdxOpt, mo: MapOpt, same_file=False, device_emit=True):
    """`minimap2-nd --step 1 ... TARGET QUERY -o OUT` (ref: minimap2/main.c:452-509).  Returns
    (ovl_bytes, stats)."""
    engine.set_options(io, mo)
    out = bytearray()
    state = np.zeros(2, dtype=np.uint32)
    n_rec = sq = 0
    work = dict(n_minimizers=0, n_anchors=0, n_segments=0, n_hits=0)
    for (t0, t1) in index_parts(targets, io):
        tp = targets.slice(t0, t1)
        engine.build_index(tp)
        for (q0, q1) in query_batches(queries, mo.mini_batch_size):
            qb = queries.slice(q0, q1)
            if device_emit:
                data, w = engine.map_batch_ovl(qb, state)
                out += data
                r, s = w["records"], w["sum_qspan"]
                work["n_hits"] += w["n_hits"]
            else:
                hits, w = engine.map_batch(qb)
                work["n_hits"] += hits.shape[0]
                r, s = engine.emit_step1(hits, qb, state, out)
            for k in ("n_minimizers", "n_anchors", "n_segments"):
                work[k] += w[k]
            n_rec += r; sq += s
    return bytes(out), dict(records=n_rec, sum_qspan=sq, mid_occ=engine.stat["mid_occ"], **work)
