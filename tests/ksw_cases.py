"""Shared generators / bindings for the ksw_extd2 parity tests (test infrastructure)."""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def simple_mat(a=2, b=4, ambi=1):
    """ref: minimap2/align.c:9-22 (ksw_gen_simple_mat), m = 5"""
    mat = np.zeros(25, dtype=np.int8)
    for i in range(4):
        for j in range(4):
            mat[i * 5 + j] = a if i == j else -b
        mat[i * 5 + 4] = -ambi
    mat[20:25] = -ambi
    return mat


class Res(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar", "reach_end")]


class Ez(C.Structure):                      # ksw_extz_t, ref: minimap2/ksw2.h:23-32
    _fields_ = [("max_zd", C.c_uint32), ("max_q", C.c_int), ("max_t", C.c_int), ("mqe", C.c_int), ("mqe_t", C.c_int), ("mte", C.c_int),
                ("mte_q", C.c_int), ("score", C.c_int), ("m_cigar", C.c_int), ("n_cigar", C.c_int), ("reach_end", C.c_int),
                ("cigar", C.POINTER(C.c_uint32))]


def oracle_lib():
    from oracle import oracle
    L = oracle.lib()
    L.ndk_extd2.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 8 + [C.POINTER(Res), C.c_void_p]
    return L


def oracle_call(L, qs, ts, mat, q, e, q2, e2, w, zdrop, eb, flag):
    r = Res()
    cg = np.zeros(len(qs) + len(ts) + 4, dtype=np.uint32)
    L.ndk_extd2(len(qs), qs.ctypes.data, len(ts), ts.ctypes.data, 5, mat.ctypes.data, q, e, q2, e2, w, zdrop, eb, flag, C.byref(r), cg.ctypes.data)
    return (r.max, r.zdropped, r.max_q, r.max_t, r.mqe, r.mqe_t, r.mte, r.mte_q, r.score, r.n_cigar, r.reach_end), [int(x) for x in cg[:r.n_cigar]]


def ref_lib():
    p = os.path.join(ROOT, "oracle", "_ref", "libminimap2_ref.so")
    if not os.path.exists(p):
        return None
    R = C.CDLL(p)
    R.ksw_extd2_sse.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int8, C.c_void_p] + [C.c_int8] * 4 + [C.c_int] * 4 + [C.POINTER(Ez)]
    return R


def ref_call(R, qs, ts, mat, q, e, q2, e2, w, zdrop, eb, flag):
    ez = Ez(); ez.m_cigar = 0; ez.cigar = None
    R.ksw_extd2_sse(None, len(qs), qs.ctypes.data, len(ts), ts.ctypes.data, 5, mat.ctypes.data, q, e, q2, e2, w, zdrop, eb, flag, C.byref(ez))
    cg = [int(ez.cigar[i]) for i in range(ez.n_cigar)]
    if ez.cigar:
        C.CDLL(None).free(ez.cigar)
    return (ez.max_zd & 0x7fffffff, ez.max_zd >> 31, ez.max_q, ez.max_t, ez.mqe, ez.mqe_t, ez.mte, ez.mte_q, ez.score, ez.n_cigar, ez.reach_end), cg


FLAGS = [0x08, 0, 0x40 | 0x02 | 0x80, 0x40, 0x01, 0x08 | 0x10, 0x02, 0x40 | 0x08]   # the combinations mm_align1 uses, and more


def random_cases(n, seed=5, max_len=1500):
    """(query codes, target codes, w, zdrop, end_bonus, flag) tuples: related pairs with ONT-like errors, ambiguous
    bases, unrelated tails (z-drop), all band/flag combinations, lengths around the 16-cell boundaries."""
    rng = np.random.default_rng(seed)
    lens = [L for L in (1, 2, 5, 15, 16, 17, 31, 32, 33, 48, 100, 200, 413, 700, 1500) if L <= max_len]
    out = []
    for _ in range(n):
        L = int(rng.choice(lens))
        t = rng.integers(0, 4, L).astype(np.uint8)
        rate = float(rng.choice([0.0, 0.05, 0.15, 0.3, 0.6]))
        u = rng.random(L)
        keep = u >= rate * 0.3
        qv = np.where((u >= rate * 0.6) & (u < rate), (t + 1 + rng.integers(0, 3, L)) % 4, t).astype(np.uint8)[keep]
        ins = rng.random(qv.shape[0]) < rate * 0.3
        qv = np.insert(qv, np.nonzero(ins)[0], rng.integers(0, 4, int(ins.sum())).astype(np.uint8))
        if qv.shape[0] == 0:
            qv = np.zeros(1, dtype=np.uint8)
        if rng.random() < 0.1: t[rng.integers(0, len(t))] = 4
        if rng.random() < 0.1: qv[rng.integers(0, len(qv))] = 4
        if rng.random() < 0.2:
            qv = np.concatenate([qv, rng.integers(0, 4, int(rng.integers(1, 600))).astype(np.uint8)])
        out.append((np.ascontiguousarray(qv), np.ascontiguousarray(t), int(rng.choice([-1, 3001, 751, 100, 20, 5, 0])),
                    int(rng.choice([400, 200, 50, -1])), int(rng.choice([-1, 5, 50])), int(rng.choice(FLAGS))))
    return out


def wide_cases(n, seed=11):
    """long pairs with wide bands (the extension / long-join calls of mm_align1): 2-6 kb, w in {3001, 751, -1}"""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        L = int(rng.choice([1100, 2500, 4000, 6000]))
        t = rng.integers(0, 4, L).astype(np.uint8)
        rate = float(rng.choice([0.05, 0.15, 0.3]))
        u = rng.random(L)
        keep = u >= rate * 0.3
        qv = np.where((u >= rate * 0.6) & (u < rate), (t + 1 + rng.integers(0, 3, L)) % 4, t).astype(np.uint8)[keep]
        ins = rng.random(qv.shape[0]) < rate * 0.3
        qv = np.insert(qv, np.nonzero(ins)[0], rng.integers(0, 4, int(ins.sum())).astype(np.uint8))
        if rng.random() < 0.5:                  # unrelated tail: the extension z-drops inside it
            cut = int(rng.integers(L // 4, L))
            qv = np.concatenate([qv[:cut], rng.integers(0, 4, int(rng.integers(500, 3000))).astype(np.uint8)])
        out.append((np.ascontiguousarray(qv), np.ascontiguousarray(t), int(rng.choice([3001, 751, -1, 1500])),
                    int(rng.choice([400, 200])), int(rng.choice([-1, 5])), int(rng.choice([0x40 | 0x02 | 0x80, 0x40, 0x08, 0]))))
    return out
