"""ctypes binding of the CPU oracle (oracle/nd_oracle.c).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by the product."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")


class Opt(C.Structure):
    _fields_ = [("k", C.c_int), ("w", C.c_int), ("hpc", C.c_int), ("flag_no_diag", C.c_int), ("flag_no_dual", C.c_int),
                ("min_cnt", C.c_int), ("min_chain_score", C.c_int), ("bw", C.c_int), ("max_gap", C.c_int),
                ("max_chain_skip", C.c_int), ("max_chain_iter", C.c_int), ("seed", C.c_int),
                ("mid_occ_frac", C.c_float), ("mid_occ", C.c_int), ("minlen", C.c_int)]


class Reg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("rid", "rev", "qs", "qe", "rs", "re", "mlen", "blen", "score", "cnt")] + [("hash", C.c_uint32)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, os.path.join(_HERE, "libnd_oracle.so"), os.path.join(_HERE, "nd_oracle")])


def lib():
    global _lib
    if _lib is None:
        p = os.path.join(_HERE, "libnd_oracle.so")
        if not os.path.exists(p):
            build()
        L = C.CDLL(p)
        L.ndo_sketch.restype = C.c_size_t
        L.ndo_sketch.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_size_t]
        L.ndo_opt_preset.argtypes = [C.POINTER(Opt), C.c_char_p]
        L.ndo_step1_files.restype = C.c_long
        L.ndo_step1_files.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(Opt), C.c_char_p, C.POINTER(C.c_uint64)]
        L.ndo_hash64.restype = C.c_uint64
        L.ndo_hash64.argtypes = [C.c_uint64, C.c_uint64]
        L.ndo_rs_sort_128x.argtypes = [C.c_void_p, C.c_void_p]
        L.ndo_ovl_encode8.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_uint32] * 7
        _lib = L
    return _lib


def make_opt(preset=None, k=None, w=None, **kw):
    o = Opt()
    lib().ndo_opt_preset(C.byref(o), None)
    if preset:
        lib().ndo_opt_preset(C.byref(o), preset.encode())
    if k is not None: o.k = k
    if w is not None: o.w = w
    for a, v in kw.items():
        setattr(o, a, v)
    return o


def sketch(seq: bytes, w, k, rid=0, hpc=0):
    out = np.zeros((len(seq) + 1, 2), dtype=np.uint64)
    n = lib().ndo_sketch(seq, len(seq), w, k, rid, hpc, out.ctypes.data_as(C.c_void_p), 0)
    return out[:n].copy()


def step1_files(target, query, opt, out_path):
    sq = C.c_uint64(0)
    n = lib().ndo_step1_files(target.encode(), query.encode(), C.byref(opt), out_path.encode() if out_path else None, C.byref(sq))
    if n < 0:
        raise RuntimeError("oracle failed on %s" % target)
    return n, sq.value


def rs_sort_128x(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    lib().ndo_rs_sort_128x(a.ctypes.data_as(C.c_void_p), C.c_void_p(a.ctypes.data + a.nbytes))
    return a


def ref_binary(name="minimap2-nd"):
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) else None
