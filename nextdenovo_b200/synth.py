"""Seeded synthetic read sets (benchmarks/tests): thin ctypes wrapper over csrc/nd_synth.c."""
import ctypes as C
import os
import numpy as np
from .formats import Reads, name_keys, x31_hash

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _L():
    global _lib
    if _lib is None:
        p = os.path.join(_HERE, "libndsynth.so")
        if not os.path.exists(p):
            raise RuntimeError("%s missing: run __graft_entry__.build()" % p)
        _lib = C.CDLL(p)
        _lib.nds_genome.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        _lib.nds_reads.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32,
                                   C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def genome(G, seed=20240923):
    g = np.zeros(G, dtype=np.uint8)
    _L().nds_genome(g.ctypes.data_as(C.c_void_p), G, seed)
    return g


def reads(g, n, L=20000, sub=0.05, ins=0.04, dele=0.06, seed=20240923, first=0, id0=None):
    """n reads of length L from genome g with the given per-base error rates; ids = first.. (or id0..)"""
    nw = (L + 15) // 16
    words = np.zeros(n * nw, dtype=np.uint32)
    q = lambda x: int(round(x * 65536))
    _L().nds_reads(g.ctypes.data_as(C.c_void_p), g.shape[0], first, n, L, q(sub), q(ins), q(dele), seed,
                   words.ctypes.data_as(C.c_void_p), None, None)
    ids = np.arange(first if id0 is None else id0, (first if id0 is None else id0) + n, dtype=np.uint32)
    off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(nw))
    return Reads(words, off, np.full(n, L, dtype=np.uint32), ids, name_keys(ids), x31_hash(ids))
