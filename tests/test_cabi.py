"""The C-ABI library loads and exports every symbol include/ndgpu.h declares; host-side entry points
(no GPU needed) behave; compute entry points fail loudly without a device.  CPU only."""
import ctypes as C
import os
import re
import numpy as np
import nextdenovo_b200 as nd


def test_exports_match_header(engine_mod):
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "ndgpu.h")).read()
    declared = set(re.findall(r"\b(ndg_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"ndg_ctx", "ndg_last_error"} - {"ndg_last_error"}
    L = engine_mod.load_library()
    assert len(declared) >= 25
    for s in sorted(declared):
        assert hasattr(L, s), s
    assert set(engine_mod.EXPORTS) <= declared
    assert b"sm_100a" in L.ndg_version()


def test_no_gpu_fails_loudly(engine_mod):
    import torch
    if torch.cuda.is_available():
        return
    try:
        engine_mod.Engine(0)
    except engine_mod.NdgError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("Engine() must raise without a GPU")


def test_host_encoder_matches_reference_bytes(engine_mod, golden_dir):
    L = engine_mod.load_library()
    buf = open(os.path.join(golden_dir, "S3_sub120_ava-pb.ovl"), "rb").read()
    rec = nd.decode_ovl8(buf)
    st = np.zeros(2, dtype=np.uint32)
    tmp = np.zeros(64, dtype=np.uint8)
    out = bytearray()
    for r in rec[:3000]:
        n = L.ndg_ovl_encode8(tmp.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p), int(r[1]), *[int(x) for x in (r[0], r[2], r[3], r[4], r[5], r[6], r[7])])
        out += tmp[:n].tobytes()
    assert bytes(out) == buf[:len(out)]
    # big values take the 5-byte path (>= 2^28) exactly like lib/ovl.c:133-143
    st[:] = 0
    n = L.ndg_ovl_encode8(tmp.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p), 1, 4294967295, 268435456, 268435457, 100000000, 5, 4000000000, 4294967295)
    assert tmp[:n].tobytes() == nd.encode_ovl8_py([(4294967295, 1, 268435456, 268435457, 100000000, 5, 4000000000, 4294967295)])


def test_emit_step1_filter(engine_mod):
    """self hits dropped, qe-qs >= minlen kept (ref: minimap2/map.c:1297-1304)"""
    L = engine_mod.load_library()
    hits = np.zeros(4, dtype=engine_mod.HIT_DTYPE)
    hits["qidx"] = [0, 0, 1, 1]; hits["rid"] = [1, 0, 0, 1]
    hits["qs"] = [10, 0, 5, 0]; hits["qe"] = [700, 900, 504, 800]; hits["rs"] = [1, 0, 9, 3]; hits["re"] = [650, 900, 600, 820]
    hits["mlen"] = [300, 900, 100, 400]; hits["rev"] = [0, 0, 1, 0]
    ids = np.array([5, 6], dtype=np.uint32); keys = nd.name_keys(ids); lens = np.array([1000, 1000], dtype=np.uint32)
    st = np.zeros(2, dtype=np.uint32); dst = np.zeros(256, dtype=np.uint8)
    nrec, sq = C.c_uint64(0), C.c_uint64(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    nb = L.ndg_ovl_emit_step1(p(hits), 4, p(ids), p(keys), p(lens), p(ids), p(keys), p(lens), 500, 0, 5000, 500, p(st), p(dst), 256, C.byref(nrec), C.byref(sq))
    rec = nd.decode_ovl8(dst[:nb].tobytes())
    assert nrec.value == 1 and sq.value == 690
    assert [int(x) for x in rec[0]] == [5, 0, 10, 700, 6, 1, 650, 300]


def test_rebase_first_record_stitches_independent_batches(engine_mod, golden_dir):
    """ndg_ovl_rebase_first: a batch encoded from state (0,0) is rebased onto the stream's true previous (qname, tname); splitting the
    golden G1 stream at several records and stitching gives back the reference's bytes (ref: lib/ovl.c:109-150)."""
    import ctypes as C
    import os
    import numpy as np
    import nextdenovo_b200 as nd
    L = engine_mod.load_library()
    L.ndg_ovl_rebase_first.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    buf = open(os.path.join(golden_dir, "G1_ava-ont.ovl"), "rb").read()
    rec = nd.decode_ovl8(buf)
    for cut in (1, 777, 65000, 130000):
        tup = [tuple(int(x) for x in r) for r in rec[:cut + 500]]
        first, second0, want = nd.encode_ovl8_py(tup[:cut]), nd.encode_ovl8_py(tup[cut:]), nd.encode_ovl8_py(tup)
        prev = np.array([rec[cut - 1][0], rec[cut - 1][4]], dtype=np.uint32)
        dst = np.zeros(48, dtype=np.uint8)
        skip = C.c_uint32(0)
        sb = np.frombuffer(bytes(second0), dtype=np.uint8)
        m = L.ndg_ovl_rebase_first(sb.ctypes.data, len(second0), prev.ctypes.data, dst.ctypes.data, C.byref(skip))
        assert m > 0
        assert bytes(first) + bytes(dst[:m]) + bytes(second0[skip.value:]) == bytes(want) == buf[:len(want)]
    assert L.ndg_ovl_rebase_first(sb.ctypes.data, 3, prev.ctypes.data, dst.ctypes.data, C.byref(skip)) < 0      # truncated
