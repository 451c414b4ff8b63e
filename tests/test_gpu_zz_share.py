"""Two contexts on one device mapping against ONE copy of the index (ndg_index_share): same hits as the owner, concurrently."""
import os
import threading
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_shared_index_two_contexts_same_bytes(engine_mod, reads_sub, golden_dir):
    import nextdenovo_b200 as nd
    io, mo = nd.IdxOpt(), nd.MapOpt()
    nd.set_opt("ava-ont", io, mo)
    mo.step, mo.minlen = 1, 500
    a, b = engine_mod.Engine(0), engine_mod.Engine(0)
    try:
        for e in (a, b):
            e.set_options(io, mo)
        a.build_index(reads_sub)
        b.share_index_of(a)
        out = {}

        def run(e, key):
            st = np.zeros(2, dtype=np.uint32)
            out[key] = e.map_batch_ovl(reads_sub, st)[0]
        ts = [threading.Thread(target=run, args=(a, "a")), threading.Thread(target=run, args=(b, "b"))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        want = open(os.path.join(golden_dir, "S1_sub120_ava-ont.ovl"), "rb").read()
        assert out["a"] == want and out["b"] == want
    finally:
        b.close()
        a.close()
