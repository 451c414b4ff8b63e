"""Pins oracle/nd_ksw.c (scalar restatement of ksw_extd2_sse) against the unmodified reference
function in oracle/_ref/libminimap2_ref.so.  CPU only."""
import pytest
from ksw_cases import simple_mat, oracle_lib, oracle_call, ref_lib, ref_call, random_cases


@pytest.mark.parametrize("params", [(4, 2, 24, 1), (6, 2, 26, 1), (4, 2, 4, 2), (24, 1, 4, 2)])
def test_ksw_oracle_equals_reference(params):
    R = ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built")
    L = oracle_lib()
    mat = simple_mat()
    q, e, q2, e2 = params
    bad = 0
    for i, (qs, ts, w, zd, eb, flag) in enumerate(random_cases(400, seed=11 + q, max_len=700)):
        a = ref_call(R, qs, ts, mat, q, e, q2, e2, w, zd, eb, flag)
        b = oracle_call(L, qs, ts, mat, q, e, q2, e2, w, zd, eb, flag)
        bad += a != b
        assert a == b, "case %d qlen=%d tlen=%d w=%d flag=%#x" % (i, len(qs), len(ts), w, flag)
