#!/usr/bin/env python
"""regime B on sub120 (for ncu: the DP launches of the real -c workload, exact-maximum extensions included): python tools/ksw_regimeb.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nextdenovo_b200 as nd
from nextdenovo_b200 import engine
from nextdenovo_b200.options import MM_F_CIGAR
reads = nd.read_2bit(os.path.join(ROOT, "tests", "golden", "sub120.2bit"), max_reads=40)
io, mo = nd.IdxOpt(), nd.MapOpt()
nd.set_opt("ava-ont", io, mo)
mo.step, mo.minlen = 1, 500
mo.flag |= MM_F_CIGAR
eng = engine.Engine(0)
eng.set_threads(0)
engine.run_step1(eng, reads, reads, io, mo)
if os.environ.get("ND_NCU_RANGE"):
    torch.cuda.cudart().cudaProfilerStart()
eng.prof_reset()
ovl, st = engine.run_step1(eng, reads, reads, io, mo)
if os.environ.get("ND_NCU_RANGE"):
    torch.cuda.cudart().cudaProfilerStop()
print(st, eng.aln_stats())
