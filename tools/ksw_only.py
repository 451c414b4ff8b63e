#!/usr/bin/env python
"""ksw_extd2 kernel alone (for ncu): python tools/ksw_only.py [n_calls]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from nextdenovo_b200 import engine
eng = engine.Engine(0)
print(bench.ksw_bench(eng, n_calls=int(sys.argv[1]) if len(sys.argv) > 1 else 6000))
