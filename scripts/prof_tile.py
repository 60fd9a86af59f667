"""Micro-benchmark of the DMMA tile kernel on one front (contribution mode): n x n block, depth k.
usage: python scripts/prof_tile.py n k [iters]   (run under ncu for a profile)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sylver_b200 as sb

L = sb.lib()
sb.require_gpu()
n, k = int(sys.argv[1]), int(sys.argv[2])
it = int(sys.argv[3]) if len(sys.argv) > 3 else 3
print("peak", L.sylver_b200_bench_dmma(0, 0, 0, 3))
print(f"tile n={n} k={k}", L.sylver_b200_bench_dmma(1, n, k, it))
