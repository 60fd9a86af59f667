"""Micro-benchmark of the DMMA tile kernel on one front (contribution mode): n x n block, depth k.
usage: python scripts/prof_tile.py n k [iters]   (run under ncu for a profile)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sylver_b200 as sb

L = sb.lib()
sb.require_gpu()
if sys.argv[1] == "potrf":
    print("potrf 128x128 + inverse: column-at-a-time %.1f us, blocked %.1f us" % (L.sylver_b200_bench_dmma(3, 0, 0, 5), L.sylver_b200_bench_dmma(4, 0, 0, 5)))
    sys.exit(0)
n, k = int(sys.argv[1]), int(sys.argv[2])
it = int(sys.argv[3]) if len(sys.argv) > 3 else 3
print("peak", L.sylver_b200_bench_dmma(0, 0, 0, 3))
print(f"tile n={n} k={k}", L.sylver_b200_bench_dmma(1, n, k, it))
