#!/bin/bash
# One GPU-box session: quick kernel sanity, the GPU test-suite, a short bench.  Every step has its
# own timeout so that a hung kernel cannot eat the box's time limit.  Logs go to gpurun_out/.
# usage: scripts/gpu_check.sh <tag> [pytest args...]
tag=${1:-run}; shift
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
echo "== quick" ; timeout 300 python -m pytest tests/test_gpu_posdef.py -m gpu -x -q -s -k "microbench or dense_front_posdef" > $out/${tag}_quick.log 2>&1; echo "quick rc=$?"; tail -5 $out/${tag}_quick.log
echo "== suite" ; timeout 1500 python -m pytest tests -m gpu -q -s "$@" > $out/${tag}_suite.log 2>&1; echo "suite rc=$?"; tail -25 $out/${tag}_suite.log
echo "== bench" ; timeout 600 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cat $out/${tag}_bench.json | cut -c1-1800
