#!/bin/bash
# A/B of APTP scheduling knobs on the LDL^T workloads (run after a warm-up)
timeout 600 python -m pytest tests/test_gpu_indef.py tests/test_gpu_posdef.py -m gpu -x -q 2>&1 | tail -2
python scripts/run_case.py dense 8192 --ncol 2048 --reps 3 > /dev/null
for e in "$@"; do
  echo "== $e"
  env $e python scripts/run_case.py dense 8192 --ncol 2048 --reps 4 | tail -1 | cut -c1-100
  env $e python scripts/run_case.py lap7 100 --indef --reps 3 --nosolve | grep '"rep": 2' | cut -c80-230
  env $e python scripts/run_case.py kkt 40 --reps 3 --nosolve | grep '"rep": 2' | cut -c80-230
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary | cut -c1-330
