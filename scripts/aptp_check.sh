#!/bin/bash
# APTP path: parity tests, then timings of the dense front (config 2), 7-point LDL^T and KKT trees.
timeout 600 python -m pytest tests/test_gpu_indef.py tests/test_gpu_local_ranks.py tests/test_gpu_x_pivot_methods.py tests/test_gpu_x_random.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_seam.py -m gpu -x -q -s 2>&1 | grep -E "passed|failed|error|config 2|_[0-9]+:" | cut -c1-330
python scripts/run_case.py dense 8192 --ncol 2048 --reps 3 > /dev/null   # warm the box up
for e in "SYLVER_B200_APTP_LEGACY=1" "SYLVER_B200_APTP_S2=0" "SYLVER_B200_APTP_S2=1"; do
  echo "== $e"
  env $e python scripts/run_case.py dense 8192 --ncol 2048 --reps 4 | tail -1 | cut -c1-120
  env $e python scripts/run_case.py lap7 100 --indef --reps 3 --nosolve | grep '"rep": 2' | cut -c1-230
  env $e python scripts/run_case.py kkt 40 --indef --reps 3 --nosolve | grep '"rep": 2' | cut -c1-230
done
