#!/bin/bash
# A/B of a kernel-variant build against the product build: correctness (dense + tree tests), the
# tile micro-benchmark and a short bench.  usage: scripts/ab_variant.sh <variant-tag|default> ...
for v in "$@"; do
  if [ "$v" = default ]; then unset SYLVER_B200_LIB; else export SYLVER_B200_LIB=$PWD/sylver_b200/libsylver_b200_$v.so; fi
  echo "===== variant $v"
  timeout 300 python -m pytest tests/test_gpu_posdef.py tests/test_gpu_indef.py -m gpu -x -q 2>&1 | tail -2
  for nk in "8192 256" "8192 128" "8192 2048" "2048 256"; do timeout 120 python scripts/prof_tile.py $nk 3 | tail -1; done
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "roofline", d["roofline"]["frac"], {k:v["ms"] for k,v in d["kernel_breakdown_ms"].items()})
PY
done
