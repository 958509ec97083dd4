#!/bin/bash
# A/B of kernel variants on the c4 index with a short BED (one device-resident step after one warm-up):
# prints the per-phase device times of the last call. Usage: tools/ab_c4p.sh "VAR=1 VAR2=x" ...
for v in "$@"; do
  echo "== $v"
  env $v IMPGX_TRACE=1 timeout 300 python bench.py --workload c4p --profile --steps 1 --warmup 1 2>&1 | grep "\[impgx\]" | tail -1
done
