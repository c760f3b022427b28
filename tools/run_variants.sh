#!/bin/bash
# Developer tool: time every variant under tools/variants/ on the GPU box (uniform problems, steady state).
for d in tools/variants/*/; do
  n=$(basename $d)
  r=$(SA_LIB=$d/libsegaligner_b200.so python tools/quick_bench.py ${1:-200} ${2:-100} 0 1 2>&1 | grep rep2 | sed 's/.*fill/fill/')
  echo "$n: $r"
done
