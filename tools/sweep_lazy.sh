#!/bin/bash
# sweep of the lazy-mapping knobs (first chunk rows, later chunk rows, SMs left free)
for cfg in "40000 200000 20" "40000 40000 20" "20000 100000 36" "40000 1000000 0" "40000 100000 48" "1100000 1100000 0"; do
  set -- $cfg
  v=$(GANSPACE_B200_LAZY_FIRST=$1 GANSPACE_B200_LAZY_LATER=$2 GANSPACE_B200_LAZY_FREE_SMS=$3 timeout 120 python bench.py --steps 4 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 | grep -o '"value": [0-9.]*' | head -1)
  echo "first=$1 later=$2 free=$3 -> $v"
done
