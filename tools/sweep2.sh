#!/bin/bash
for w in 4 2 1; do
  echo -n "feat_warps=$w "; KPL_FEAT_WARPS=$w python bench.py --workload view1m --no-cpu --steps 5 --warmup 2 2>/dev/null | python tools/show_bench.py /dev/stdin
done
for w in 4 1; do
  echo -n "10M feat_warps=$w "; KPL_FEAT_WARPS=$w python bench.py --no-cpu --steps 2 --warmup 2 2>/dev/null | python tools/show_bench.py /dev/stdin
done
