#!/bin/bash
# grid resolution sweep with the final kernels (10 M-point scene)
for cpr in 3 4 5 6 8; do
  echo -n "cpr=$cpr "; python bench.py --no-cpu --steps 2 --warmup 2 --cpr $cpr 2>/dev/null | python tools/show_bench.py /dev/stdin
done
