#!/bin/bash
out=gpurun_out/sweep_dense.txt; : > $out
run() {
  for wl in scene10m view1m; do
    steps=3; [ $wl != scene10m ] && steps=10
    line=$(env "${@:2}" python bench.py --no-cpu --no-extras --workload $wl --steps $steps --warmup 2 2>/dev/null | tail -1)
    python - "$1" "$wl" "$line" >> $out <<'PY'
import json,sys
d=json.loads(sys.argv[3]); r=d['roofline']
print("%-22s %-9s feat %8.3f ms  step %8.3f ms  digest %s" % (sys.argv[1], sys.argv[2], r['kernel_ms'], d['ms_per_step'], d.get('digest_match')))
PY
  done
}
for t in 32 31 30 29 28 26 24 20; do run dense$t KPL_DENSE_TILE=$t; done
cat $out
