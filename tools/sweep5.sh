#!/bin/bash
out=gpurun_out/sweep_groups.txt; : > $out
run() {
  for wl in scene10m view1m cheff001; do
    steps=3; [ $wl != scene10m ] && steps=10
    line=$(env "${@:2}" python bench.py --no-cpu --no-extras --workload $wl --steps $steps --warmup 2 2>/dev/null | tail -1)
    python - "$1" "$wl" "$line" >> $out <<'PY'
import json,sys
d=json.loads(sys.argv[3]); r=d['roofline']
print("%-22s %-9s feat %8.3f ms  step %8.3f ms  accept %.3f  grid %.2f  digest %s" % (sys.argv[1], sys.argv[2], r['kernel_ms'], d['ms_per_step'], r['acceptance'], r['stage_ms']['grid_ms'], d.get('digest_match')))
PY
  done
}
run E3 KPL_GROUP_E=3
run E4 KPL_GROUP_E=4
run E5 KPL_GROUP_E=5
run E4jump2 KPL_GROUP_E=4 KPL_JUMP=2
cat $out
