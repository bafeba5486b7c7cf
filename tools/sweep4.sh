#!/bin/bash
# experiments build (-DKPL_EXPERIMENTS): query-order parameters of the feature kernel on the three single-GPU workloads
out=gpurun_out/sweep_qorder.txt; : > $out
run() { # label, env...
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
run base X=1
run E1 KPL_GROUP_E=1
run E3 KPL_GROUP_E=3
run sub1 KPL_SUB=1
run sub3 KPL_SUB=3
run jump2 KPL_JUMP=2
run jump2E3 KPL_JUMP=2 KPL_GROUP_E=3
cat $out
