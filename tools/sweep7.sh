#!/bin/bash
out=gpurun_out/sweep_knn_ball.txt; : > $out
for f in 9.55 7.6 6.4 5.1; do
  for wl in scene10m view1m; do
    steps=3; [ $wl != scene10m ] && steps=10
    line=$(KPL_KNN_BALL=$f python bench.py --no-cpu --no-extras --workload $wl --steps $steps --warmup 2 2>/dev/null | tail -1)
    python - "$f" "$wl" "$line" >> $out <<'PY'
import json,sys
d=json.loads(sys.argv[3]); r=d['roofline']
print("ball %-6s %-9s normals %7.3f ms  step %8.3f ms  digest %s" % (sys.argv[1], sys.argv[2], r['stage_ms']['normals_ms'], d['ms_per_step'], d.get('digest_match')))
PY
  done
done
cat $out
