#!/bin/bash
# tuning sweep: grid resolution x group span on the 1M-point view (feature kernel ms from bench.py)
for cpr in 3 4 5 6; do for span in 0 1 2 3; do
  echo -n "cpr=$cpr span=$span "; KPL_FEAT_SPAN=$span python bench.py --workload view1m --no-cpu --steps 3 --warmup 2 --cpr $cpr 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('ms/step %.2f feat %.2f acc %.3f normals %.2f kp %d' % (d['ms_per_step'], r['kernel_ms'], r['acceptance'], r['stage_ms']['normals_ms'], d['keypoints']))
"; done; done
