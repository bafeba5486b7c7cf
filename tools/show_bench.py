import json, sys
for f in sys.argv[1:]:
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); r = d['roofline']
            print(f.split('/')[-1], 'ms/step', round(d['ms_per_step'], 2), 'Mpts/s', round(d['value'] / 1e6, 2), 'e2e', round(d['e2e']['value'] / 1e6, 2) if d.get('e2e') else None,
                  'feat ms', round(r['kernel_ms'], 2), 'frac', round(r['frac'], 3), {k: round(v, 2) for k, v in r['stage_ms'].items()}, 'kp', d['keypoints'], 'fast', r['fast_math_selftest_passed'])
        elif l.strip():
            print(l[:300].rstrip())
