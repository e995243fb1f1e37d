#!/bin/bash
# quick knob sweep on the GPU box: bench with a few env settings (each line = one bench JSON, trimmed)
for cfg in "${@:-GGB_GEMV_CTAS_PER_SM=1}"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 128 --warmup 8 --no-cpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('tok/s %.1f  gemv GB/s %.0f (frac %.3f) e2e %.1f' % (j['value'], j['roofline']['achieved'], j['roofline']['frac'], j['e2e']['value']))
"
done
