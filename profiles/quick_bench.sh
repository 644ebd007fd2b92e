#!/bin/bash
# quick kernel-time check of the main workloads (no CPU leg)
for wl in headline c2; do
  python bench.py --steps 100 --warmup 5 --workload $wl --no-cpu --no-cloud 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$wl', 'ms/step %.4f' % d['ms_per_step'], 'kernel_ms %.4f' % d['roofline']['kernel_ms'], 'frac %.3f' % d['roofline']['frac'], 'Mpix/s %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'])
"
done
