#!/bin/bash
# time one workload with every experimental build under ndsimulator_b200/csrc/exp/
wl=${1:-c4}
for so in ndsimulator_b200/csrc/exp/libndsb200_*.so; do
  tag=$(basename $so .so | sed 's/libndsb200_//')
  NDSB200_LIB=$PWD/$so python bench.py --workload $wl --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-workloads 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('$tag', '$wl', '%.4g' % d['value'], 'ms', round(d['ms_per_step'], 4), d['config'].get('kernel'), 'frac %.3f' % d['roofline']['frac'])"
done
