#!/bin/bash
# A/B two builds of the library in the same gpurun call (same box, same clocks):
#   tools/ab.sh [voices] [steps]
V=${1:-65536}; S=${2:-100}
for round in 1 2; do
  for lib in oscen_b200/lib/liboscen_b200_prev.so oscen_b200/lib/liboscen_b200.so; do
    OSCEN_B200_LIB=$PWD/$lib python bench.py --steps $S --warmup 20 --no-cpu-baseline --voices $V 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', d['config']['voices_per_gpu'], 'value %.4g' % d['value'], 'kernel_ms %.4f' % d['roofline']['kernel_ms'], 'e2e %.4g' % d['e2e']['value'])"
  done
done
