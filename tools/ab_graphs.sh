#!/bin/bash
# A/B of two library builds on tools/bench_graphs.py cases (same box): tools/ab_graphs.sh <case-substring> [steps]
for lib in oscen_b200/lib/liboscen_b200_prev.so oscen_b200/lib/liboscen_b200.so; do
  OSCEN_B200_LIB=$PWD/$lib python tools/bench_graphs.py ${2:-8} "$1" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lib'.split('/')[-1], d['graph'][:48], '%.4g v-s/s %.3f ms' % (d['voice_samples_per_s'], d['kernel_ms']))"
done
