#!/bin/bash
# round-2 call U: SatGraph plain-loop unroll factor A/B (instruction-cache pressure vs register renaming of the FIR shifts)
mkdir -p gpurun_out
for rep in 1 2; do for u in 1 2 3 4 8; do
  OSCEN_B200_LIB=$PWD/oscen_b200/lib/ab_sat_u$u.so python tools/bench_graphs.py 10 "sat_" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('[unroll $u]', d['graph'][:32], '%.4g v-s/s %.4f ms' % (d['voice_samples_per_s'], d['kernel_ms']))"
done; done 2>&1 | tee gpurun_out/r02u_ab.log
