#!/bin/bash
# round-2 call Q: segmented TwinPeaks / SatGraph kernels: parity + A/B against the previous build
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02q_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02q_tests.log
for round in 1 2; do for lib in liboscen_b200_prev.so liboscen_b200.so; do
  for c in "device LCG" "stream read from HBM" "sat_4x" "sat_1x"; do
  OSCEN_B200_LIB=$PWD/oscen_b200/lib/$lib python tools/bench_graphs.py 8 "$c" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lib', d['graph'][:58], '%.4g v-s/s %.3f ms' % (d['voice_samples_per_s'], d['kernel_ms']))"
  done
done; done 2>&1 | tee gpurun_out/r02q_ab.log
