#!/bin/bash
# round-2 call I: A/B ElectricPiano lane variants (4 lanes x 8 harmonics vs 2 x 16), select-form amplitude step, select-form tanhf (TwinPeaks)
mkdir -p gpurun_out
for lib in liboscen_b200_ep2x3.so liboscen_b200_ep2x4.so liboscen_b200.so; do
  OSCEN_B200_LIB=$PWD/oscen_b200/lib/$lib timeout 600 python -m pytest tests -m gpu -q -x -k "piano or epiano or twin or golden" 2>&1 | tail -1 | sed "s/^/$lib: /"
done 2>&1 | tee gpurun_out/r02i_tests.log
for round in 1 2; do for lib in liboscen_b200_prev.so liboscen_b200.so liboscen_b200_ep2x3.so liboscen_b200_ep2x4.so; do
  for c in "x 16 voices" "device LCG"; do
  OSCEN_B200_LIB=$PWD/oscen_b200/lib/$lib python tools/bench_graphs.py 8 "$c" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lib', d['graph'][:48], '%.4g v-s/s %.3f ms' % (d['voice_samples_per_s'], d['kernel_ms']))"
  done
done; done 2>&1 | tee gpurun_out/r02i_ab.log
