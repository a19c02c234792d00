#!/bin/bash
# round-2 call L: register-resident FIR histories in the generated kernels' cross-rate edges and the clip family: parity + A/B
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02l_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02l_tests.log
for round in 1 2; do for lib in liboscen_b200_prev2.so liboscen_b200.so; do
  OSCEN_B200_LIB=$PWD/oscen_b200/lib/$lib python tools/bench_graphs.py 8 "clip_os sinc" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lib', d['graph'][:58], '%.4g v-s/s %.3f ms' % (d['voice_samples_per_s'], d['kernel_ms']))"
  for V in 262144 1048576; do OSCEN_BENCH_WAYS=gen OSCEN_B200_LIB=$PWD/oscen_b200/lib/$lib python tools/bench_dsl.py 8 $V sat_4x 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lib', d['voices'], d['graph'], d['way'][:22], '%.4g v-s/s %.3f ms' % (d['voice_samples_per_s'], d['kernel_ms']))"; done
done; done 2>&1 | tee gpurun_out/r02l_ab.log
