#!/bin/bash
# round-2 call O: segmented frame loop in the generated kernels: DSL GPU tests + A/B (OSCEN_B200_JIT_SEGMENTS=0|1)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02o_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02o_tests.log
for round in 1 2; do for sg in 0 1; do for V in 65536 1048576; do for gname in voice_fm3 synth_adsr; do
  OSCEN_B200_JIT_SEGMENTS=$sg OSCEN_BENCH_WAYS=gen python tools/bench_dsl.py 10 $V $gname 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('[SEGMENTS=$sg]', d['voices'], d['graph'], d['way'][:22], '%.4g v-s/s %.3f ms' % (d['voice_samples_per_s'], d['kernel_ms']))"
done; done; done; done 2>&1 | tee gpurun_out/r02o_ab.log
