#!/bin/bash
# round-2 call T: generator knobs on the segmented kernels (A/B only): CTAs-per-SM target, uniform inputs in the constant bank
mkdir -p gpurun_out
for mb in 6 8 10; do for V in 65536 1048576; do
  OSCEN_B200_JIT_MIN_BLOCKS=$mb OSCEN_BENCH_WAYS=gen python tools/bench_dsl.py 10 $V voice_fm3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('[MIN_BLOCKS=$mb]', d['voices'], d['graph'], '%.4g v-s/s %.3f ms' % (d['voice_samples_per_s'], d['kernel_ms']))"
done; done 2>&1 | tee gpurun_out/r02t_ab.log
for um in 0 0x0FFFFFFE; do for V in 65536 1048576; do
  OSCEN_B200_JIT_UNIFORM_MASK=$um OSCEN_BENCH_WAYS=gen python tools/bench_dsl.py 10 $V voice_fm3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('[UNIFORM_MASK=$um]', d['voices'], d['graph'], '%.4g v-s/s %.3f ms' % (d['voice_samples_per_s'], d['kernel_ms']))"
done; done 2>&1 | tee -a gpurun_out/r02t_ab.log
