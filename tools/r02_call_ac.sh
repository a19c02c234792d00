#!/bin/bash
# round-2 call AC: ramp segments in generated kernels — GPU parity, then A/B on the bench line's graph-compiler FMVoice entry
# (blocks 3..12 of the schedule: ramp blocks from block 10) and on a ramp-only window
mkdir -p gpurun_out
python -m pytest tests/test_specialize.py tests/test_variants_gpu.py tests/test_dsl_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02ac_tests.log
for rs in 0 1; do
  OSCEN_B200_JIT_RAMP_SEGMENTS=$rs python bench.py --steps 20 --warmup 5 --no-cpu-baseline --other-steps 16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for o in d['other_configs']:
    if o['config'].startswith('graph compiler: FMVoice'): print('[RAMP_SEGMENTS=$rs]', '%.4g v-s/s' % o['value'], '%.4f ms mean over blocks 3..18' % o['kernel_ms'])"
done 2>&1 | tee gpurun_out/r02ac_ab_ramp_segments.log
