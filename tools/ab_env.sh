#!/bin/bash
# A/B of one build under two environments in the same gpurun call (same box, same clocks):
#   tools/ab_env.sh "OSCEN_B200_FM_SKEW=0" "" [voices] [steps] [warmup]
A="$1"; B="$2"; V=${3:-65536}; S=${4:-20}; W=${5:-5}
for round in 1 2; do
  for e in "$A" "$B"; do
    env $e python bench.py --steps $S --warmup $W --no-cpu-baseline --no-other-configs --voices $V 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$e]', d['config']['voices_per_gpu'], 'value %.4g' % d['value'], 'kernel_ms %.4f' % d['roofline']['kernel_ms'], 'e2e %.4g' % d['e2e']['value'], [(k['kernel'][-16:], round(k['mean_ms'],4)) for k in d['kernels']])"
  done
done
