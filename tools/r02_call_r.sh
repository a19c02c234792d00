#!/bin/bash
# round-2 call R: segmented general / legacy FMVoice kernels: parity + A/B (OSCEN_B200_FM_KERNEL=general|legacy forces them on the bench workload)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02r_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02r_tests.log
for round in 1 2; do for lib in liboscen_b200_prev.so liboscen_b200.so; do for k in general legacy; do for V in 65536 1048576; do
  OSCEN_B200_FM_KERNEL=$k OSCEN_B200_LIB=$PWD/oscen_b200/lib/$lib python bench.py --steps 20 --warmup 25 --voices $V --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', '$k', d['config']['voices_per_gpu'], 'value %.4g' % d['value'], 'ms %.4f' % d['ms_per_step'], [k['kernel'] for k in d['kernels']])"
done; done; done; done 2>&1 | tee gpurun_out/r02r_ab.log
