#!/bin/bash
# round-2 call K: A/B of the rotated isum rows (bank conflicts of the mix-partial fold) on every fused kernel + full suite
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02k_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02k_tests.log
for round in 1 2; do for lib in liboscen_b200_prev.so liboscen_b200.so; do
  OSCEN_B200_LIB=$PWD/oscen_b200/lib/$lib python tools/bench_graphs.py 8 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lib', d['graph'][:58], '%.4g v-s/s %.3f ms' % (d['voice_samples_per_s'], d['kernel_ms']))"
  OSCEN_B200_LIB=$PWD/oscen_b200/lib/$lib python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'FM 65536 20 steps: value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'])"
  OSCEN_B200_LIB=$PWD/oscen_b200/lib/$lib python bench.py --steps 20 --warmup 25 --voices 1048576 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'FM 1M: value %.4g' % d['value'], 'ms %.4f' % d['ms_per_step'])"
  OSCEN_B200_LIB=$PWD/oscen_b200/lib/$lib python bench.py --plan nodes --steps 20 --warmup 25 --voices 1048576 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'node plan 1M: value %.4g' % d['value'], 'ms %.4f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'])"
done; done 2>&1 | tee gpurun_out/r02k_ab.log
