#!/bin/bash
# round-2 call G: full GPU suite on the final library; A/B against the previous build: generated kernels (lean FmOperator),
# default bench (staging ring), node plan (instance fold loads, per-wave plain selection in ramp blocks)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02g_tests.log
for V in 65536 1048576; do for lib in oscen_b200/lib/liboscen_b200_prev.so oscen_b200/lib/liboscen_b200.so; do
  OSCEN_B200_LIB=$PWD/$lib python tools/bench_dsl.py 10 $V 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lib'.split('/')[-1], d['voices'], d['graph'], d['way'][:22], '%.4g v-s/s %.3f ms' % (d['voice_samples_per_s'], d['kernel_ms']))"
done; done 2>&1 | tee gpurun_out/r02g_ab_dsl.log
for lib in oscen_b200/lib/liboscen_b200_prev.so oscen_b200/lib/liboscen_b200.so; do
  OSCEN_B200_LIB=$PWD/$lib python bench.py --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib'.split('/')[-1], 'default bench: value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'e2e step max %.2f ms' % d['e2e']['step_ms']['max'])"
  OSCEN_B200_LIB=$PWD/$lib python bench.py --plan nodes --voices 1048576 --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib'.split('/')[-1], 'node plan 1M, blocks 5-24 (ramp blocks 10-18): value %.4g' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'])"
  OSCEN_B200_LIB=$PWD/$lib python bench.py --plan nodes --voices 1048576 --steps 20 --warmup 25 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib'.split('/')[-1], 'node plan 1M, plain blocks: value %.4g' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'])"
done 2>&1 | tee gpurun_out/r02g_ab_bench.log
timeout 600 python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?"
