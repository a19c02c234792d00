#!/bin/bash
# round-2 call H: full GPU suite on the final library; A/B of the zero-copy mix-bus mirror (end-to-end path)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02h_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02h_tests.log
for round in 1 2; do for lib in oscen_b200/lib/liboscen_b200_prev.so oscen_b200/lib/liboscen_b200.so; do
  OSCEN_B200_LIB=$PWD/$lib python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib'.split('/')[-1], '20 steps: value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'e2e ms/step %.4f' % d['e2e']['ms_per_step'], 'median %.4f' % d['e2e']['step_ms']['median'])"
  OSCEN_B200_LIB=$PWD/$lib python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib'.split('/')[-1], '200 steps: value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'e2e ms/step %.4f' % d['e2e']['ms_per_step'], 'median %.4f' % d['e2e']['step_ms']['median'])"
done; done 2>&1 | tee gpurun_out/r02h_ab_e2e.log
timeout 600 python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02h_bench_reference.json 2> gpurun_out/r02h_bench_reference.err; echo "reference arm rc=$?"; tail -c 600 gpurun_out/r02h_bench_reference.json
