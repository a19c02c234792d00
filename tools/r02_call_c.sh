#!/bin/bash
# round-2 call C: full GPU suite on the new kernels; A/B sat / epiano against the previous build; node plan plain kernels A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02c_tests.log
for c in "sat_4x" "sat_1x" "x 16 voices" "device LCG"; do bash tools/ab_graphs.sh "$c" 8; done 2>&1 | tee gpurun_out/r02c_ab.log
for pl in 0 1; do for V in 1048576 65536; do
OSCEN_B200_NODE_PLAIN=$pl python bench.py --plan nodes --voices $V --steps 20 --warmup 25 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[NODE_PLAIN=$pl]', d['config']['voices_per_gpu'], 'value %.4g' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], 'frac', d['roofline'].get('frac'))"
done; done 2>&1 | tee -a gpurun_out/r02c_ab.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 160 -c 40 --csv --log-file gpurun_out/r02c_launches_nodes_1m.csv python bench.py --plan nodes --voices 1048576 --steps 2 --warmup 25 --no-cpu-baseline --no-other-configs > /dev/null 2>&1
P="ncu --set full --clock-control none --import-source on"
timeout 400 $P -k regex:epiano -s 3 -c 1 -f -o gpurun_out/r02c_ep python tools/bench_graphs.py 2 "x 16 voices" > gpurun_out/r02c_ep.log 2>&1
bash tools/ncu_pack.sh r02c_ep "epiano_kernel, 4 lanes per voice, 128 registers" 262144 256
timeout 300 $P -k regex:sat_kernel -s 3 -c 1 -f -o gpurun_out/r02c_sat4 python tools/bench_graphs.py 2 "sat_4x" > gpurun_out/r02c_sat4.log 2>&1
bash tools/ncu_pack.sh r02c_sat4 "sat_kernel<4>, register rings, unconditional body" 8192 256
timeout 400 $P -k regex:"node_plain" -s 12 -c 6 -f -o gpurun_out/r02c_nodes python bench.py --plan nodes --voices 1048576 --steps 2 --warmup 25 --no-cpu-baseline --no-other-configs > gpurun_out/r02c_nodes.log 2>&1
bash tools/ncu_pack.sh r02c_nodes "node plan plain kernels, 1M voices"
