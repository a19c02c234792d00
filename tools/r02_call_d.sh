#!/bin/bash
# round-2 call D: kernel-variant parity tests, node plan with / without TMA tiles, launch list + ncu of the TMA kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_variants_gpu.py tests/test_parity_gpu.py tests/test_dsl_gpu.py -m gpu -x -q > gpurun_out/r02d_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02d_tests.log
for tm in 0 1; do for V in 1048576 65536; do
OSCEN_B200_NODE_TMA=$tm python bench.py --plan nodes --voices $V --steps 20 --warmup 25 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[NODE_TMA=$tm]', d['config']['voices_per_gpu'], 'value %.4g' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], 'frac', d['roofline'].get('frac'))"
done; done 2>&1 | tee gpurun_out/r02d_ab.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 60 --csv --log-file gpurun_out/r02d_launches_nodes_1m.csv python bench.py --plan nodes --voices 1048576 --steps 4 --warmup 25 --no-cpu-baseline --no-other-configs > /dev/null 2>&1
P="ncu --set full --clock-control none --import-source on"
timeout 400 $P -k regex:"node_tma|node_plain" -s 12 -c 6 -f -o gpurun_out/r02d_nodes python bench.py --plan nodes --voices 1048576 --steps 2 --warmup 25 --no-cpu-baseline --no-other-configs > gpurun_out/r02d_nodes.log 2>&1
bash tools/ncu_pack.sh r02d_nodes "node plan TMA kernels, 1M voices"
