#!/bin/bash
# round-2 call X: pipelined instance_mix_node_kernel — node-plan parity tests, node plan timing at 65 536 / 1 M voices, per-kernel times
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "node or plan or dsl_fm or variants" 2>&1 | tail -3 | tee gpurun_out/r02x_tests.log
for V in 65536 1048576; do python tools/bench_dsl.py 10 $V voice_fm3 2>/dev/null | grep "^{" | grep "node plan"; done | tee gpurun_out/r02x_node_plan.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:instance_mix -c 6 --csv --log-file gpurun_out/r02x_instance_mix_launches.csv python tools/bench_dsl.py 2 1048576 voice_fm3 > /dev/null 2>&1
grep instance_mix gpurun_out/r02x_instance_mix_launches.csv | awk -F'","' '{print $5, $NF}' | tail -4
