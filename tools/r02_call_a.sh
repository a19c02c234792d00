#!/bin/bash
# round-2 call A: GPU suite, default bench line, launch list of the bench, full ncu captures of the config kernels
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs > /dev/null 2>&1
P="ncu --set full --clock-control none --import-source on"
timeout 300 $P -k regex:twin_peaks -s 3 -c 1 -f -o gpurun_out/r02a_tp python tools/bench_graphs.py 2 "device LCG" > gpurun_out/r02a_tp.log 2>&1
bash tools/ncu_pack.sh r02a_tp "r02a_tp" 32768 256
timeout 300 $P -k regex:sat_kernel -s 3 -c 1 -f -o gpurun_out/r02a_sat4 python tools/bench_graphs.py 2 "sat_4x" > gpurun_out/r02a_sat4.log 2>&1
bash tools/ncu_pack.sh r02a_sat4 "r02a_sat4" 8192 256
timeout 400 $P -k regex:epiano -s 3 -c 1 -f -o gpurun_out/r02a_ep python tools/bench_graphs.py 2 "x 16 voices" > gpurun_out/r02a_ep.log 2>&1
bash tools/ncu_pack.sh r02a_ep "r02a_ep" 65536 256
timeout 300 $P -k regex:clip_kernel -s 3 -c 1 -f -o gpurun_out/r02a_clipref python tools/bench_graphs.py 2 "clip_ref" > gpurun_out/r02a_clipref.log 2>&1
bash tools/ncu_pack.sh r02a_clipref "r02a_clipref" 32768 256
timeout 300 $P -k regex:fm_voice -s 7 -c 1 -f -o gpurun_out/r02a_fm00 python tools/prof_fm.py 65536 9 > gpurun_out/r02a_fm00.log 2>&1
bash tools/ncu_pack.sh r02a_fm00 "r02a_fm00" 2048 256
timeout 300 $P -k regex:fm_voice -s 14 -c 1 -f -o gpurun_out/r02a_fm11 python tools/prof_fm.py 65536 16 > gpurun_out/r02a_fm11.log 2>&1
bash tools/ncu_pack.sh r02a_fm11 "r02a_fm11" 2048 256
timeout 300 $P -k regex:fm_voice -s 7 -c 1 -f -o gpurun_out/r02a_fm00_1m python tools/prof_fm.py 1048576 9 > gpurun_out/r02a_fm00_1m.log 2>&1
bash tools/ncu_pack.sh r02a_fm00_1m "r02a_fm00_1m" 32768 256
OSCEN_BENCH_PLAN=nodes timeout 400 $P -k regex:"fmop_node|tpt_node|adsr_node" -s 30 -c 6 -f -o gpurun_out/r02a_nodes python bench.py --plan nodes --voices 1048576 --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r02a_nodes.log 2>&1
bash tools/ncu_pack.sh r02a_nodes "r02a_nodes"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 60 --csv --log-file gpurun_out/r02a_launches_nodes_1m.csv python bench.py --plan nodes --voices 1048576 --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > /dev/null 2>&1
ls -la gpurun_out | grep r02a; du -sh gpurun_out
