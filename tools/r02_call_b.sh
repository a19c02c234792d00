#!/bin/bash
# round-2 call B: A/B of the register-ring sat kernel and the 4-lane epiano kernel against the previous build, parity tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "sat or piano or epiano or twin or synth or clip or golden or fullsize or config" > gpurun_out/r02b_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02b_tests.log
for c in "sat_4x" "sat_1x" "x 16 voices" "x 8 voices" "device LCG" "synth_adsr"; do bash tools/ab_graphs.sh "$c" 8; done 2>&1 | tee gpurun_out/r02b_ab.log
P="ncu --set full --clock-control none --import-source on"
timeout 300 $P -k regex:sat_kernel -s 3 -c 1 -f -o gpurun_out/r02b_sat4 python tools/bench_graphs.py 2 "sat_4x" > gpurun_out/r02b_sat4.log 2>&1
bash tools/ncu_pack.sh r02b_sat4 "sat_kernel<4>, register rings" 8192 256
timeout 400 $P -k regex:epiano -s 3 -c 1 -f -o gpurun_out/r02b_ep python tools/bench_graphs.py 2 "x 16 voices" > gpurun_out/r02b_ep.log 2>&1
bash tools/ncu_pack.sh r02b_ep "epiano_kernel, 4 lanes per voice" 262144 256
