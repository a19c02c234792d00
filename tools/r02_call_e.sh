#!/bin/bash
# round-2 call E: full GPU suite, the default bench line, ncu captures of every config kernel from THIS binary (+ traffic table)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02e_tests.log
timeout 600 python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs > gpurun_out/r02e_bench_20.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs > /dev/null 2>&1
P="ncu --set full --clock-control none --import-source on"
cap() { # name kernel-regex skip title warps -- command...
  local name=$1 rx=$2 skip=$3 title=$4 warps=$5; shift 5
  timeout 400 $P -k regex:$rx -s $skip -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  bash tools/ncu_pack.sh $name "$title" $warps 256
}
cap r02e_fm_65536 fm_voice 25 "fm_voice_seg_kernel, block 25 of the bench workload, 65 536 voices" 2048 python tools/prof_fm.py 65536 27
cap r02e_fm_ramp_65536 fm_voice 14 "fm_voice_seg_kernel<ramp>, block 14 of the bench workload, 65 536 voices" 2048 python tools/prof_fm.py 65536 16
cap r02e_fm_1m fm_voice 25 "fm_voice_seg_kernel, block 25 of the bench workload, 1 048 576 voices" 32768 python tools/prof_fm.py 1048576 27
cap r02e_tp twin_peaks 3 "twin_peaks_kernel, 1 048 576 voices, device LCG input" 32768 python tools/bench_graphs.py 2 "device LCG"
cap r02e_sat4 sat_kernel 3 "sat_kernel<4>, 262 144 voices" 8192 python tools/bench_graphs.py 2 "sat_4x"
cap r02e_ep epiano 3 "epiano_kernel, 131 072 instances x 16 voices" 262144 python tools/bench_graphs.py 2 "x 16 voices"
ls gpurun_out | grep r02e | head -40; du -sh gpurun_out
