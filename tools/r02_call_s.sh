#!/bin/bash
# round-2 call S: FINAL library — full GPU suite, smoke, default bench + the driver's 20-step window + reference arm,
# every config kernel re-captured with ncu (profiles/r02_final_ncu_*.md, profiles/ncu_traffic.json), launch list of the bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02s_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02s_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02s_bench_20.json 2>/dev/null; echo "bench20 rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02s_bench_reference.json 2>/dev/null; echo "reference rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02s_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs > /dev/null 2>&1
P="ncu --set full --clock-control none --import-source on"
cap() { local name=$1 rx=$2 skip=$3 title=$4 warps=$5; shift 5
  timeout 400 $P -k regex:$rx -s $skip -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  bash tools/ncu_pack.sh $name "$title" $warps 256; rm -f gpurun_out/$name.source.csv.gz; }
cap r02s_fm_65536 fm_voice 25 "fm_voice_seg_kernel, block 25 of the bench workload, 65 536 voices" 2048 python tools/prof_fm.py 65536 27
cap r02s_fm_ramp_65536 fm_voice 14 "fm_voice_seg_kernel<ramp>, block 14 of the bench workload, 65 536 voices" 2048 python tools/prof_fm.py 65536 16
cap r02s_fm_plain_65536 fm_voice 7 "fm_voice_seg_kernel<ramp=0,fmod=0>, block 7 of the bench workload, 65 536 voices" 2048 python tools/prof_fm.py 65536 9
cap r02s_fm_1m fm_voice 25 "fm_voice_seg_kernel, block 25 of the bench workload, 1 048 576 voices" 32768 python tools/prof_fm.py 1048576 27
cap r02s_tp twin_peaks 3 "twin_peaks_kernel, 1 048 576 voices, device LCG input" 32768 python tools/bench_graphs.py 2 "device LCG"
cap r02s_sat4 sat_kernel 3 "sat_kernel<4>, 262 144 voices" 8192 python tools/bench_graphs.py 2 "sat_4x"
cap r02s_ep epiano 3 "epiano_kernel (two lanes x 16 harmonics per voice), 131 072 instances x 16 voices" 131072 python tools/bench_graphs.py 2 "x 16 voices"
python tools/bench_graphs.py 8 > gpurun_out/r02s_bench_graphs.jsonl 2>/dev/null
for V in 65536 1048576; do python tools/bench_dsl.py 10 $V 2>/dev/null; done > gpurun_out/r02s_bench_dsl_three_ways.jsonl
ls gpurun_out | grep r02s | wc -l; du -sh gpurun_out
