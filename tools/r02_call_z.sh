#!/bin/bash
# round-2 call Z: FINAL library (after specialise, array-state-in-shared-memory, segmented Synth kernel, SatGraph unroll 2, pipelined
# instance_mix) — full GPU suite, smoke, default bench + the driver's 20-step window + reference arm, ncu recapture of the kernels
# that changed since call S, graph / DSL benches
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02z_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02z_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02z_bench_20.json 2>/dev/null; echo "bench20 rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02z_bench_reference.json 2>/dev/null; echo "reference rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs > /dev/null 2>&1
P="ncu --set full --clock-control none --import-source on"
cap() { local name=$1 rx=$2 skip=$3 title=$4 warps=$5; shift 5
  timeout 400 $P -k regex:$rx -s $skip -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  bash tools/ncu_pack.sh $name "$title" $warps 256; rm -f gpurun_out/$name.source.csv.gz; }
cap r02z_sat4 sat_kernel 3 "sat_kernel<4> (plain loop unrolled by 2), 262 144 voices" 8192 python tools/bench_graphs.py 2 "sat_4x"
cap r02z_synth synth_kernel 3 "synth_kernel<true> (segmented), 1 048 576 voices" 32768 python tools/bench_graphs.py 2 "synth_adsr"
cap r02z_ep_generated oscen_b200_generated 3 "ElectricPiano compiled from its graph text: generated kernel, array-node state in shared memory, 16 384 x 16 voices" 8192 python tools/bench_epiano_dsl.py 16384 2
python tools/bench_graphs.py 8 > gpurun_out/r02z_bench_graphs.jsonl 2>/dev/null
for V in 65536 1048576; do python tools/bench_dsl.py 10 $V 2>/dev/null; done > gpurun_out/r02z_bench_dsl_four_ways.jsonl
ls gpurun_out | grep r02z | wc -l; du -sh gpurun_out
