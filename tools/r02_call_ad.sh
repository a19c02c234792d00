#!/bin/bash
# round-2 call AD: the library as committed at the end of the round — full GPU suite, smoke, default bench, the driver's 20-step
# window, the reference arm, launch list of the bench under ncu, graph / DSL benches
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02ad_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02ad_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02ad_bench.json 2> gpurun_out/r02ad_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02ad_bench_20.json 2>/dev/null; echo "bench20 rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02ad_bench_reference.json 2>/dev/null; echo "reference rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02ad_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs > /dev/null 2>&1
python tools/bench_graphs.py 8 > gpurun_out/r02ad_bench_graphs.jsonl 2>/dev/null
for V in 65536 1048576; do python tools/bench_dsl.py 10 $V 2>/dev/null; done > gpurun_out/r02ad_bench_dsl_four_ways.jsonl
python tools/bench_epiano_dsl.py 16384 5 2>/dev/null | grep "^{" > gpurun_out/r02ad_bench_epiano_dsl.jsonl
ls gpurun_out | grep r02ad | wc -l
