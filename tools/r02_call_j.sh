#!/bin/bash
# round-2 call J: final library — full GPU suite, default bench + the driver's 20-step window, reference arm, ncu recapture of the two kernels
# that changed since call E (TwinPeaks: select-form tanhf; ElectricPiano: two lanes per voice)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02j_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02j_tests.log
timeout 600 python bench.py > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02j_bench_20.json 2>/dev/null; echo "bench20 rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02j_bench_reference.json 2>/dev/null; echo "reference rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
P="ncu --set full --clock-control none --import-source on"
cap() { local name=$1 rx=$2 skip=$3 title=$4 warps=$5; shift 5
  timeout 400 $P -k regex:$rx -s $skip -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  bash tools/ncu_pack.sh $name "$title" $warps 256; }
cap r02j_tp twin_peaks 3 "twin_peaks_kernel (select-form tanhf), 1 048 576 voices, device LCG input" 32768 python tools/bench_graphs.py 2 "device LCG"
cap r02j_ep epiano 3 "epiano_kernel (two lanes x 16 harmonics per voice), 131 072 instances x 16 voices" 131072 python tools/bench_graphs.py 2 "x 16 voices"
python tools/bench_graphs.py 8 > gpurun_out/r02j_bench_graphs.jsonl 2>/dev/null; python - <<'PY'
import json
for l in open("gpurun_out/r02j_bench_graphs.jsonl"):
    d=json.loads(l); print("%-72s %.4g v-s/s %.3f ms" % (d["graph"][:72], d["voice_samples_per_s"], d["kernel_ms"]))
PY
