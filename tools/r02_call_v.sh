#!/bin/bash
# round-2 call V: oscen_b200_specialize (generated-kernel variant for bank-uniform inputs) — GPU parity + gain on every DSL
# graph; the segmented hand-written synth kernel and SatGraph unroll 2 (parity + timing)
mkdir -p gpurun_out
python -m pytest tests/test_specialize.py tests/test_variants_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02v_tests.log
for V in 65536 1048576; do python tools/bench_dsl.py 10 $V 2>/dev/null | grep "^{" ; done > gpurun_out/r02v_bench_dsl.jsonl
python tools/bench_graphs.py 10 "s" 2>/dev/null | grep "^{" > gpurun_out/r02v_bench_graphs.jsonl
python - <<'PY'
import json
for f in ("gpurun_out/r02v_bench_dsl.jsonl", "gpurun_out/r02v_bench_graphs.jsonl"):
    for l in open(f):
        d = json.loads(l); print(d["graph"][:40], d.get("way", "")[:60], d["voices"], "%.4g v-s/s %.4f ms" % (d["voice_samples_per_s"], d["kernel_ms"]), d.get("kernel", ""))
PY
