#!/bin/bash
# round-2 call P: segmented generated kernels for every graph: GPU suite + three-ways table + smoke
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02p_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02p_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for V in 65536 1048576; do python tools/bench_dsl.py 10 $V 2>/dev/null; done > gpurun_out/r02p_bench_dsl_three_ways.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/r02p_bench_dsl_three_ways.jsonl"):
    d=json.loads(l); print(d["voices"], d["graph"], d["way"][:24], "%.4g v-s/s %.3f ms" % (d["voice_samples_per_s"], d["kernel_ms"]))
PY
for sg in 0 1; do OSCEN_B200_JIT_SEGMENTS=$sg OSCEN_BENCH_WAYS=gen python tools/bench_dsl.py 10 1048576 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('[SEGMENTS=$sg]', d['voices'], d['graph'], '%.4g v-s/s %.3f ms' % (d['voice_samples_per_s'], d['kernel_ms']))"; done
