#!/bin/bash
# round-2 call W: segmented ElectricPiano kernel — GPU parity, then A/B against the same library built with
# -DOSCEN_B200_EP_SEGMENTS=0; ElectricPiano from its own graph text with the array nodes' state in shared memory
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "piano or epiano or midi or specialis or variants or config3" 2>&1 | tail -3 | tee gpurun_out/r02w_tests.log
for rep in 1 2; do for lib in ab_ep_noseg.so liboscen_b200.so; do
  OSCEN_B200_LIB=$PWD/oscen_b200/lib/$lib python tools/bench_graphs.py 6 "epiano" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('[$lib]', d['graph'][:52], '%.4g v-s/s %.4f ms' % (d['voice_samples_per_s'], d['kernel_ms']))"
done; done 2>&1 | tee gpurun_out/r02w_ab_epiano_segments.log
python tools/bench_epiano_dsl.py 16384 5 2>/dev/null | grep "^{" | tee gpurun_out/r02w_bench_epiano_dsl.jsonl
OSCEN_B200_JIT_ARRAY_SMEM=0 python tools/bench_epiano_dsl.py 16384 3 2>/dev/null | grep "^{" | head -1 | tee -a gpurun_out/r02w_bench_epiano_dsl.jsonl
