#!/bin/bash
# round-2 call Y: tile_flush_mono with __restrict__ + four interleaved instance folds — A/B against the previous build on the
# FMVoice bench window, on the config graphs and on the node plan
mkdir -p gpurun_out
{ bash tools/ab.sh 65536 20 --steps20; bash tools/ab.sh 1048576 10; } 2>&1 | tee gpurun_out/r02y_ab_flush.log
bash tools/ab_graphs.sh "config" 8 2>&1 | tee -a gpurun_out/r02y_ab_flush.log
for lib in liboscen_b200_prev.so liboscen_b200.so; do OSCEN_B200_LIB=$PWD/oscen_b200/lib/$lib python tools/bench_dsl.py 10 1048576 voice_fm3 2>/dev/null | grep "^{" | grep "node plan" | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('[$lib] node plan 1M', '%.4f ms' % d['kernel_ms'])"; done 2>&1 | tee -a gpurun_out/r02y_ab_flush.log
