#!/usr/bin/env python
"""The bench workload's first BLOCKS blocks on one engine, nothing else: the process ncu wraps for a capture of the FMVoice
graph kernel (block k of BASELINE configs[1] is the k-th launch matching `fm_voice`).

    ncu --set full --clock-control none --import-source on -k regex:fm_voice -s 7 -c 1 -o gpurun_out/prof python tools/prof_fm.py [voices] [blocks]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oscen_b200 import abi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 26
wl = bench.Workload(n, 0, n)
eng = abi.Engine(abi.GRAPH_FM_VOICE, n, bench.VPI, bench.SR, 0, abi.PLAN_FUSED, abi.KEEP_INSTANCE_OUT)
for b in range(blocks):
    wl.apply(eng, b, abi.EVENT_DTYPE, abi.EVENT_GATE)
    eng.l2_flush()
    eng.process_block(bench.FRAMES)
    print(b, eng.last_kernel_name(), flush=True)
eng.synchronize()
eng.close()
