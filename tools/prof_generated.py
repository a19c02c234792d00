#!/usr/bin/env python
"""Workload for an `ncu --set full` capture of a GENERATED kernel: the FMVoice graph from its DSL text
(tests/graphs/voice_fm3.graph), 65 536 voices x 256 frames, steady-state sustain block.

    ncu --set full --clock-control none --import-source on -k regex:oscen_b200_generated_kernel -s 30 -c 1 \
        -o gpurun_out/prof_generated_fm python tools/prof_generated.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oscen_b200 import abi  # noqa: E402

n, vpi = 65536, 8
g = abi.Graph(open(os.path.join(ROOT, "tests", "graphs", "voice_fm3.graph")).read())
if len(sys.argv) > 1 and sys.argv[1] == "builtin":      # the hand-written kernel on the identical workload (-k regex:fm_voice_kernel)
    e = abi.Engine(abi.GRAPH_FM_VOICE, n, vpi, 48000.0, 0, abi.PLAN_FUSED, abi.KEEP_INSTANCE_OUT)
else:
    e = abi.Engine.from_graph(g, n, vpi, keep=abi.KEEP_INSTANCE_OUT, plan=abi.PLAN_FUSED)
rng = np.random.default_rng(1)
notes = 36 + rng.integers(0, 60, size=n)
e.set_value(0, (440.0 * np.exp2((notes - 69) / 12.0)).astype(np.float32))
ev = np.zeros(n, dtype=abi.EVENT_DTYPE)
ev["voice"] = np.arange(n); ev["frame_offset"] = rng.integers(0, 256, size=n); ev["input_id"] = abi.EVENT_GATE; ev["scalar"] = 0.8
e.push_events(ev)
for _ in range(34):
    e.l2_flush(); e.process_block(256)
e.synchronize()
