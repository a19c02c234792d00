#!/usr/bin/env python
"""ElectricPiano compiled from its own graph! text (tests/graphs/electric_piano.graph) vs the hand-written epiano_kernel:
device time of one 256-frame block.  tools/bench_epiano_dsl.py [instances] [steps]   (OSCEN_B200_JIT_ARRAY_SMEM=0: HBM rows)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oscen_b200 import abi  # noqa: E402

n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
vpi, frames = 16, 256
n = n_inst * vpi
g = abi.Graph(open(os.path.join(ROOT, "tests", "graphs", "electric_piano.graph")).read())
for name, e in (("generated kernel", abi.Engine.from_graph(g, n, plan=abi.PLAN_FUSED, keep=abi.KEEP_INSTANCE_OUT)),
                ("hand-written epiano_kernel", abi.Engine(abi.GRAPH_EPIANO, n, vpi, 48000.0, 0, abi.PLAN_FUSED, abi.KEEP_INSTANCE_OUT))):
    rng = np.random.default_rng(1)
    notes = 36 + rng.integers(0, 60, size=n)
    fid = g.input_id("frequency") if "generated" in name and "frequency" in [g.input_info(i)["name"] for i in range(g.n_inputs)] else 0
    e.set_value(fid, (440.0 * np.exp2((notes - 69) / 12.0)).astype(np.float32))
    ev = np.zeros(n, dtype=abi.EVENT_DTYPE)
    ev["voice"] = np.arange(n); ev["frame_offset"] = rng.integers(0, frames, size=n); ev["input_id"] = abi.EVENT_GATE; ev["scalar"] = 0.8
    e.push_events(ev)
    e.enable_timing(True)
    for _ in range(2):
        e.l2_flush(); e.process_block(frames)
    e.synchronize()
    ms = []
    for _ in range(steps):
        e.l2_flush(); e.process_block(frames); ms.append(e.last_kernel_ms())
    k = float(np.mean(ms))
    print(json.dumps({"graph": "electric_piano.graph", "way": name, "array_smem": os.environ.get("OSCEN_B200_JIT_ARRAY_SMEM", "1"), "instances": n_inst,
                      "voices": n, "kernel_ms": k, "voice_samples_per_s": n * frames / (k * 1e-3), "kernel": e.last_kernel_name()}), flush=True)
    e.close()
