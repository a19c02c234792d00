#!/usr/bin/env python
"""EXPERIMENT (round-2 groundwork): the generated FMVoice kernel with the inputs that are uniform over the bank read from
the kernel-argument constant bank (OSCEN_B200_JIT_UNIFORM_MASK) instead of per-thread registers — timing and a bit-identity
check against the default generated kernel and the hand-written one.   tools/ab_uniform.py [voices] [steps]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oscen_b200 import abi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
text = open(os.path.join(ROOT, "tests", "graphs", "voice_fm3.graph")).read()
rng = np.random.default_rng(1)
notes = 36 + rng.integers(0, 60, size=n)
freq = (440.0 * np.exp2((notes - 69) / 12.0)).astype(np.float32)
ev = np.zeros(n, dtype=abi.EVENT_DTYPE)
ev["voice"] = np.arange(n); ev["frame_offset"] = rng.integers(0, 256, size=n); ev["input_id"] = abi.EVENT_GATE; ev["scalar"] = 0.8


def run(e):
    e.set_value(0, freq); e.push_events(ev)
    e.enable_timing(True)
    outs = []
    for k in range(3):
        e.l2_flush(); e.process_block(256)
        outs.append(e.read_instance_out())
    ms = []
    for _ in range(steps):
        e.l2_flush(); e.process_block(256); ms.append(e.last_kernel_ms())
    return float(np.mean(ms)), np.concatenate(outs)


base = None
for name, mb, mask in (("generated, default", "8", ""), ("generated, uniform mask", "8", "0xFFFFFFE"), ("generated, uniform mask, 10 CTAs/SM", "10", "0xFFFFFFE"),
                       ("hand-written", None, None)):
    if mb is None:
        e = abi.Engine(abi.GRAPH_FM_VOICE, n, 8, 48000.0, 0, abi.PLAN_FUSED, abi.KEEP_INSTANCE_OUT)
    else:
        os.environ["OSCEN_B200_JIT_MIN_BLOCKS"] = mb
        os.environ["OSCEN_B200_JIT_UNIFORM_MASK"] = mask
        e = abi.Engine.from_graph(abi.Graph(text), n, 8, keep=abi.KEEP_INSTANCE_OUT, plan=abi.PLAN_FUSED)
    k_ms, out = run(e)
    if base is None:
        base = out
    print(json.dumps({"variant": name, "voices": n, "kernel_ms": k_ms, "voice_samples_per_s": n * 256 / (k_ms * 1e-3),
                      "bit_identical_to_default": bool(np.array_equal(out, base))}), flush=True)
    e.close()
