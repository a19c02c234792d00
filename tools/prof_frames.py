#!/usr/bin/env python
"""Fixed cost vs per-frame cost of the fused FMVoice kernel: device time of one block at 65 536 voices for block sizes
32..512, envelopes held in Decay (long decay times) with the cutoff following its envelope (the bench window's regime)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oscen_b200 import abi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
e = abi.Engine(abi.GRAPH_FM_VOICE, n, 8, 48000.0, 0, abi.PLAN_FUSED, abi.KEEP_INSTANCE_OUT)
rng = np.random.default_rng(1)
notes = 36 + rng.integers(0, 60, size=n)
e.set_value(0, (440.0 * np.exp2((notes - 69) / 12.0)).astype(np.float32))
for i in (5, 12, 18, 25):            # op3 / op2 / op1 / filter decay
    e.set_value(i, 20.0)
e.set_value(27, 1500.0)              # filter_env_amount
ev = np.zeros(n, dtype=abi.EVENT_DTYPE)
ev["voice"] = np.arange(n); ev["frame_offset"] = rng.integers(0, 256, size=n); ev["input_id"] = abi.EVENT_GATE; ev["scalar"] = 0.8
e.push_events(ev)
e.enable_timing(True)
for _ in range(8):
    e.process_block(256)
e.synchronize()
out = {}
for frames in (32, 64, 128, 256, 512, 256, 128, 64, 32):
    ms = []
    for _ in range(12):
        e.l2_flush(); e.process_block(frames); ms.append(e.last_kernel_ms())
    out.setdefault(frames, []).append(float(np.median(ms)))
    name = e.last_kernel_name()
res = {f: float(np.mean(v)) for f, v in out.items()}
fs = np.array(sorted(res)); ts = np.array([res[f] for f in fs])
slope, icpt = np.polyfit(fs, ts, 1)
print(json.dumps({"voices": n, "kernel": name, "ms_by_frames": res, "per_frame_us": slope * 1e3, "fixed_us": icpt * 1e3}))
