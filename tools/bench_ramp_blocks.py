#!/usr/bin/env python
"""Device time of a block in which a value input ramps, for the hand-written Synth+ADSR / TwinPeaks / SatGraph_4x kernels
(ramp blocks run in segments unless the ramp lands on an envelope input): tools/bench_ramp_blocks.py [voices]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oscen_b200 import abi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
FRAMES = 256
for name, kind, freq_id, ramp_id, lo, hi in (("synth_adsr: cutoff ramps", abi.GRAPH_SYNTH_ADSR, 2, 3, 800.0, 4000.0),
                                             ("twin_peaks: cutoff_a ramps", abi.GRAPH_TWIN_PEAKS, None, 0, 600.0, 3000.0),
                                             ("sat_4x: amplitude ramps", abi.GRAPH_SAT_4X, 0, 1, 0.2, 0.9)):
    e = abi.Engine(kind, n, 1, 48000.0, 0, abi.PLAN_FUSED, 0)
    rng = np.random.default_rng(1)
    if freq_id is not None:
        e.set_value(freq_id, (440.0 * np.exp2((36 + rng.integers(0, 60, size=n) - 69) / 12.0)).astype(np.float32))
    if kind == abi.GRAPH_TWIN_PEAKS:
        e.stream_in_lcg(0, 1000, 0)
    if kind == abi.GRAPH_SYNTH_ADSR:
        ev = np.zeros(n, dtype=abi.EVENT_DTYPE)
        ev["voice"] = np.arange(n); ev["frame_offset"] = rng.integers(0, FRAMES, size=n); ev["input_id"] = abi.EVENT_GATE; ev["scalar"] = 0.8
        e.push_events(ev)
    e.enable_timing(True)
    for _ in range(3):
        e.l2_flush(); e.process_block(FRAMES)
    plain, ramp = [], []
    for k in range(6):
        e.l2_flush(); e.process_block(FRAMES); plain.append(e.last_kernel_ms())
    for k in range(6):
        e.set_value_ramp(ramp_id, hi if k % 2 == 0 else lo, FRAMES)        # one ramp per block, the whole block long
        e.l2_flush(); e.process_block(FRAMES); ramp.append(e.last_kernel_ms())
    print(json.dumps({"case": name, "voices": n, "plain_block_ms": float(np.mean(plain)), "ramp_block_ms": float(np.mean(ramp)),
                      "lib": os.path.basename(os.environ.get("OSCEN_B200_LIB", "liboscen_b200.so"))}), flush=True)
    e.close()
