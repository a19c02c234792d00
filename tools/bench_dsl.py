#!/usr/bin/env python
"""Device-timed throughput of DSL-compiled graphs (tests/graphs/*.graph): the generated fused kernel
(codegen.hpp + NVRTC), the node plan, and — where one exists — the hand-written fused kernel of the
same graph.  One JSON line per (graph, way):  tools/bench_dsl.py [steps] [voices]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oscen_b200 import abi  # noqa: E402

FRAMES = 256
CASES = [("voice_fm3.graph", abi.GRAPH_FM_VOICE, 8, 0), ("synth_adsr.graph", abi.GRAPH_SYNTH_ADSR, 1, 2),
         ("twin_peaks.graph", abi.GRAPH_TWIN_PEAKS, 1, None), ("sat_1x.graph", abi.GRAPH_SAT_1X, 1, 0),
         ("sat_4x.graph", abi.GRAPH_SAT_4X, 1, 0)]


def run(e, n, freq_id, steps, has_gate, has_in, specialize=False):
    rng = np.random.default_rng(1)
    if has_in:
        e.stream_in_lcg(0, 1000, 0)
    if freq_id is not None:
        notes = 36 + rng.integers(0, 60, size=n)
        e.set_value(freq_id, (440.0 * np.exp2((notes - 69) / 12.0)).astype(np.float32))
    if has_gate:
        ev = np.zeros(n, dtype=abi.EVENT_DTYPE)
        ev["voice"] = np.arange(n); ev["frame_offset"] = rng.integers(0, FRAMES, size=n)
        ev["input_id"] = abi.EVENT_GATE; ev["scalar"] = 0.8
        e.push_events(ev)
    if specialize:
        e.specialize()               # after the per-voice frequencies are in: they stay in registers
    e.enable_timing(True)
    for _ in range(3):
        e.l2_flush(); e.process_block(FRAMES)
    e.synchronize()
    ms = []
    for _ in range(steps):
        e.l2_flush(); e.process_block(FRAMES); ms.append(e.last_kernel_ms())
    return float(np.mean(ms))


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    only = sys.argv[3] if len(sys.argv) > 3 else None
    for file, kind, vpi, freq_id in CASES:
        if only and only not in file:
            continue
        text = open(os.path.join(ROOT, "tests", "graphs", file)).read()
        g = abi.Graph(text)
        has_gate = "event input:" in g.describe()
        has_in = "stream input:" in g.describe()
        keep = abi.KEEP_INSTANCE_OUT
        ways = []
        t0 = time.perf_counter()
        ways.append(("generated fused kernel (NVRTC)", abi.Engine.from_graph(g, n, vpi, keep=keep, plan=abi.PLAN_FUSED)))
        jit_s = time.perf_counter() - t0
        ways.append(("generated fused kernel, specialised for the uniform inputs (oscen_b200_specialize)", abi.Engine.from_graph(g, n, vpi, keep=keep, plan=abi.PLAN_FUSED)))
        if "Edge [" not in g.describe():       # multirate graphs run as generated kernels only
            ways.append(("node plan (CUDA graph)", abi.Engine.from_graph(g, n, vpi, keep=keep, plan=abi.PLAN_NODES)))
        ways.append(("hand-written fused kernel", abi.Engine(kind, n, vpi, 48000.0, 0, abi.PLAN_FUSED, keep)))
        if os.environ.get("OSCEN_BENCH_WAYS") == "gen":
            ways = ways[:2]
        for name, e in ways:
            k_ms = run(e, n, freq_id, steps, has_gate, has_in, specialize="specialised" in name)
            print(json.dumps({"graph": file, "way": name, "voices": n, "frames": FRAMES, "kernel_ms": k_ms,
                              "voice_samples_per_s": n * FRAMES / (k_ms * 1e-3),
                              "kernel": e.last_kernel_name(),
                              "create_s_incl_nvrtc": jit_s if "NVRTC" in name else None}), flush=True)
            e.close()


if __name__ == "__main__":
    main()
