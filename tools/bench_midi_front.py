#!/usr/bin/env python
"""Cost of the device-side MIDI front (midi_device.cuh) at config-3 polyphony: 131 072 ElectricPiano
instances x 16 voices, every instance receives `k` note-ons in one block.  Reports host-clock time of
push_midi_batch + process_block + synchronize with and without the note burst."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oscen_b200 import abi  # noqa: E402


def main():
    n_inst, vpi, frames, k = 131072, 16, 256, 3
    e = abi.Engine(abi.GRAPH_EPIANO, n_inst * vpi, vpi, 48000.0, 0, abi.PLAN_FUSED, abi.KEEP_INSTANCE_OUT)
    rng = np.random.default_rng(1)

    def burst(status):
        m = np.zeros(n_inst * k, dtype=abi.MIDI_DTYPE)
        m["instance"] = np.repeat(np.arange(n_inst), k)
        m["frame_offset"] = rng.integers(0, frames, size=n_inst * k)
        m["bytes"][:, 0] = status
        m["bytes"][:, 1] = np.tile(np.array([60, 64, 67]), n_inst)
        m["bytes"][:, 2] = 100
        m["len"] = 3
        return m

    on, off = burst(0x90), burst(0x80)
    e.push_midi_batch(on[:k]); e.process_block(frames); e.synchronize()      # allocate the front's buffers
    res = {}
    for name, msgs in (("quiet", None), ("note_on_burst", on), ("quiet2", None), ("note_off_burst", off), ("quiet3", None)):
        ts = []
        for _ in range(3 if msgs is None else 1):
            t0 = time.perf_counter()
            if msgs is not None:
                e.push_midi_batch(msgs)
            t1 = time.perf_counter()
            e.process_block(frames); e.synchronize()
            ts.append((time.perf_counter() - t0, t1 - t0))
        res[name] = {"block_ms": 1e3 * min(t[0] for t in ts), "push_ms": 1e3 * min(t[1] for t in ts)}
    print(json.dumps({"workload": f"epiano {n_inst} instances x {vpi} voices, {k} messages per instance per burst ({n_inst * k} messages)",
                      "host_clock_ms": res, "stats": e.stats()}))


if __name__ == "__main__":
    main()
