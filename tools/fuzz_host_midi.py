"""Fuzz of the DEVICE-side MIDI front WITHOUT a GPU: random raw-MIDI streams -> midi_front_kernel executed on the host ->
per-voice events -> the hand-written voice kernel executed on the host, against the oracle bank fed the same messages.
TEST TOOL.    PYTHONPATH=tests:. python tools/fuzz_host_midi.py FIRST_SEED LAST_SEED"""
import sys
import time

import numpy as np

import oracle_lib as O
from host_kernel import HostBuiltinEngine, HostMidiFront
from test_parity_gpu import _random_midi


def run(kind, vpi, seed):
    rng = np.random.default_rng(5550000 + 100 * kind + seed)
    n_inst = int(rng.integers(1, 7))
    n = n_inst * vpi
    ref, host, front = O.OracleBank(kind, n, vpi), HostBuiltinEngine(kind, n, vpi), HostMidiFront(n, vpi, 0)
    ok = True
    for blk in range(int(rng.integers(2, 5))):
        frames = int(rng.choice([1, 16, 33, 64, 128, 200]))
        msgs = _random_midi(rng, n_inst, frames, int(rng.choice([3, 10, 45])), int(rng.choice([4, vpi, 2 * vpi, 3 * vpi])))
        for inst, frame, data in msgs:
            ref.push_midi(inst, frame, data)
        host.push_events(front.run(msgs, frames))
        ref.process_block(frames); host.process_block(frames)
        ok &= np.array_equal(ref.read_voice_out(), host.read_voice_out()) and np.array_equal(ref.read_instance_out(), host.read_instance_out())
    front.close()
    if not ok:
        print("MISMATCH kind", kind, "seed", seed, flush=True)
    return "ok" if ok else "MISMATCH"


if __name__ == "__main__":
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    tally, t0 = {}, time.time()
    for seed in range(lo, hi):
        for kind, vpi in ((O.K_FM, 8), (O.K_EPIANO, 16), (O.K_FM, 4), (O.K_FM, 16)):
            try:
                r = run(kind, vpi, seed)
            except Exception as ex:
                r = "EXC"; print("EXC", kind, vpi, seed, repr(ex)[:300], flush=True)
            tally[(kind, vpi, r)] = tally.get((kind, vpi, r), 0) + 1
    print(sorted(tally.items()), "%.0fs" % (time.time() - t0), flush=True)
