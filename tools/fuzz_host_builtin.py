"""Script fuzzer for the HAND-WRITTEN fused kernels WITHOUT a GPU: a random sequence of host calls (uniform / per-voice
set_value, set_value_ramp, gate events, frame-accurate value events, odd block sizes) is played on a built-in graph kind's
kernel executed on the host (tests/host_kernel.py HostBuiltinEngine) and on the oracle bank; outputs must be
bit-identical.  TEST TOOL.
    PYTHONPATH=tests:. python tools/fuzz_host_builtin.py FIRST_SEED LAST_SEED"""
import sys
import time

import numpy as np

import oracle_lib as O
from host_kernel import HostBuiltinEngine
from oscen_b200 import abi
from parity_util import lcg_noise

GATE = O.EV_GATE
# per kind: inputs that may take any of (lo, hi); input ids follow the reference graphs' declaration order
RANGES = {
    O.K_FM: {0: (60, 1500), 1: (0.5, 4), 2: (0, 1), 3: (0, 0.6), 4: (0.001, 0.02), 5: (0.005, 0.1), 6: (0.1, 0.9), 7: (0.01, 0.3),
             8: (0.5, 3), 9: (0, 1), 10: (0, 0.5), 13: (0.1, 0.9), 15: (0.5, 2), 18: (0.2, 0.9), 20: (0, 1), 21: (300, 6000),
             22: (0.4, 3), 25: (0.1, 0.9), 27: (0, 2500)},
    O.K_TWIN: {0: (200, 4000), 1: (400, 6000), 2: (0.0, 0.9)},
    O.K_SAT1: {0: (100, 5000), 1: (0.1, 1.0)}, O.K_SAT4: {0: (100, 5000), 1: (0.1, 1.0)},
    O.K_SYNTH: {0: (0.5, 12), 1: (0, 0.5), 2: (50, 3000), 3: (200, 9000)},
    O.K_SYNTH_ADSR: {0: (0.5, 12), 1: (0, 0.5), 2: (50, 3000), 3: (200, 9000), 4: (0.001, 0.02), 5: (0.005, 0.1), 6: (0.1, 0.9), 7: (0.01, 0.3)},
    O.K_EPIANO: {0: (60, 1200), 1: (0, 100), 2: (0, 100), 3: (0, 100), 4: (0, 100), 5: (0, 100), 6: (0, 100), 7: (0, 1), 8: (0.5, 9)},
}
HAS_GATE = {O.K_FM, O.K_SYNTH_ADSR, O.K_EPIANO}


def make_script(kind, seed):
    """-> (n_voices, voices_per_instance, [(frames, [call, ...]), ...], stream input or None)"""
    rng = np.random.default_rng(880000 + 100 * kind + seed)
    vpi = {O.K_FM: 8, O.K_EPIANO: 16}.get(kind, int(rng.choice([1, 2, 4])))
    n = vpi * int(rng.integers(1, 4)) if vpi > 1 else int(rng.integers(1, 6))
    if rng.random() < 0.3:
        n = (70 // vpi + 1) * vpi                                   # two CTAs, the second partly filled
    ranges = RANGES[kind]
    ids = sorted(ranges)
    n_blocks = int(rng.integers(2, 5))
    script = []
    ramped, per_voice, evented = set(), set(), set()
    for blk in range(n_blocks):
        frames = int(rng.choice([1, 7, 31, 32, 33, 64, 100, 129]))
        calls = []
        for _ in range(int(rng.integers(0, 4))):
            i = int(rng.choice(ids)); lo, hi = ranges[i]
            r = rng.random()
            instance_level = kind == O.K_EPIANO and i in (7, 8)    # the instance's Tremolo: uniform values and ramps only
            if r < 0.35 and i not in ramped and not instance_level:
                calls.append(("set_value", i, rng.uniform(lo, hi, n).astype(np.float32))); per_voice.add(i)
            elif r < 0.6:
                calls.append(("set_value", i, float(np.float32(rng.uniform(lo, hi))))); per_voice.discard(i)
            elif r < 0.8 and i not in per_voice and i not in evented:
                calls.append(("set_value_ramp", i, float(np.float32(rng.uniform(lo, hi))), int(rng.choice([1, 5, 40, 150]))))
                ramped.add(i)
            elif i not in ramped and not instance_level:
                ev = [(int(v), int(rng.integers(0, frames)), i, float(np.float32(rng.uniform(lo, hi)))) for v in rng.choice(n, size=min(n, 3), replace=False)]
                calls.append(("push_events", ev)); evented.add(i); per_voice.add(i)
        if kind in HAS_GATE and rng.random() < 0.8:
            ev = [(int(v), int(rng.integers(0, frames)), GATE, float(np.float32(rng.choice([0.0, 0.3, 0.9])))) for v in rng.choice(n, size=min(n, 5), replace=False)]
            calls.append(("push_events", ev))
        script.append((frames, calls))
    x = np.stack([lcg_noise(600, 50 + seed + v) for v in range(n)], axis=1) if kind == O.K_TWIN else None
    return n, vpi, script, x


def play(b, script, x):
    ys, insts, pos = [], [], 0
    for frames, calls in script:
        for c in calls:
            getattr(b, c[0])(*c[1:])
        if x is not None:
            b.stream_in(0, x[pos:pos + frames])
        b.process_block(frames); ys.append(b.read_voice_out()); insts.append(b.read_instance_out()); pos += frames
    return np.concatenate(ys), np.concatenate(insts)


def describe(script):
    for frames, calls in script:
        print("   block", frames, [(c[0],) + tuple((c[1][:3] if c[0] == "push_events" else (c[1], "per-voice" if np.ndim(c[2]) else round(float(c[2]), 4)) + tuple(c[3:]))) for c in calls], flush=True)


def run(kind, seed):
    n, vpi, script, x = make_script(kind, seed)
    (vr, ir), (vh, ih) = play(O.OracleBank(kind, n, vpi), script, x), play(HostBuiltinEngine(kind, n, vpi), script, x)
    if not np.isfinite(vr).all():
        return "skip"
    if np.array_equal(vr, vh) and np.array_equal(ir, ih):
        return "ok"
    print("MISMATCH kind", kind, "seed", seed, "n", n, "vpi", vpi, "voice diff", float(np.nanmax(np.abs(vr - vh))), "inst equal", np.array_equal(ir, ih), flush=True)
    describe(script)
    return "MISMATCH"


if __name__ == "__main__":
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    tally, t0 = {}, time.time()
    for seed in range(lo, hi):
        for kind in sorted(RANGES):
            try:
                r = run(kind, seed)
            except Exception as ex:
                r = "EXC"; print("EXC", kind, seed, repr(ex)[:300], flush=True)
            tally[(kind, r)] = tally.get((kind, r), 0) + 1
    print(sorted(tally.items()), "%.0fs" % (time.time() - t0), flush=True)
