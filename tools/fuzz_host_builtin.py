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


def make_script_fm_uniform(seed):
    """FMVoice scripts that stay on the SEGMENTED kernel's ground (kernel_fm_seg.cuh): `frequency` per voice, every other
    input uniform, uniform set_value between blocks, ramps (the eight inputs the ramp variant follows, sometimes an envelope
    time / ratio / sustain: those must fall back to the general kernel), gates and `frequency` events at random frames,
    short envelope times (stage boundaries inside blocks, different per lane), now and then levels large enough to break
    the sine-argument bound (general kernel again)"""
    rng = np.random.default_rng(990000 + seed)
    vpi = 8
    n = vpi * int(rng.choice([1, 2, 4, 5, 9, 12]))
    ranges = dict(RANGES[O.K_FM])
    ranges.update({4: (0.0, 0.004), 5: (0.0005, 0.01), 7: (0.0, 0.01), 11: (0.0, 0.004), 12: (0.0005, 0.02), 14: (0.0005, 0.01),
                   16: (0.0, 0.003), 17: (0.0005, 0.02), 19: (0.0, 0.02), 23: (0.0, 0.004), 24: (0.0005, 0.02), 26: (0.0005, 0.02)})
    seg_ramp_ids = [2, 3, 9, 10, 20, 21, 22, 27]
    ids = [i for i in sorted(ranges) if i != 0]
    script = []
    ramped = set()
    for blk in range(int(rng.integers(3, 7))):
        frames = int(rng.choice([1, 31, 32, 33, 64, 100, 129, 256]))
        calls = []
        if blk == 0:
            calls.append(("set_value", 0, rng.uniform(60, 1500, n).astype(np.float32)))
            for i in rng.choice(ids, size=6, replace=False):
                lo, hi = ranges[int(i)]
                calls.append(("set_value", int(i), float(np.float32(rng.uniform(lo, hi)))))
        for _ in range(int(rng.integers(0, 3))):
            r = rng.random()
            if r < 0.45:
                i = int(rng.choice(seg_ramp_ids if rng.random() < 0.85 else ids)); lo, hi = ranges[i]
                calls.append(("set_value_ramp", i, float(np.float32(rng.uniform(lo, hi))), int(rng.choice([1, 5, 40, 150, 400]))))
                ramped.add(i)
            elif r < 0.8:
                i = int(rng.choice([j for j in ids if j not in ramped])); lo, hi = ranges[i]
                big = i in (2, 9) and rng.random() < 0.1
                calls.append(("set_value", i, float(np.float32(rng.uniform(15, 30) if big else rng.uniform(lo, hi)))))
            else:
                ev = [(int(v), int(rng.integers(0, frames)), 0, float(np.float32(rng.uniform(60, 1500)))) for v in rng.choice(n, size=min(n, 3), replace=False)]
                calls.append(("push_events", ev))
        if rng.random() < 0.85:
            k = int(rng.choice([1, 5, n]))
            ev = [(int(v), int(rng.integers(0, frames)), GATE, float(np.float32(rng.choice([0.0, 0.3, 0.9])))) for v in rng.choice(n, size=min(n, k), replace=False)]
            calls.append(("push_events", ev))
        script.append((frames, calls))
    return n, vpi, script, None


def run_fm_uniform(seed, tally):
    n, vpi, script, x = make_script_fm_uniform(seed)
    hb = HostBuiltinEngine(O.K_FM, n, vpi)
    (vr, ir), (vh, ih) = play(O.OracleBank(O.K_FM, n, vpi), script, x), play(hb, script, x)
    for v in hb.variants:
        tally[("variant", v)] = tally.get(("variant", v), 0) + 1
    if not np.isfinite(vr).all():
        return "skip"
    if np.array_equal(vr, vh) and np.array_equal(ir, ih):
        return "ok"
    print("MISMATCH fm-uniform seed", seed, "n", n, "voice diff", float(np.nanmax(np.abs(vr - vh))), "inst equal", np.array_equal(ir, ih), hb.variants, flush=True)
    describe(script)
    return "MISMATCH"


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
    if "--fm-uniform" in sys.argv:
        for seed in range(lo, hi):
            try:
                r = run_fm_uniform(seed, tally)
            except Exception as ex:
                r = "EXC"; print("EXC fm-uniform", seed, repr(ex)[:300], flush=True)
            tally[("fm-uniform", r)] = tally.get(("fm-uniform", r), 0) + 1
        print(sorted(tally.items()), "%.0fs" % (time.time() - t0), flush=True)
        sys.exit(0)
    for seed in range(lo, hi):
        for kind in sorted(RANGES):
            try:
                r = run(kind, seed)
            except Exception as ex:
                r = "EXC"; print("EXC", kind, seed, repr(ex)[:300], flush=True)
            tally[(kind, r)] = tally.get((kind, r), 0) + 1
    print(sorted(tally.items()), "%.0fs" % (time.time() - t0), flush=True)
