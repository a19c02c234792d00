"""Differential fuzz of the generated kernels WITHOUT a GPU: random graphs (tests/random_graphs.py) -> DSL text -> generated CUDA text
-> executed on the host (tests/host_kernel.py) vs the per-sample walk over the oracle.  TEST TOOL.
    PYTHONPATH=tests:. python tools/fuzz_host_kernels.py FIRST_SEED LAST_SEED [--ramps]
--ramps: on odd seeds the first value input is uniform and ramps (set_<v>_with_ramp) across the block boundary; on seeds
divisible by 3 the last value input takes frame-accurate per-voice value events.
--extended: oscillator frequencies, envelope times, filter cutoffs, delay time / feedback also come from value inputs."""
import sys, time
import numpy as np
import random_graphs as R
import oracle_lib as O
from parity_util import lcg_noise
from host_kernel import HostEngine, HostNodePlanEngine
from oscen_b200 import abi
GATE = O.EV_GATE
BLOCKS = {"uniform": 0, "general": 0}        # blocks of the generated kernel rendered by each variant (--specialize)


def run(kind, seed, ramps=False, extended=False, specialize=False):
    rng = np.random.default_rng(777000 + seed)
    rg = R.make_graph(rng, feedback=(kind == "feedback"), multirate=(kind == "multirate"), extended=extended)
    src = rg.dsl()
    n, frames = 2, 72
    vpi = 1
    if ramps and seed % 4 == 0:                                # instances of two voices: the in-order instance fold is checked too
        n, vpi = 4, 2
    g = abi.Graph(src)
    vals = {name: (d * (1.0 + 0.1 * np.arange(n))).astype(np.float32) for name, d in rg.values.items()}
    uniform = set()
    if specialize:                                             # --specialize: a random subset of the inputs holds ONE value for the bank
        for name in sorted(rg.values):
            if rng.random() < 0.6:
                uniform.add(name); vals[name] = np.full(n, np.float32(rg.values[name]) * np.float32(1.25), dtype=np.float32)
    x = np.stack([lcg_noise(frames, 300 + seed * 8 + v) for v in range(n)], axis=1) if rg.has_input else None
    gates = [(v, 2 + 3 * v, 0.5 + 0.1 * v) for v in range(n)] + [(v, 50 + v, 0.0) for v in range(n)] if rg.has_gate else []
    # odd seeds: the first value input is uniform and RAMPS (set_<v>_with_ramp) from frame 0 across the block boundary
    ramp = None
    if ramps and seed % 2 == 1 and rg.values:
        import ctypes as C
        name = sorted(rg.values)[0]
        d = float(np.float32(rg.values[name]))
        r = O.Ramp(); O.lib().orc_ramp_new(C.byref(r), d)
        O.lib().orc_ramp_set_with_ramp(C.byref(r), float(np.float32(d * 1.5 + 0.1)), 45)
        table = []
        for _ in range(frames):
            O.lib().orc_ramp_tick(C.byref(r)); table.append(np.float32(r.current))
        ramp = (name, d, float(np.float32(d * 1.5 + 0.1)), 45)
        vals = dict(vals); vals[name] = (lambda f, table=table: table[f])
    # seeds divisible by 3 (with --ramps): frame-accurate VALUE EVENTS on the last value input, one per voice at its own frame
    value_events = []
    if ramps and seed % 3 == 0 and rg.values:
        name = sorted(rg.values)[-1]
        if not (ramp and ramp[0] == name):
            base = np.asarray(vals[name], dtype=np.float32)
            new = (base * np.float32(0.7) + np.float32(0.05)).astype(np.float32)
            when = [17 + 9 * v for v in range(n)]                       # voice 1's lands right before the block boundary
            step = lambda f, v, base=base, new=new, when=when: (new[v] if f >= when[v] else base[v])
            step.per_voice = True
            vals = dict(vals); vals[name] = step
            value_events = [(v, when[v], name, float(new[v])) for v in range(n)]
    want = rg.walk(n, frames, vals, gates, x)
    if not np.isfinite(want).all(): return "skip"
    engines = [("generated kernel", HostEngine(g, n, vpi))] + ([("node plan", HostNodePlanEngine(src, n, vpi))] if kind == "plain" else [])
    for label, eng in engines:
        for name, arr in vals.items():
            if ramp and name == ramp[0]:
                eng.set_value(g.input_id(name), ramp[1]); eng.set_value_ramp(g.input_id(name), ramp[2], ramp[3])
            elif callable(arr):
                eng.set_value(g.input_id(name), float(base[0]) if name in uniform else base)
            elif name in uniform:
                eng.set_value(g.input_id(name), float(arr[0]))                 # broadcast
            else:
                eng.set_value(g.input_id(name), arr)
        if specialize and label == "generated kernel":
            eng.specialize()                   # oscen_b200_specialize: the constant-bank variant, picked per block with fallback
        got, inst = [], []
        for lo, hi in ((0, 29), (29, 72)):
            if x is not None: eng.stream_in(0, x[lo:hi])
            ev = [(v, f - lo, GATE, s) for v, f, s in gates if lo <= f < hi]
            ev += [(v, f - lo, g.input_id(nm), s) for v, f, nm, s in value_events if lo <= f < hi]
            if ev: eng.push_events(ev)
            eng.process_block(hi - lo)
            if label == "generated kernel": BLOCKS[eng.last_kernel] += 1
            got.append(eng.read_voice_out()); inst.append(eng.read_instance_out())
        got, inst = np.concatenate(got), np.concatenate(inst)
        eng.close()
        fold = np.zeros((frames, n // vpi), dtype=np.float32)     # `.iter().map(|v| v.out).sum()` in voice order
        for k in range(vpi):
            fold = fold + want.reshape(frames, n // vpi, vpi)[:, :, k]
        if not np.array_equal(inst[:, :, 0], fold):
            print("MISMATCH (instance fold)", kind, seed, label, flush=True)
            return "MISMATCH"
        if not np.array_equal(got, want):
            print("MISMATCH", kind, seed, label, float(np.nanmax(np.abs(got - want))), "\n", src, flush=True)
            return "MISMATCH"
    return "ok"
if __name__ == "__main__":
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    ramps = "--ramps" in sys.argv
    extended = "--extended" in sys.argv
    specialize = "--specialize" in sys.argv
    tally = {}
    t0 = time.time()
    for seed in range(lo, hi):
        for kind in ("plain", "feedback", "multirate"):
            try:
                r = run(kind, seed, ramps, extended, specialize)
            except Exception as ex:
                r = "EXC"; print("EXC", kind, seed, repr(ex)[:300], flush=True)
            tally[(kind, r)] = tally.get((kind, r), 0) + 1
    print(sorted(tally.items()), "%.0fs" % (time.time() - t0), ("blocks by variant: %s" % BLOCKS) if specialize else "", flush=True)
