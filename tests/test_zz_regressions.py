"""Regressions found by the host-side differential fuzzer (tools/fuzz_host_kernels.py), kept as explicit cases on the GPU
(`-m gpu`, generated kernel) and on the host-executed kernel (CPU suite).
(The file name sorts last on purpose: the GPU variants were added after the round's GPU budget was spent.)"""
import numpy as np
import pytest

import oracle_lib as O
from test_dsl_gpu import dsl_engine, fused_plan, plan  # noqa: F401  (fixtures)

GATE = O.EV_GATE
SR = 48000.0

INNER_PARAMS = """
    input ratio: value = 1.5;
    input attack: value = 0.003;
    input gate: event;
    output stream out;
    nodes {
        env = AdsrEnvelope::new(0.003, 0.004, 0.5, 0.002) * 4;
        op = FmOperator::new() * 4;
        clip = HardClip::new() * 4;
    }
    connections {
        gate -> env.gate;
        attack -> env.attack;                    // an inner-rate envelope whose times come from a value input
        ratio -> op.ratio;                       // an inner-rate operator whose increment comes from a value input
        330.0 -> op.base_freq;
        env.output -> op.envelope;
        op.output -> clip.input;
        [sinc] clip.output -> out;
    }
"""


def test_inner_rate_nodes_keep_their_rate_when_a_frame_carries_events(fused_plan):
    """seeds 777000+1047 / +1131 of the host fuzzer: on a frame that carries an event or a ramp the generated kernel
    re-derives an FmOperator's phase increment and an AdsrEnvelope's step data from their value inputs — for a `* N` node
    that must happen at sr * N (codegen/emit_struct.rs:578-581), not at the graph's rate.  Properties that need no
    oracle: a value event that rewrites the value already there changes nothing, and any block split renders the same"""
    n, frames = 4, 160
    ratio = (1.0 + 0.25 * np.arange(n)).astype(np.float32)
    attack = (0.002 + 0.001 * np.arange(n)).astype(np.float32)
    gates = [(v, 3 + 2 * v, GATE, 0.8) for v in range(n)] + [(v, 90 + v, GATE, 0.0) for v in range(n)]

    def render(blocks, extra_events=()):
        eng = dsl_engine(INNER_PARAMS, n, 1, SR, plan=fused_plan)
        eng.set_value(0, ratio); eng.set_value(1, attack)
        out, pos = [], 0
        for b in blocks:
            ev = [(v, f - pos, i, s) for v, f, i, s in list(gates) + list(extra_events) if pos <= f < pos + b]
            if ev:
                eng.push_events(ev)
            eng.process_block(b); out.append(eng.read_voice_out()); pos += b
        return np.concatenate(out)

    base = render((frames,))
    assert np.abs(base).max() > 0.05
    assert np.array_equal(render((41, 7, 112)), base)                                     # block == per-sample for any split
    same_value = [(v, 30 + v, 0, float(ratio[v])) for v in range(n)] + [(v, 50, 1, float(attack[v])) for v in range(n)]
    assert np.array_equal(render((frames,), same_value), base)                            # rewriting a value with itself is a no-op
    # and the rate really is 4 x: one octave up in `ratio` is not what a wrong rate would look like — compare with base_freq doubled
    assert not np.array_equal(render((frames,), [(v, 30 + v, 0, float(ratio[v] * 2.0)) for v in range(n)]), base)


def test_value_events_on_a_ramped_input_are_refused(fused_plan):
    """found by running the sat graphs with a ramp AND value events on `frequency` through tests/host_kernel.py: once the
    ramp had completed, the engine kept the event's value while the oracle went back to the ramp's `.current` (which the
    generated struct reads every frame, codegen/emit_node.rs:317-321).  The combination has no counterpart in the
    reference — a ramped input is one ValueRampState — so engine, oracle and host harness now all refuse it"""
    import os
    from oscen_b200 import abi
    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "graphs", "sat_1x.graph")).read()
    eng = dsl_engine(src, 4, 1, SR, plan=fused_plan)
    eng.set_value(0, 900.0)
    eng.push_events([(1, 4, 0, 700.0)])                    # fine while the input has never been ramped
    eng.process_block(16)
    with pytest.raises(abi.EngineError):
        eng.set_value_ramp(0, 2500.0, 8)                   # the event left the input per-voice: a ramp needs a uniform input
    eng.set_value(0, 900.0)
    eng.set_value_ramp(0, 2500.0, 8)
    with pytest.raises(abi.EngineError):
        eng.push_events([(1, 4, 0, 700.0)])
    eng.push_events([(1, 4, 1, 0.5)])                      # the other input is unaffected
    eng.process_block(16)
    ref = O.OracleBank(O.K_SAT1, 4, 1)
    ref.set_value(0, 900.0); ref.push_events([(1, 4, 0, 700.0)]); ref.process_block(16)
    ref.set_value(0, 900.0); ref.set_value_ramp(0, 2500.0, 8); ref.push_events([(1, 4, 1, 0.5)]); ref.process_block(16)
    assert np.array_equal(eng.read_voice_out(), ref.read_voice_out())


def test_gate_off_starts_from_the_release_length_cached_by_the_previous_process(plan):
    """found by tools/fuzz_host_builtin.py: AdsrEnvelope::handle_gate_event (adsr.rs:252-273) counts a release down from
    `release_samples` as cached by the LAST process(), so a note-off on the frame where `release` grew — here: a host write
    between blocks with the note-off at frame 0, and a value event on the note-off's frame — releases over the old length
    (then clamped to the new one).  The length travels in the ADSR stage word across blocks (nodes.cuh AdsrState::r_gate)."""
    import graph_walk as W
    src = ("input r: value = 0.004; input gate: event; output stream out; node env = AdsrEnvelope::new(0.002, 0.003, 0.6, 0.004); "
           "connection gate -> env.gate; connection r -> env.release; connection env.output -> out;")
    n = 3
    eng = dsl_engine(src, n, 1, SR, plan=plan)
    envs = [W.adsr(0.002, 0.003, 0.6, 0.004, SR) for _ in range(n)]
    rel = np.full(n, 0.004, dtype=np.float32)
    got, want = [], []
    # block 0: notes on.  block 1: release doubled by the host, voice 0 released at frame 0, voice 1 at frame 9.
    # block 2: voice 2 gets a longer release by a value event on the frame of its note-off.
    script = [(200, None, [(v, 0, GATE, 0.9) for v in range(n)]),
              (100, 0.008, [(0, 0, GATE, 0.0), (1, 9, GATE, 0.0)]),
              (120, None, [(2, 7, 0, 0.02), (2, 7, GATE, 0.0)])]
    for frames, new_rel, events in script:
        if new_rel is not None:
            eng.set_value(0, new_rel); rel[:] = new_rel
        eng.push_events(events)
        eng.process_block(frames); got.append(eng.read_voice_out())
        out = np.zeros((frames, n), dtype=np.float32)
        for f in range(frames):
            for v in range(n):
                for ev, ef, ei, es in events:
                    if ev == v and ef == f and ei == 0:
                        rel[v] = es
                envs[v].set(release=rel[v])
                for ev, ef, ei, es in events:
                    if ev == v and ef == f and ei == GATE:
                        envs[v].on_gate(es)
                out[f, v] = envs[v].process()
        want.append(out)
    got, want = np.concatenate(got), np.concatenate(want)
    assert np.abs(want).max() > 0.5
    assert np.array_equal(got, want), float(np.abs(got - want).max())
