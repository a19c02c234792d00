"""GPU parity tests of graphs compiled from `graph!` DSL text (oscen_b200_graph_compile +
oscen_b200_create_from_graph): (1) the reference graphs written as DSL against the oracle's
hand-written banks of the same graphs, bit for bit; (2) graphs the oracle has no bank for — the
reference's own expression / fan-in test graphs (oscen-lib/tests/binary_expression_test.rs,
stream_fanin.rs) and endpoint-default cases — against per-sample walks over the oracle's node
primitives (tests/graph_walk.py), also bit for bit."""
import os

import numpy as np
import pytest

import graph_walk as W
import oracle_lib as O
from parity_util import assert_close, lcg_ints, lcg_noise, lcg_uniform, seq_sum_f32
from test_parity_gpu import _fm_bank_script

pytestmark = pytest.mark.gpu

GATE = O.EV_GATE
GRAPHS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "graphs")
SR = 48000.0
f32 = np.float32


def text(name):
    with open(os.path.join(GRAPHS, name)) as f:
        return f.read()


def dsl_engine(src, n, vpi=1, sr=SR):
    from oscen_b200 import abi
    return abi.Engine.from_graph(src, n, vpi, sr)


# ------------------------------------------------------------ (1) reference graphs, DSL vs oracle banks
def test_dsl_fm_voice_bank_bit_exact_vs_oracle_and_fused():
    from oscen_b200 import abi
    n, vpi = 512, 8
    ref = O.OracleBank(O.K_FM, n, vpi)
    dsl = dsl_engine(text("fm_voice.graph"), n, vpi)
    fused = abi.Engine(abi.GRAPH_FM_VOICE, n, vpi)
    (vr, ir, mr), (vd, idd, md), (vf, iff, mf) = [_fm_bank_script(b, n) for b in (ref, dsl, fused)]
    assert np.abs(vr).max() > 0.05
    assert np.array_equal(vd, vr) and np.array_equal(vd, vf)
    assert np.array_equal(idd, ir)
    assert np.array_equal(seq_sum_f32(vd.reshape(vd.shape[0], n // vpi, vpi), axis=2), idd[:, :, 0])
    scale = np.abs(ir.astype(np.float64)).sum(axis=1)
    assert (np.abs(md.astype(np.float64) - mr.astype(np.float64)) <= 1e-6 + 2e-6 * scale).all()


@pytest.mark.parametrize("file,kind", [("synth.graph", O.K_SYNTH), ("synth_adsr.graph", O.K_SYNTH_ADSR),
                                       ("sat_1x.graph", O.K_SAT1), ("twin_peaks.graph", O.K_TWIN)])
def test_dsl_reference_graphs_bit_exact_vs_oracle(file, kind):
    n = 160
    ref, dsl = O.OracleBank(kind, n, 1), dsl_engine(text(file), n, 1)
    x = np.stack([lcg_noise(512, 77 + v) for v in range(n)], axis=1) if kind == O.K_TWIN else None

    def script(b):
        out = []
        if kind in (O.K_SYNTH, O.K_SYNTH_ADSR):
            b.set_value(2, (80.0 + 9.0 * np.arange(n)).astype(np.float32))
            b.set_value(3, (400.0 + 50.0 * np.arange(n)).astype(np.float32))
        elif kind == O.K_SAT1:
            b.set_value(0, (500.0 + 20.0 * np.arange(n)).astype(np.float32))
        for blk, frames in enumerate((256, 100, 156)):
            if kind == O.K_TWIN:
                b.stream_in(0, x[blk * 156:blk * 156 + frames])
            if kind == O.K_SYNTH_ADSR and blk < 2:
                b.push_events([(v, (v + 3 * blk) % 100, GATE, 0.9 if blk == 0 else 0.0) for v in range(n)])
            if blk == 1 and kind != O.K_SAT1:
                b.set_value_ramp(0, 7.0 if kind != O.K_TWIN else 1200.0, 100)
                b.push_events([(5, 40, 1, 0.3 if kind != O.K_TWIN else 2500.0)])      # a value event at frame 40
            b.process_block(frames)
            out.append(b.read_voice_out())
        return np.concatenate(out)

    yr, yd = script(ref), script(dsl)
    assert np.abs(yr).max() > 0.01
    assert np.array_equal(yd, yr), float(np.abs(yd - yr).max())


def test_dsl_twin_peaks_device_lcg_input():
    n = 130
    ref, dsl = O.OracleBank(O.K_TWIN, n, 1), dsl_engine(text("twin_peaks.graph"), n, 1)
    x = np.stack([lcg_noise(256, 1000 + 7 + v) for v in range(n)], axis=1)
    ref.stream_in(0, x); ref.process_block(256)
    dsl.stream_in_lcg(0, 1000, 7); dsl.process_block(256)
    assert np.array_equal(dsl.read_voice_out(), ref.read_voice_out())


# ------------------------------------------------------------ (2) graphs without an oracle bank: walks
TWO_OSC = """
    output stream out;
    nodes { osc1 = PolyBlepOscillator::sine(440.0, 0.5); osc2 = PolyBlepOscillator::saw(660.0, 0.25); osc3 = PolyBlepOscillator::square(110.0, 0.3); }
"""


def two_osc_voice(expr):
    def make(v):
        o1, o2, o3 = W.polyblep("sine", 440.0, 0.5, SR), W.polyblep("saw", 660.0, 0.25, SR), W.polyblep("square", 110.0, 0.3, SR)

        def frame(vals, x):
            return expr(o1.process(), o2.process(), o3.process(), vals)
        return frame, None
    return make


# oscen-lib/tests/binary_expression_test.rs:15,49,83,119,154,188,223 plus the f32 methods
EXPR_CASES = [
    ("osc1.output * gain -> out;", {"gain": 0.5}, lambda a, b, c, p: f32(a * p["gain"])),
    ("osc1.output + osc2.output -> out;", {}, lambda a, b, c, p: f32(a + b)),
    ("osc1.output - osc2.output -> out;", {}, lambda a, b, c, p: f32(a - b)),
    ("osc1.output * envelope + offset -> out;", {"envelope": 0.75, "offset": 0.125}, lambda a, b, c, p: f32(f32(a * p["envelope"]) + p["offset"])),
    ("osc1.output + osc2.output + osc3.output -> out;", {}, lambda a, b, c, p: f32(f32(a + b) + c)),
    ("osc1.output * osc2.output -> out;", {}, lambda a, b, c, p: f32(a * b)),
    ("osc1.output / divisor -> out;", {"divisor": 3.0}, lambda a, b, c, p: f32(a / p["divisor"])),
    ("2.0 - osc1.output -> out;", {}, lambda a, b, c, p: f32(f32(2.0) - a)),
    ("gain / (osc3.output + 1.5) -> out;", {"gain": 0.3}, lambda a, b, c, p: f32(p["gain"] / f32(c + f32(1.5)))),
    ("osc1.output / (osc2.output + 2.0) -> out;", {}, lambda a, b, c, p: f32(a / f32(b + f32(2.0)))),
    ("(osc1.output * 3.0).tanh() -> out;", {}, lambda a, b, c, p: W.tanh32(f32(a * f32(3.0)))),
    ("osc1.output.abs() - osc2.output.clamp(-0.1, 0.2) -> out;", {}, lambda a, b, c, p: f32(f32(abs(a)) - W.clamp32(b, -0.1, 0.2))),
    ("osc1.output.min(0.3) + osc2.output.max(-0.2) -> out;", {}, lambda a, b, c, p: f32(min(a, f32(0.3)) + max(b, f32(-0.2)))),
    ("gain * 2.0 + osc1.output -> out;", {"gain": 0.2}, lambda a, b, c, p: f32(f32(p["gain"] * f32(2.0)) + a)),
    ("osc1.output -> out; osc2.output -> out; bias -> out; osc3.output -> out;", {"bias": 0.05},
     lambda a, b, c, p: f32(f32(f32(a + b) + p["bias"]) + c)),               # fan-in into the graph output, statement order
]


@pytest.mark.parametrize("case", range(len(EXPR_CASES)))
def test_dsl_expression_edges_match_per_sample_walk(case):
    conns, params, expr = EXPR_CASES[case]
    decls = "".join(f"input {k}: value = {v};\n" for k, v in params.items())
    src = decls + TWO_OSC + "connections { " + conns + " }"
    n, frames = 8, 200
    eng = dsl_engine(src, n)
    vals = {k: f32(v) for k, v in params.items()}
    want = W.render(two_osc_voice(expr), 1, frames, vals)
    eng.process_block(frames)
    got = eng.read_voice_out()
    assert np.abs(want).max() > 0.01
    for v in range(n):                          # every voice renders the same graph with the same inputs
        assert np.array_equal(got[:, v], want[:, 0]), (conns, float(np.abs(got[:, v] - want[:, 0]).max()))


def test_dsl_per_voice_and_ramped_values_in_expressions():
    """a value input in an expression is read per frame: per-voice values, a ramp (`.current` after
    tick_ramps, codegen/emit_node.rs:317-321) and a frame-accurate value event"""
    src = "input gain: value = 0.5;\ninput offset: value = 0.0;\n" + TWO_OSC + "connection osc1.output * gain + offset -> out;"
    n, frames = 6, 160
    eng = dsl_engine(src, n)
    gains = (0.1 * (1 + np.arange(n))).astype(np.float32)
    eng.set_value(0, gains)
    eng.set_value_ramp(1, 0.5, 100)
    eng.push_events([(2, 37, 0, 0.9)])
    eng.process_block(frames)
    got = eng.read_voice_out()
    ramp = O.Ramp(); L = O.lib(); L.orc_ramp_new(C_byref(ramp), 0.0); L.orc_ramp_set_with_ramp(C_byref(ramp), 0.5, 100)
    offs = []
    for f in range(frames):
        L.orc_ramp_tick(C_byref(ramp)); offs.append(f32(ramp.current))
    for v in range(n):
        o1 = W.polyblep("sine", 440.0, 0.5, SR)
        for f in range(frames):
            g = f32(0.9) if (v == 2 and f >= 37) else gains[v]
            want = f32(f32(o1.process() * g) + offs[f])
            assert got[f, v] == want, (v, f, got[f, v], want)


def C_byref(x):
    import ctypes
    return ctypes.byref(x)


def test_dsl_stream_fan_in_and_value_to_stream_edges():
    """stream_fanin.rs: >= 2 sources into one stream input are summed in statement order, a single source stays a
    plain copy; Value -> Stream edges broadcast the value per frame (lower.rs:1163); unconnected stream inputs keep
    their constructor value (FmOperator.envelope = 1.0, fm_operator.rs:41)"""
    src = """
        name: FanInVoice;
        input freq_offset: value = 0.25;
        input cutoff: value = 900.0;
        input fm_depth: value = 0.1;
        output audio_out: stream;
        nodes {
            lfo = PolyBlepOscillator::triangle(7.0, 1.0);
            a = PolyBlepOscillator::saw(220.0, 0.4);
            b = PolyBlepOscillator::square(331.0, 0.3);
            op = FmOperator::new();
            filter = TptFilter::new(1000.0, 0.9);
            clip = HardClip::new();
        }
        connections {
            freq_offset -> a.frequency_mod;           // Value -> Stream: broadcast
            lfo.output * fm_depth -> b.phase_mod;
            a.output -> filter.input;                 // three-way fan-in, statement order
            b.output -> filter.input;
            op.output -> filter.input;
            cutoff -> filter.cutoff;
            lfo.output * 0.5 -> filter.f_mod;
            a.output * 0.125 -> op.phase_mod;
            filter.output -> clip.input;
            clip.output -> audio_out;
        }
    """
    n, frames = 5, 300
    eng = dsl_engine(src, n)
    cut = (500.0 + 700.0 * np.arange(n)).astype(np.float32)
    eng.set_value(1, cut)
    eng.process_block(frames)
    got = eng.read_voice_out()

    def make(v):
        lfo, a, b = W.polyblep("triangle", 7.0, 1.0, SR), W.polyblep("saw", 220.0, 0.4, SR), W.polyblep("square", 331.0, 0.3, SR)
        op, flt = W.fmop(SR), W.tpt(1000.0, 0.9, SR)

        def frame(p, x):
            l = lfo.process()
            a.set(frequency_mod=p["freq_offset"]); ya = a.process()
            b.set(phase_mod=f32(l * p["fm_depth"])); yb = b.process()
            op.set(phase_mod=f32(ya * f32(0.125))); yo = op.process()
            flt.set(input=f32(f32(ya + yb) + yo), cutoff=p["cutoff"], f_mod=f32(l * f32(0.5)))
            return W.hardclip(flt.process())
        return frame, None

    want = W.render(make, n, frames, {"freq_offset": 0.25, "cutoff": cut, "fm_depth": 0.1})
    assert np.abs(want).max() > 0.05
    assert np.array_equal(got, want), float(np.abs(got - want).max())


def test_dsl_envelope_gate_wiring_and_unconnected_envelope():
    """only envelopes the graph's event input is connected to react to gate events; an unconnected one stays idle"""
    src = """
        input gate: event;
        input level: value = 0.8;
        output stream out;
        nodes {
            osc = PolyBlepOscillator::sine(330.0, 1.0);
            env = AdsrEnvelope::new(0.002, 0.01, 0.6, 0.001);
            idle_env = AdsrEnvelope::new(0.002, 0.01, 0.6, 0.001);
            vca = Gain::new(1.0);
        }
        connections {
            gate -> env.gate;
            osc.output -> vca.input;
            env.output -> vca.gain;
            vca.output * level + idle_env.output -> out;
        }
    """
    n, frames = 4, 400
    eng = dsl_engine(src, n)
    gates = [(v, 10 * v, 0.5 + 0.1 * v) for v in range(n)] + [(v, 250 + v, 0.0) for v in range(n)]
    eng.push_events([(v, f, GATE, s) for v, f, s in gates])
    eng.process_block(frames)
    got = eng.read_voice_out()

    def make(v):
        osc, env, idle = W.polyblep("sine", 330.0, 1.0, SR), W.adsr(0.002, 0.01, 0.6, 0.001, SR), W.adsr(0.002, 0.01, 0.6, 0.001, SR)

        def frame(p, x):
            e = env.process(); i = idle.process()
            return f32(f32(f32(osc.process() * e) * p["level"]) + i)
        return frame, env.on_gate

    want = W.render(make, n, frames, {"level": 0.8}, gates)
    assert np.abs(want).max() > 0.1 and np.abs(want[320:]).max() < 1e-3
    assert np.array_equal(got, want), float(np.abs(got - want).max())


def test_dsl_stream_input_passthrough_and_filters():
    src = """
        input audio_in: stream;
        input cutoff: value = 700.0;
        output stream out;
        node lp = LP18Filter::new(1500.0, 0.4);
        node post = TptFilter::new(3000.0, 0.5);
        connection audio_in -> lp.input;
        connection cutoff -> lp.cutoff;
        connection lp.output -> post.input;
        connection post.output * 0.5 + audio_in * 0.25 -> out;
    """
    n, frames = 6, 256
    eng = dsl_engine(src, n)
    x = np.stack([lcg_noise(frames, 500 + v) for v in range(n)], axis=1)
    eng.stream_in(0, x)
    eng.process_block(frames)
    got = eng.read_voice_out()

    def make(v):
        lp, post = W.lp18(1500.0, 0.4, SR), W.tpt(3000.0, 0.5, SR)

        def frame(p, xin):
            lp.set(input=xin, cutoff=p["cutoff"])
            post.set(input=lp.process())
            return f32(f32(post.process() * f32(0.5)) + f32(xin * f32(0.25)))
        return frame, None

    want = W.render(make, n, frames, {"cutoff": 700.0}, None, x)
    assert np.array_equal(got, want), float(np.abs(got - want).max())


def test_dsl_block_split_invariance_and_instances():
    """block == per-sample for any split (oscen-lib/tests/block_processing_test.rs) and instance sums in voice order"""
    n, vpi = 64, 4
    src = text("synth_adsr.graph")
    a, b = dsl_engine(src, n, vpi), dsl_engine(src, n, vpi)
    freqs = (60.0 + 11.0 * np.arange(n)).astype(np.float32)
    for e in (a, b):
        e.set_value(2, freqs)
    ev = [(v, 3 * v, GATE, 0.7) for v in range(n)] + [(v, 300 + v, GATE, 0.0) for v in range(n)]
    a.push_events(ev); a.process_block(512); ya, ia = a.read_voice_out(), a.read_instance_out()
    outs, insts, pos = [], [], 0
    for frames in (100, 1, 155, 256):
        b.push_events([(v, f - pos, i, s) for v, f, i, s in ev if pos <= f < pos + frames])
        b.process_block(frames); outs.append(b.read_voice_out()); insts.append(b.read_instance_out()); pos += frames
    assert np.array_equal(np.concatenate(outs), ya)
    assert np.array_equal(np.concatenate(insts), ia)
    assert np.array_equal(seq_sum_f32(ya.reshape(512, n // vpi, vpi), axis=2), ia[:, :, 0])


def test_dsl_full_size_fm_bank_65536_voices_matches_fused_kernel():
    """BASELINE configs[1] size through the DSL-compiled node plan: identical bits to the fused kernel"""
    from oscen_b200 import abi
    n, vpi = 65536, 8
    dsl = dsl_engine(text("fm_voice.graph"), n, vpi)
    fused = abi.Engine(abi.GRAPH_FM_VOICE, n, vpi)
    notes = 36 + lcg_ints(n, 1, 60)
    freq = (440.0 * np.exp2((notes - 69) / 12.0)).astype(np.float32)
    vel = (0.25 + 0.75 * lcg_uniform(n, 2)).astype(np.float32)
    on = lcg_ints(n, 3, 256)
    ev = np.zeros(n, dtype=abi.EVENT_DTYPE)
    ev["voice"] = np.arange(n); ev["frame_offset"] = on; ev["input_id"] = GATE; ev["scalar"] = vel
    for e in (dsl, fused):
        e.set_value(0, freq)
        e.push_events(ev)
        e.process_block(256)
    assert np.array_equal(dsl.read_instance_out(), fused.read_instance_out())
    sl = slice(40000, 40512)
    assert np.array_equal(dsl.read_voice_out(sl.start, 512), fused.read_voice_out(sl.start, 512))
