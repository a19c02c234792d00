"""Node arrays of library types (`oscs = [PolyBlepOscillator::saw(..); N]`; oscen-lib/tests/array_test.rs,
array_fanin_stream_input.rs, multirate_array_fanout.rs): Broadcast / Parallel / FanIn edges and indexed elements, on the
GPU (node plan and generated kernel, `-m gpu`) and on the host-executed kernels (CPU suite), against a per-sample walk over
the oracle's primitives and against the hand-unrolled scalar graph.
(The file name sorts last on purpose: the GPU variants were added after the round's GPU budget was spent.)"""
import numpy as np
import pytest

import graph_walk as W
import oracle_lib as O
from test_dsl_gpu import dsl_engine, fused_plan, plan  # noqa: F401  (fixtures)

GATE = O.EV_GATE
f32 = np.float32
SR = 48000.0

UNISON = """
    name: Unison;
    input value detune = 0.0;
    input value cutoff = 1800.0;
    input gate: event;
    output stream out;
    nodes {
        oscs = [PolyBlepOscillator::saw(220.0, 0.4); 3];
        envs = [AdsrEnvelope::new(0.002, 0.01, 0.6, 0.02); 3];
        vcas = [Gain::new(1.0); 3];
        lp = TptFilter::new(1800.0, 0.9);
    }
    connections {
        gate -> envs.gate;                       // Broadcast (event)
        detune -> oscs[1].frequency_mod;         // one element
        detune * -1.0 -> oscs[2].frequency_mod;
        oscs.output -> vcas.input;               // Parallel
        envs.output -> vcas.gain;
        vcas.output -> lp.input;                 // FanIn: 0.0 + v0 + v1 + v2
        cutoff -> lp.cutoff;
        lp.output + oscs[0].output * 0.1 -> out;
    }
"""


def test_node_arrays_broadcast_parallel_fanin_and_indexed_elements_match_the_oracle_walk(plan):
    n, blocks = 5, (70, 58)
    frames = sum(blocks)
    detune = (0.002 * np.arange(n)).astype(np.float32)
    cutoff = (900.0 + 500.0 * np.arange(n)).astype(np.float32)
    gates = [(v, 3 + 2 * v, 0.5 + 0.1 * v) for v in range(n)] + [(v, 80 + v, 0.0) for v in range(n)]
    eng = dsl_engine(UNISON, n, 1, SR, plan=plan)
    eng.set_value(0, detune); eng.set_value(1, cutoff)
    got, pos = [], 0
    for b in blocks:
        ev = [(v, f - pos, GATE, s) for v, f, s in gates if pos <= f < pos + b]
        if ev:
            eng.push_events(ev)
        eng.process_block(b); got.append(eng.read_voice_out()); pos += b
    got = np.concatenate(got)

    def make(v):
        oscs = [W.polyblep("saw", 220.0, 0.4, SR) for _ in range(3)]
        envs = [W.adsr(0.002, 0.01, 0.6, 0.02, SR) for _ in range(3)]
        lp = W.tpt(1800.0, 0.9, SR)

        def frame(vals, x):
            d = vals["detune"]
            oscs[1].set(frequency_mod=d); oscs[2].set(frequency_mod=f32(d * f32(-1.0)))
            o = [osc.process() for osc in oscs]
            e = [env.process() for env in envs]
            vc = [f32(o[k] * e[k]) for k in range(3)]
            s = f32(0.0)
            for k in range(3):                    # `let mut sum = 0.0; for i { sum += elt }` (codegen/emit_edge.rs:208-233)
                s = f32(s + vc[k])
            lp.set(input=s, cutoff=vals["cutoff"])
            return f32(lp.process() + f32(o[0] * f32(0.1)))

        def gate(s):
            for env in envs:
                env.on_gate(s)
        return frame, gate

    want = W.render(make, n, frames, {"detune": detune, "cutoff": cutoff}, gates)
    assert np.abs(want).max() > 0.05
    assert np.array_equal(got, want), float(np.abs(got - want).max())


ARRAY_4X = """
    input value level = 0.5;
    output stream out;
    nodes { emitters = [PolyBlepOscillator::square(330.0, 0.3); 3] * 4; shapers = [HardClip::new(); 3] * 4; post = TptFilter::new(6000.0, 0.7); }
    connections {
        level -> emitters.amplitude;             // Broadcast of a value into inner-rate nodes
        emitters.output -> shapers.input;        // Parallel, same (inner) rate
        [sinc_iir] shapers.output -> post.input; // FanIn across the rate boundary: summed at 4x, then ONE resampler
        post.output -> out;
    }
"""
UNROLLED_4X = """
    input value level = 0.5;
    output stream out;
    nodes {
        e0 = PolyBlepOscillator::square(330.0, 0.3) * 4; e1 = PolyBlepOscillator::square(330.0, 0.3) * 4; e2 = PolyBlepOscillator::square(330.0, 0.3) * 4;
        s0 = HardClip::new() * 4; s1 = HardClip::new() * 4; s2 = HardClip::new() * 4;
        post = TptFilter::new(6000.0, 0.7);
    }
    connections {
        level -> e0.amplitude; level -> e1.amplitude; level -> e2.amplitude;
        e0.output -> s0.input; e1.output -> s1.input; e2.output -> s2.input;
        [sinc_iir] 0.0 + s0.output + s1.output + s2.output -> post.input;
        post.output -> out;
    }
"""


def test_oversampled_node_array_renders_like_the_hand_unrolled_graph(fused_plan):
    """multirate_array_fanout.rs shapes on a `* 4` array: the unrolled scalar graph is what the oracle-checked multirate
    lowering already renders, so equality with it checks the array shapes (one resampler after the fan-in sum)"""
    n = 4
    a, b = dsl_engine(ARRAY_4X, n, 1, SR, plan=fused_plan), dsl_engine(UNROLLED_4X, n, 1, SR, plan=fused_plan)
    lv = (0.2 + 0.2 * np.arange(n)).astype(np.float32)
    outs = []
    for e in (a, b):
        e.set_value(0, lv)
        y = []
        for frames in (64, 50):
            e.process_block(frames); y.append(e.read_voice_out())
        outs.append(np.concatenate(y))
    assert np.abs(outs[0]).max() > 0.05
    assert np.array_equal(outs[0], outs[1])


def test_node_array_plans_and_errors():
    from oscen_b200 import abi
    d = abi.Graph(UNISON).describe()
    assert "AdsrEnvelope [envs[0], envs[1], envs[2]]" in d                  # one launch for the whole array (blockIdx.y = element)
    assert "PolyBlepOscillator::saw [oscs[0], oscs[1]]" in d and "PolyBlepOscillator::saw [oscs[2]]" in d   # oscs[2] waits for `detune * -1.0`
    assert "Expr [0 + vcas[0].output {a + v}]" in d                       # the fan-in starts from 0.0
    d4 = abi.Graph(ARRAY_4X).describe()
    assert d4.count("DownEdge [") == 1 and d4.count("UpEdge [") == 0    # value edges latch; one resampler after the sum
    # the reference's own array graphs (oscen-lib/tests/array_test.rs:5-56)
    g = abi.Graph("name: ArrayTest; output stream out; nodes { oscs = [PolyBlepOscillator::saw(440.0, 0.6); 4]; } connections { oscs[0].output -> out; }")
    assert g.launches() == ["0: PolyBlepOscillator::saw [oscs[0]]"]        # the unused elements are dead nodes
    g = abi.Graph("name: ArrayConnectionTest; input value freq = 440.0; output stream out; nodes { oscs = [PolyBlepOscillator::saw(440.0, 0.6); 4]; } "
                  "connections { freq -> oscs.frequency_mod; oscs[0].output + oscs[1].output + oscs[2].output + oscs[3].output -> out; }")
    assert g.launches()[0] == "0: PolyBlepOscillator::saw [oscs[0], oscs[1], oscs[2], oscs[3]]" and g.jit_check() > 10000
    # an array wider than one launch (4 nodes per launch) and a Parallel edge between arrays of different sizes: min(N, M)
    g = abi.Graph("output stream out; nodes { a = [PolyBlepOscillator::sine(100.0, 0.1); 6]; b = [Gain::new(0.5); 5]; } connections { a.output -> b.input; b.output -> out; }")
    d = g.describe()
    assert "a[4]" in d and "a[5]" not in d and "b[4]" in d
    for src, frag in [
        ("output stream out; nodes { o = [PolyBlepOscillator::saw(1.0, 1.0); 2]; } connections { o.output * 0.5 -> out; }", "a node array inside a source expression"),
        ("output stream out; nodes { o = [PolyBlepOscillator::saw(1.0, 1.0); 2]; } connections { o[2].output -> out; }", "cannot resolve the source"),
        ("output stream out; nodes { o = [PolyBlepOscillator::saw(1.0, 1.0); 2]; } connections { o[x].output -> out; }", "array-indexed endpoints take the form"),
        ("output stream out; nodes { o = [PolyBlepOscillator::saw(1.0, 1.0); 2]; g = Gain::new(1.0); } connections { o.output -> [3] -> g.input; g.output -> out; }",
         "feedback edges (`-> [..] ->`) take scalar endpoints"),
        ("output stream out; nodes { o = [PolyBlepOscillator::saw(1.0, 1.0); 65]; } connections { o.output -> out; }", "more than 64 elements"),
        ("output stream out; nodes { o = [PolyBlepOscillator::saw(1.0, 1.0); 2]; } connections { o -> out; }", "a node array inside a source expression"),
    ]:
        with pytest.raises(abi.GraphCompileError) as ei:
            abi.Graph(src)
        assert frag in str(ei.value), str(ei.value)


@pytest.mark.parametrize("seed", range(6))
def test_random_array_shapes_render_like_their_hand_unrolled_graphs(seed):
    """arrays of random sizes N, M (1..6) with a Broadcast value edge, an indexed element, a Parallel edge (min(N, M)
    elements) and a FanIn: the array text against the same graph written out node by node (CPU: host-executed kernel)"""
    from host_kernel import HostEngine
    rng = np.random.default_rng(4200 + seed)
    n_osc, n_gain = int(rng.integers(1, 7)), int(rng.integers(1, 7))
    k = int(rng.integers(0, n_osc))
    freqs = [float(np.float32(rng.uniform(100, 900))) for _ in range(2)]
    head = "input value spread = 0.01; input value level = 0.5; output stream out;\n"
    arr = head + f"""nodes {{ oscs = [PolyBlepOscillator::saw({freqs[0]!r}, 0.4); {n_osc}]; gains = [Gain::new(0.7); {n_gain}]; lp = TptFilter::new(2500.0, 0.8); }}
        connections {{
            spread -> oscs.frequency_mod;
            spread * 3.0 -> oscs[{k}].phase_mod;
            oscs.output -> gains.input;
            level -> gains.gain;
            gains.output -> lp.input;
            lp.output + oscs[{k}].output * 0.2 -> out;
        }}"""
    m = min(n_osc, n_gain)
    nodes = "".join(f"o{i} = PolyBlepOscillator::saw({freqs[0]!r}, 0.4); " for i in range(n_osc)) + \
        "".join(f"g{i} = Gain::new(0.7); " for i in range(n_gain)) + "lp = TptFilter::new(2500.0, 0.8);"
    conns = "".join(f"spread -> o{i}.frequency_mod; " for i in range(n_osc)) + f"spread * 3.0 -> o{k}.phase_mod; " + \
        "".join(f"o{i}.output -> g{i}.input; " for i in range(m)) + "".join(f"level -> g{i}.gain; " for i in range(n_gain)) + \
        "0.0" + "".join(f" + g{i}.output" for i in range(n_gain)) + " -> lp.input; " + f"lp.output + o{k}.output * 0.2 -> out;"
    flat = head + "nodes { " + nodes + " } connections { " + conns + " }"
    n = 3
    outs = []
    for src in (arr, flat):
        e = HostEngine(src, n, 1, SR)
        e.set_value(0, (0.01 * (1 + np.arange(n))).astype(np.float32)); e.set_value(1, (0.3 + 0.2 * np.arange(n)).astype(np.float32))
        y = []
        for frames in (40, 57):
            e.process_block(frames); y.append(e.read_voice_out())
        outs.append(np.concatenate(y))
    assert np.abs(outs[0]).max() > 0.01
    assert np.array_equal(outs[0], outs[1]), (n_osc, n_gain, k)
