"""CPU tests of the `graph!` DSL front end (oscen_b200/csrc/dsl.hpp through the C ABI): parsing of
both declaration spellings, lowering rules (kind compatibility, fan-in, last-write-wins, constructor
defaults, dead nodes), the launch plans of the reference graphs, and the compile errors — the
counterpart of the reference's compiler tests (oscen-graph-compiler/tests/{lower,accumulation,
expr_lowering}.rs) and compile-fail UI tests (oscen-macros/tests/ui/*.rs).  No GPU needed."""
import os
import re

import pytest

from oscen_b200 import abi

GRAPHS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "graphs")


def text(name):
    with open(os.path.join(GRAPHS, name)) as f:
        return f.read()


def groups(lines):
    """launch lines -> list of node-name lists (type label dropped)"""
    out = []
    for l in lines:
        names = re.search(r"\[(.*)\]", l).group(1)
        out.append([re.sub(r" \{.*\}$", "", n.strip()) for n in names.split(", ")])
    return out


def builtin_plan(kind):
    import ctypes as C
    L = abi.load()
    buf = C.create_string_buffer(4096)
    n = L.oscen_b200_plan_describe(kind, buf, len(buf))
    assert n > 0
    return buf.value.decode().strip().splitlines()


@pytest.mark.parametrize("file,kind", [("voice_fm3.graph", abi.GRAPH_FM_VOICE), ("synth.graph", abi.GRAPH_SYNTH),
                                       ("synth_adsr.graph", abi.GRAPH_SYNTH_ADSR), ("sat_1x.graph", abi.GRAPH_SAT_1X)])
def test_reference_graphs_compile_to_the_builtin_launch_plans(file, kind):
    g = abi.Graph(text(file))
    assert g.launches() == builtin_plan(kind)


def test_twin_peaks_plan_and_expression_edge():
    g = abi.Graph(text("twin_peaks.graph"))
    want = groups(builtin_plan(abi.GRAPH_TWIN_PEAKS))
    assert groups(g.launches()) == want == [["filter_a", "filter_b"], ["filter_a.output - filter_b.output"]]
    d = g.describe()
    assert "stream input: audio_in" in d and "{a - b}" in d and "state words per voice: 18" in d


def test_fm_voice_inputs_follow_declaration_order():
    g = abi.Graph(text("voice_fm3.graph"))
    assert g.n_inputs == 28
    assert g.input_id("frequency") == 0 and g.input_id("op3_ratio") == 1 and g.input_id("route") == 20
    assert g.input_id("filter_env_amount") == 27
    assert g.input_info(21) == {"name": "filter_cutoff", "default": 2000.0, "range": None, "log": False, "ramp": 0}
    with pytest.raises(KeyError):
        g.input_id("gate")                      # the event input is not a value input
    d = g.describe()
    assert "event input: gate" in d and "state words per voice: 29" in d
    # the four envelopes share one launch, so do op-independent stateless nodes of one wave
    assert g.launches()[0] == "0: AdsrEnvelope [env3, env2, env1, env_filter]"


def test_param_spec_forms():
    g = abi.Graph("""
        input cutoff: value = 3000.0 [20.0..20000.0, log, ramp: 1323];
        input value q = 0.707 [0.1..10.0];
        input volume: value = 0.8 { range: 0.0..1.0, unit: " dB", smoother: 50.0, name: "Vol", group: "Out", ramp: 64 };
        input detune: value = -0.5 [-1.0..1.0, linear];
        input big: value = 2_000.5f32;
        input tiny: value = 1e-3;
        output stream out;
        node f = TptFilter::new(1000.0, 0.7);
        node o = PolyBlepOscillator::saw(220.0, 1.0);
        connection o.output -> f.input; connection cutoff -> f.cutoff; connection q -> f.q;
        connection f.output * volume -> out;
    """)
    assert g.input_info(0) == {"name": "cutoff", "default": 3000.0, "range": (20.0, 20000.0), "log": True, "ramp": 1323}
    assert g.input_info(1)["range"] == (pytest.approx(0.1), 10.0) and g.input_info(1)["ramp"] == 0
    assert g.input_info(2)["range"] == (0.0, 1.0) and g.input_info(2)["ramp"] == 64
    assert g.input_info(3)["default"] == -0.5 and g.input_info(3)["log"] is False
    assert g.input_info(4)["default"] == 2000.5 and g.input_info(5)["default"] == pytest.approx(1e-3)


def test_fan_in_sums_in_statement_order_and_dead_nodes_are_dropped():
    g = abi.Graph("""
        name: FanIn;
        input bias: value = 0.25;
        output out: stream;
        nodes {
            a = PolyBlepOscillator::sine(100.0, 1.0);
            b = PolyBlepOscillator::saw(200.0, 1.0);
            c = PolyBlepOscillator::square(300.0, 1.0);
            unused = PolyBlepOscillator::triangle(50.0, 1.0);
            sink = Gain::new(1.0);
            f = TptFilter::new(500.0, 0.7);
        }
        connections {
            a.output -> f.input;
            b.output -> f.input;
            c.output -> f.input;
            bias -> f.input;
            unused.output -> sink.input;
            f.output -> out;
            a.output -> out;
        }
    """)
    d = g.describe()
    assert "unused" not in d and "sink" not in d
    gr = groups(g.launches())
    # ((a + b) + c) + bias into the filter: two Mixer launches (stream + stream) then stream + value
    flat = [n for grp in gr for n in grp]
    assert flat.count("f.input (fan-in)") == 3 and flat.count("out (fan-in)") == 1
    assert flat.index("a") < flat.index("f.input (fan-in)") < flat.index("f") < flat.index("out (fan-in)")
    assert "state words per voice: 10" in d           # 3 oscillator phases + TptFilter(7); the dead oscillator holds none


def test_value_endpoint_last_write_wins_and_constructor_defaults():
    g = abi.Graph("""
        input x: value = 1.0;
        input y: value = 2.0;
        output stream out;
        node o = PolyBlepOscillator::saw(220.0, 0.5);
        connection x -> o.frequency;
        connection y -> o.frequency;
        connection 0.25 -> o.amplitude;
        connection o.output -> out;
    """)
    assert g.launches() == ["0: PolyBlepOscillator::saw [o]"]


ONE_OSC = "output stream out; node o = PolyBlepOscillator::saw(2000.0, 0.6); "


@pytest.mark.parametrize("src,frag", [
    ("output stream out; node x = Foo::new(); connection x.output -> out;", "unknown node type `Foo`"),
    ("output stream out; node g = Gain::new(1.0); connection g.output -> g.input; connection g.output -> out;", "cycle in graph"),
    ("input gate: event; output stream out; node f = TptFilter::new(100.0, 0.7); connection gate -> f.input; connection f.output -> out;",
     "Type mismatch in connection: source is Event but destination expects Stream"),
    ("output stream out; node o = PolyBlepOscillator::saw(1.0, 1.0); node e = AdsrEnvelope::new(0.1, 0.1, 0.5, 0.1); "
     "connection o.output -> e.attack; connection e.output -> out;",
     "Type mismatch in connection: source is Stream but destination expects Value"),
    ("input v: value = 1.0; output stream out; node e = AdsrEnvelope::new(0.1, 0.1, 0.5, 0.1); connection v -> e.gate; connection e.output -> out;",
     "Type mismatch in connection: source is Value but destination expects Event"),
    ("output stream out; node o = PolyBlepOscillator::saw(2_000.0, 0.6) / 2; connection o.output -> out;", "node undersampling (`/ N`) is not yet supported"),
    ("output stream out; node a = PolyBlepOscillator::saw(1.0, 1.0) * 4; node b = Gain::new(1.0) * 2; connection a.output -> out;", "one oversampling factor per graph"),
    ("output stream out; node a = PolyBlepOscillator::saw(1.0, 1.0) * 4; node b = PolyBlepOscillator::saw(2.0, 1.0); connection a.output + b.output -> out;",
     "arithmetic between streams of different rates"),
    ("output stream out; node a = PolyBlepOscillator::saw(1.0, 1.0) * 4; node b = PolyBlepOscillator::saw(2.0, 1.0) * 4; connection a.output -> out; connection b.output -> out;",
     "fan-in summing supports only same-rate"),
    ("output stream out; node a = PolyBlepOscillator::saw(1.0, 1.0) * 2; node f = TptFilter::new(500.0, 0.7); node c = HardClip::new() * 2; "
     "connection a.output -> f.input; connection f.output -> c.input; connection c.output -> out;", "cannot feed it again in the same frame"),
    ("output stream out; node o = PolyBlepOscillator::saw(2000.0, 0.6) * 3; connection o.output -> out;", "rate factor must be in {1, 2, 4, 8}"),
    ("output stream out; node v = [UserVoice::new(); 4]; connection v.output -> out;", "is not a graph of this text"),
    ("output stream out; node o = PolyBlepOscillator::saw(2000.0, 0.6)\n connection o.output -> out;", "2:2: expected `;`"),
    (ONE_OSC + "connection o.output -> [4] -> out;", "must end at a node input"),
    (ONE_OSC + "node g = Gain::new(1.0); connection o.output -> [g] -> g.input; connection g.output -> out;", "does not implement AllowsFeedback (only Delay does)"),
    (ONE_OSC + "node g = Gain::new(1.0); connection o.output -> [nowhere] -> g.input; connection g.output -> out;", "no such node"),
    (ONE_OSC + "node g = Gain::new(1.0); connection o.output -> [1.5] -> g.input; connection g.output -> out;", "must be an integer literal"),
    (ONE_OSC + "connection o.outptu -> out;", "node `o` has no endpoint `outptu`"),
    (ONE_OSC + "connection osc.output -> out;", "cannot resolve the source"),
    (ONE_OSC + "connection o.output -> nowhere;", "cannot resolve the destination"),
    (ONE_OSC + "connection o.frequency -> out;", "is an input: it cannot be a connection source"),
    (ONE_OSC + "node p = PolyBlepOscillator::saw(1.0, 1.0); connection p.output -> o.output; connection o.output -> out;",
     "is an output: it cannot be a connection destination"),
    ("output stream out; node o = PolyBlepOscillator::saw(2000.0); connection o.output -> out;", "takes 2 argument(s), found 1"),
    ("output stream out; node o = PolyBlepOscillator::pulse(2000.0, 1.0); connection o.output -> out;", "unknown constructor `PolyBlepOscillator::pulse`"),
    (ONE_OSC + "node o = Gain::new(1.0); connection o.output -> out;", "duplicate name `o`"),
    (ONE_OSC + "connection half(o.output) -> out;", "free function `half`"),
    (ONE_OSC + "connection o.output.powf(2.0) -> out;", "method `.powf()` is not supported"),
    ("node o = PolyBlepOscillator::saw(2000.0, 0.6);", "declares no stream output"),
    (ONE_OSC, "nothing is connected to the output"),
    ("output stream out; output stream out2; node o = PolyBlepOscillator::saw(1.0, 1.0); connection o.output -> out;", "one stream output per graph"),
    ("output stream out; input a: stream; input b: stream; connection a -> out;", "one stream input per graph"),
    ("input audio: stream: Frame<2>; output stream out; connection audio -> out;", "typed endpoints"),
    ("output stream out; external ir: ImpulseResponse; node o = PolyBlepOscillator::saw(1.0, 1.0); connection o.output -> out;", "only AudioAsset slots are supported (found `ImpulseResponse`)"),
    ("output stream out; external s: AudioAsset; node o = PolyBlepOscillator::saw(1.0, 1.0); connection s -> o.frequency; connection o.output -> out;",
     "Type mismatch in connection: source is Asset but destination expects Value"),
    ("output stream out; external s: AudioAsset; node p = SamplePlayer::new(); connection s -> out;", "an external asset connects to a node's asset endpoint"),
    ("output stream out; external s: AudioAsset; node p = SamplePlayer::new(); connection s * 2.0 -> p.buf; connection p.output -> out;", "is an external asset slot: it connects to an asset endpoint only"),
    (ONE_OSC + "node p = SamplePlayer::new(); connection o.output -> p.buf; connection p.output -> out;", "Type mismatch in connection: source is Stream but destination expects Asset"),
    ("output stream out; external s: AudioAsset; external s: AudioAsset; node p = SamplePlayer::new(); connection p.output -> out;", "duplicate name `s`"),
    ("output stream out; external a: AudioAsset; external b: AudioAsset; external c: AudioAsset; external d: AudioAsset; external e: AudioAsset; "
     "node p = SamplePlayer::new(); connection p.output -> out;", "more than 4 external asset slots"),
    ("output stream out; node p = SamplePlayer::new(1.0); connection p.output -> out;", "SamplePlayer::new takes 0 argument(s), found 1"),
    ("output stream out; name: Late; node o = PolyBlepOscillator::saw(1.0, 1.0); connection o.output -> out;", "`name:` must be the first item"),
    (ONE_OSC + "connection [cubic] o.output -> out;", "unknown connection policy `cubic`"),
    ("output stream out; /* never closed", "unterminated block comment"),
    (ONE_OSC + "connection o.output -> out; @", "unexpected character `@`"),
    # run-time text must not be able to overflow the stack or a fixed buffer
    (ONE_OSC + "connection " + "(" * 200000 + "o.output" + ")" * 200000 + " -> out;", "nested deeper than 64 parentheses"),
    (ONE_OSC + "connection o.output" + " + o.output" * 5000 + " -> out;", "expression too long"),
    (ONE_OSC + "connection o.output" + ".abs()" * 5000 + " -> out;", "expression too long"),
    ("input " + "c" * 300 + ": value = 1.0; " + ONE_OSC + "connection o.output -> out;", "identifier longer than 63 characters"),
])
def test_compile_errors(src, frag):
    with pytest.raises(abi.GraphCompileError) as ei:
        abi.Graph(src)
    msg = str(ei.value)
    assert frag in msg, msg
    assert re.match(r"^\d+:\d+: ", msg), msg          # every diagnostic carries line:col


def test_longest_identifiers_survive_code_generation():
    """63-character names (the lexer's cap) reach the generated kernel's text whole: every `/* name */` comment is closed and
    every input is loaded (a 300-character name used to truncate a line inside its comment and swallow the next)"""
    a, b = "a" * 63, "b" * 63
    g = abi.Graph(f"input {a}: value = 100.0; input {b}: value = 0.5; output stream out; "
                  f"node o = PolyBlepOscillator::saw(1.0, 1.0); connection {a} -> o.frequency; connection {b} -> o.amplitude; connection o.output -> out;")
    src = g.codegen()
    assert src.count("/*") == src.count("*/")
    assert f"/* {a} */" in src and f"/* {b} */" in src and "P[0] = a.pv[" in src and "P[1] = a.pv[" in src
    assert g.input_info(1)["name"] == b
    assert g.jit_check() > 0


def test_expressions_lower_to_one_operator_per_launch():
    g = abi.Graph("""
        input gain: value = 0.5;
        input offset: value = 0.1;
        output stream out;
        node osc = PolyBlepOscillator::sine(440.0, 1.0);
        node lfo = PolyBlepOscillator::sine(2.0, 1.0);
        connection (osc.output * gain + offset) / (lfo.output + 2.0) - 1.0 / lfo.output.abs().max(0.25) -> out;
    """)
    ops = re.findall(r"\{([^}]*)\}", g.describe())
    assert sorted(ops) == sorted(["a * v", "a + v", "a + v", "a / b", "abs", "max", "v / a", "a - b"])
    # literal (op) literal is folded with f32 arithmetic
    g2 = abi.Graph("output stream out; node o = PolyBlepOscillator::sine(1.0, 1.0); connection o.output * (0.5 * 0.25) -> out;")
    assert re.findall(r"\{([^}]*)\}", g2.describe()) == ["a * v"]


def test_graph_objects_are_independent_of_the_gpu():
    # compiling needs no device; creating an engine from it does (and fails loudly without one)
    g = abi.Graph(text("synth.graph"))
    import conftest
    if not conftest.HAS_GPU:
        with pytest.raises(abi.EngineError) as ei:
            abi.Engine.from_graph(g, 64)
        assert "no CUDA device" in str(ei.value)


# ------------------------------------------------------------------ generated fused kernels (codegen.hpp + NVRTC)
@pytest.mark.parametrize("file", ["voice_fm3.graph", "synth.graph", "synth_adsr.graph", "twin_peaks.graph", "sat_1x.graph"])
def test_generated_kernel_source_compiles_for_sm_100a(file):
    """NVRTC needs no GPU: every reference graph's generated kernel must compile for sm_100a here"""
    g = abi.Graph(text(file))
    src = g.codegen()
    assert 'extern "C" __global__ void __launch_bounds__(kCta, 8) oscen_b200_generated_kernel' in src
    assert '#include "state_io.cuh"' in src and "tile_flush_mono" in src
    assert g.jit_check() > 10000           # cubin bytes


def test_generated_kernel_walks_nodes_in_topological_order():
    g = abi.Graph(text("voice_fm3.graph"))
    src = g.codegen()
    order = [m.group(1) for m in re.finditer(r"const float s\d+ = .*/\* (\w+) \*/$", src, re.M)]
    pos = {n: i for i, n in enumerate(order)}
    # every edge of fm_voice.rs: producer before consumer
    for a, b in [("env3", "op3_osc"), ("op3_osc", "op3_route"), ("op3_route", "op2_osc"), ("op2_osc", "op1_mod_mixer"),
                 ("op1_mod_mixer", "op1_osc"), ("op1_osc", "filter"), ("env_filter", "filter_env_gain"),
                 ("filter_env_gain", "cutoff_mod"), ("cutoff_mod", "filter"), ("filter", "output_gain")]:
        assert pos[a] < pos[b], (a, b)
    # the walk is emitted three times: the plain-segment copy (envelopes in their segment form), its twin that loads the frame's
    # ramp row first (ramp blocks whose ramps stay off the envelopes' and the operators' hoisted state), and the general frame
    assert src.count("adsr_on_gate(") == 4 and src.count("fmop_process_k(") == 9 and src.count("adsr_ad(") == 8
    # hoisted state sits behind frequency, the three ratios and the sixteen envelope times / sustain levels
    hoisted = [0, 1, 8, 15] + [4, 5, 6, 7, 11, 12, 13, 14, 16, 17, 18, 19, 23, 24, 25, 26]
    assert "(ramp_mask & 0x%xu) == 0u" % sum(1 << i for i in hoisted) in src
    assert "if ((ramp_mask >> 21) & 1u) P[21] = row[21];" in src and src.count("P[21] = row[21]") == 2 and src.count("P[4] = row[4]") == 1
    assert src.count("__reduce_min_sync(") == 1 and "oscen::umin(np, adsr_plain_frames(n0))" in src
    # literals are emitted bit-exactly
    assert "0x3e99999au /* 0.300000012 */" in src      # output_gain = Gain::new(0.3)


def test_generated_kernel_for_expression_graph():
    g = abi.Graph("""
        input gain: value = 0.5;
        input gate: event;
        output stream out;
        node osc = PolyBlepOscillator::triangle(440.0, 1.0);
        node env = AdsrEnvelope::new(0.01, 0.1, 0.5, 0.2);
        node idle = AdsrEnvelope::new(0.01, 0.1, 0.5, 0.2);
        connection gate -> env.gate;
        connection ((osc.output * gain).tanh() / 2.0 + idle.output).clamp(-0.5, 0.5) * env.output -> out;
    """)
    src = g.codegen()
    assert "om_tanhf_u(" in src and "clampf(" in src and src.count("adsr_on_gate(") == 1
    assert g.jit_check() > 10000


# ------------------------------------------------------------------ multirate graphs (`* N` nodes, cross-rate edges)
def test_multirate_lowering_inserts_one_resampler_per_cross_rate_edge():
    g = abi.Graph(text("sat_4x.graph"))
    d = g.describe()
    assert g.launches() == ["0: PolyBlepOscillator::saw [osc]", "1: HardClip [clip]", "2: DownEdge [clip.output -> audio_out [sinc down x4]]"]
    assert "state words per voice: 49" in d                    # phase + SincDownFir<4>: 2 stages x 24 taps (== the built-in SAT layout)
    src = g.codegen()
    assert "DownEdge<0, 2>" in src and "const float sri = sr * 4.0f;" in src and 'for (int inner = 0; inner < 4; ++inner)' in src
    assert src.count(", sri);") == 3                           # only the oscillator sees the inner sample rate (in the three copies of the walk: plain segment, ramp segment, general frame)
    assert g.jit_check() > 10000


@pytest.mark.parametrize("policy,words", [("sinc", 2 * 12 + 2 * 24), ("sinc_iir", 2 * 8 + 2 * 9), ("linear", 1), ("latch", 0)])
def test_multirate_policies_default_and_state_words(policy, words):
    prefix = f"[{policy}] " if policy != "sinc" else ""          # stream edges default to [sinc] (ir/lower.rs refine_kernels)
    g = abi.Graph(f"""
        input audio_in: stream; input drive: value = 1.0; output stream out;
        node g = Gain::new(1.0) * 4;
        node post = Gain::new(0.5);
        connection {prefix}audio_in * drive -> g.input;           // the compound source is evaluated at the outer rate, then upsampled
        connection {prefix}g.output.clamp(-0.7, 0.7) -> post.input;
        connection post.output -> out;
    """)
    d = g.describe()
    assert f"[{policy} up x4]" in d and f"[{policy} down x4]" in d
    assert f"state words per voice: {words}" in d
    src = g.codegen()
    # phases: outer product -> Up warm-up -> 4 inner ticks (gain, clamp, capture) -> Down finalize -> post gain
    order = [src.rindex(k) for k in ("x_in * P[0]", ".up(", "for (int inner = 0", "clampf(", ".down(", "/* post */")]
    assert order == sorted(order)
    assert g.jit_check() > 10000


def test_value_edges_into_inner_nodes_latch():
    """a value input read by an inner-rate node is a plain field read on every inner tick (Latch, lower.rs refine_kernels)"""
    g = abi.Graph("""
        input f: value = 440.0; input gate: event; output stream out;
        node osc = PolyBlepOscillator::sine(440.0, 1.0) * 2;
        node env = AdsrEnvelope::new(0.01, 0.1, 0.5, 0.1) * 2;
        node vca = Gain::new(1.0) * 2;
        connection f -> osc.frequency; connection gate -> env.gate;
        connection osc.output -> vca.input; connection env.output -> vca.gain;
        connection [sinc_iir] vca.output -> out;
    """)
    src = g.codegen()
    assert "adsr_derive_r(d1" in src and ", sri);" in src and "UpEdge" not in src
    assert g.jit_check() > 10000


def test_stringify_spacing_is_accepted():
    """Rust's stringify!() of the macro tokens separates most tokens with spaces: the front end lexes tokens, not text"""
    spaced = ("name : Synth ; input cutoff : value = 1200.0 [20.0 .. 20000.0 , log , ramp : 1323] ; output audio_out : stream ; "
              "nodes { osc = PolyBlepOscillator :: saw (2_000.0 , 0.5) * 2 ; filter = TptFilter :: new (1200.0 , 0.707) ; } "
              "connections { [sinc_iir] osc . output -> filter . input ; cutoff -> filter . cutoff ; filter . output . tanh () * 0.5 -> audio_out ; }")
    g = abi.Graph(spaced)
    assert g.input_info(0)["ramp"] == 1323 and "[sinc_iir down x2]" in g.describe() and "{tanh}" in g.describe()


# ------------------------------------------------------------------ polyphonic graphs (voice arrays + the MIDI front)
def test_polyphonic_fm_graph_compiles_to_the_voice_plan_with_outer_names():
    g = abi.Graph(text("voice_fm3.graph") + text("poly_fm8.graph"))
    assert g.voices_per_instance == 8 and g.has_midi
    assert g.launches() == builtin_plan(abi.GRAPH_FM_VOICE)
    d = g.describe()
    assert "graph FMGraph" in d and "polyphony: 8 voices per instance, MIDI front -> frequency = input 0, gate" in d
    assert "sum of 8 x FMVoice.audio_out" in d
    # ids are the voice graph's (OSCEN_B200_FM_*); names and ramps come from the outer declarations
    assert g.input_id("op3_ratio") == 1 and g.input_id("filter_cutoff") == 21 and g.input_id("frequency") == 0
    assert g.input_info(21)["ramp"] == 2205 and g.input_info(21)["range"] == (20.0, 20000.0)
    assert g.input_info(2)["ramp"] == 2205 and g.input_info(1)["ramp"] == 0
    assert g.jit_check() > 10000


@pytest.mark.parametrize("mutation,frag", [
    (("voices = [FMVoice::new(); 8];", "voices = [FMVoice::new(); 6];"), "power of two"),
    (("voice_handlers = [MidiVoiceHandler::new(); 8];", "voice_handlers = [MidiVoiceHandler::new(); 4];"), "same array size"),
    (("voices = [FMVoice::new(); 8];", "voices = [Unknown::new(); 8];"), "is not a graph of this text"),
    (("voice_handlers.gate -> voices.gate;", "voice_handlers.gate -> voices.trigger;"), "has no event input `trigger`"),
    (("route -> voices.route;", "route -> voices.rout;"), "has no value input `rout`"),
    (("connection voices.audio_out -> audio_out;", ""), "is not connected to the output"),
    (("node midi_parser = MidiParser::new();", ""), "MIDI front needs all of"),
    (("route -> voices.route;", "route -> voices.route; connection op3_ratio -> midi_parser.midi_in;"), "not part of the polyphonic pattern"),
])
def test_polyphonic_graph_errors(mutation, frag):
    outer = text("poly_fm8.graph")
    assert mutation[0] in outer
    with pytest.raises(abi.GraphCompileError) as ei:
        abi.Graph(text("voice_fm3.graph") + outer.replace(mutation[0], mutation[1]))
    assert frag in str(ei.value), str(ei.value)


# ------------------------------------------------------------------ array-typed streams, Frame<2>, the instance-level tail
def test_electric_piano_graph_compiles_from_the_reference_text():
    """examples/electric-piano: `amplitudes: [f32; 32]` between AmplitudeSource and OscillatorBank (electric_piano_voice.rs:362-402),
    16 voices -> sum -> Tremolo -> `output out: stream: Frame<2>` with event outputs tapped off the parser (main.rs:33-97)"""
    g = abi.Graph(text("electric_piano.graph"))
    assert g.voices_per_instance == 16 and g.has_midi and g.channels == 2
    d = g.describe()
    assert "graph ElectricPianoGraph" in d and "state words per voice: 259" in d
    assert "out <- Tremolo (Frame<2>) <- sum of 16 x ElectricPianoVoiceNode.output" in d
    # the voice's inputs keep the hand-written kernel's ids (OSCEN_B200_EP_*); the tremolo's follow them
    assert [g.input_id(k) for k in ("frequency", "brightness", "velocity_scaling", "decay_rate", "harmonic_decay", "key_scaling",
                                    "release_rate", "vibrato_intensity", "vibrato_speed")] == list(range(9))
    assert g.input_info(7)["default"] == pytest.approx(0.3) and g.input_info(8)["default"] == 5.0
    assert [l.split(":")[1].split("[")[0].strip() for l in g.launches()] == ["AmplitudeSource", "OscillatorBank"]
    src = g.codegen()
    assert "float s0[32]; ampsrc_process(" in src and "oscbank_process(" in src and "tile_flush_tremolo(a, tile, isum2, f0, fe, trem_phase, 8u," in src
    assert g.jit_check() > 10000
    voice = abi.Graph(text("electric_piano.graph"), name="ElectricPianoVoiceNode")      # the voice graph alone: mono, 1 voice per instance
    assert voice.channels == 1 and voice.voices_per_instance == 1 and voice.jit_check() > 10000
    with pytest.raises(abi.EngineError) as ei:                                            # refused before any device is touched
        abi.Engine.from_graph(g, 64, plan=abi.PLAN_NODES)
    assert "array-typed streams" in str(ei.value) and "PLAN_FUSED" in str(ei.value)


@pytest.mark.parametrize("mutation,frag", [
    (("amplitude_source.amplitudes -> oscillator_bank.amplitudes;", "amplitude_source.amplitudes -> output;"), "source is [f32; 32] but the graph output is f32"),
    (("amplitude_source.amplitudes -> oscillator_bank.amplitudes;", "oscillator_bank.output -> oscillator_bank.amplitudes;"), "source is f32 but `oscillator_bank.amplitudes` expects [f32; 32]"),
    (("amplitude_source.amplitudes -> oscillator_bank.amplitudes;", "amplitude_source.amplitudes * 2.0 -> oscillator_bank.amplitudes;"), "connects one output to one input directly"),
    (("amplitude_source.amplitudes -> oscillator_bank.amplitudes;",
      "amplitude_source.amplitudes -> oscillator_bank.amplitudes; amplitude_source.amplitudes -> oscillator_bank.amplitudes;"), "fan-in summing supports only scalar"),
    (("output out: stream: Frame<2>;", "output out: stream;"), "`Tremolo.output` is Frame<2> but `out` is f32"),
    (("output out: stream: Frame<2>;", "output out: stream: Frame<3>;"), "mono (`f32`, `Frame<1>`) or stereo (`Frame<2>`)"),
    (("output out: stream: Frame<2>;", "output out: stream: Complex;"), "unknown endpoint type `Complex`"),
    (("tremolo.output -> out;", ""), "must sit between the voice sum and the output"),
    (("vibrato_speed -> tremolo.rate;", "vibrato_speed -> tremolo.speed;"), "has no endpoint `speed`"),
    (("tremolo = crate::tremolo::Tremolo::new();", "tremolo = crate::tremolo::Tremolo::new(); chorus = crate::tremolo::Tremolo::new();"), "at most one Tremolo"),
    (("midi_parser.note_on -> note_on_out;", "midi_parser.velocity -> note_on_out;"), "MidiParser has no output `velocity`"),
    (("output output: stream;", "output output: stream: Frame<2>;"), "needs a node that produces one"),
])
def test_array_stream_and_frame_output_errors(mutation, frag):
    src = text("electric_piano.graph")
    assert mutation[0] in src
    with pytest.raises(abi.GraphCompileError) as ei:
        abi.Graph(src.replace(mutation[0], mutation[1]))
    assert frag in str(ei.value), str(ei.value)


# ------------------------------------------------------------------ feedback edges and the rest of the node library
FEEDBACK = """
    name: Echo;
    input time: value = 480.0;
    input fb: value = 0.4;
    output stream out;
    nodes {
        osc = Oscillator::saw(220.0, 0.5);
        echo = Delay::new(480.0, 0.0);
        lp = IirLowpass::new(3000.0, 0.707);
        tap = Gain::new(0.5);
        sum = Mixer::new();
    }
    connections {
        osc.output -> sum.input_a;
        lp.output * fb -> [echo] -> sum.input_b;
        time -> echo.delay_samples;
        sum.output -> lp.input;
        sum.output -> [1] -> tap.input;
        sum.output + tap.output -> out;
    }
"""


def test_feedback_edges_close_cycles_through_delays():
    g = abi.Graph(FEEDBACK)
    flat = [n for grp in groups(g.launches()) for n in grp]
    # the legs OUT of the delays do not order the walk: sum and tap run before the delays they read
    assert flat.index("sum") < flat.index("lp") < flat.index("lp.output * fb") < flat.index("echo")
    assert flat.index("tap") < flat.index("sum.output -> [1]")
    # ring sizes: `time` is connected, so `echo` keeps the reference capacity (131072 at 48 kHz); `-> [1] ->` needs 2 slots
    assert "state words per voice: %d" % (1 + 8 + (3 + 131072) + (3 + 2)) in g.describe()
    src = g.codegen()
    assert "delay_process<true, false>" in src and "delay_process<false, false>" in src and "iirlp_process(" in src and "osc_process<OSC_SAW>" in src
    assert g.jit_check() > 10000
    # without the via the same wiring is a cycle
    with pytest.raises(abi.GraphCompileError) as ei:
        abi.Graph(FEEDBACK.replace("-> [echo] -> sum.input_b", "-> sum.input_b"))
    assert "cycle in graph" in str(ei.value)


def test_constant_whole_sample_delays_get_small_rings():
    g = abi.Graph("""
        output stream out;
        node o = Oscillator::sine(100.0, 1.0);
        node d = Delay::new(11025.0, 0.3);              // examples/simple-echo/src/lib.rs:18
        node e = Delay::from_seconds(0.25, 0.0, 48000.0);
        node f = Delay::new(10.5, 0.0);                 // fractional: cubic reads depend on the reference capacity
        connection o.output -> d.input; connection o.output -> e.input; connection o.output -> f.input;
        connection d.output + e.output + f.output -> out;
    """)
    assert "state words per voice: %d" % (1 + (3 + 16384) + (3 + 16384) + (3 + 131072)) in g.describe()


def test_node_plan_refuses_feedback_and_multirate_graphs():
    g = abi.Graph(FEEDBACK)
    with pytest.raises(abi.EngineError) as ei:
        abi.Engine.from_graph(g, 64, plan=abi.PLAN_NODES)
    assert "needs the per-sample walk" in str(ei.value)
    with pytest.raises(abi.EngineError) as ei:
        abi.Engine.from_graph(abi.Graph(text("sat_4x.graph")), 64, plan=abi.PLAN_NODES)
    assert "multirate" in str(ei.value)


def test_external_asset_slot_and_sample_player_plan():
    """`external NAME: AudioAsset;` + `NAME -> player.buf` (oscen-lib/tests/sample_player_graph.rs:15-31): the slot table, the
    one-word player state, the generated kernel compiles for sm_100a, and the node plan has a launch for the player"""
    g = abi.Graph(text("player.graph"))
    d = g.describe()
    assert "external 0: clip (AudioAsset) -> deck.buf" in d and "state words per voice: 1" in d
    assert g.launches() == ["0: SamplePlayer [deck]"]
    assert g.n_assets() == 1 and g.asset_id("clip") == 0
    with pytest.raises(KeyError):
        g.asset_id("nope")
    src = g.codegen()
    assert "sample_player_process(" in src and "a.assets->reset[0]" in src
    assert g.jit_check() > 10000
    # a slot nobody reads and a player nobody feeds are both legal (the reference emits silence for the latter)
    g2 = abi.Graph("output stream out; external spare: AudioAsset; node p = SamplePlayer::new() * 2; connection p.output -> out;")
    assert "external 0: spare (AudioAsset) -> (not connected)" in g2.describe()
    assert "`buf` is not connected: silence" in g2.codegen()
    assert g2.jit_check() > 10000


# ------------------------------------------------------------------ the front end under sanitizers + fuzzing
@pytest.fixture(scope="module")
def graphc(tmp_path_factory):
    """oscen_b200/tools/graphc.cpp (the front end + code generator as a CLI) built with ASan + UBSan"""
    import subprocess
    exe = str(tmp_path_factory.mktemp("graphc") / "graphc")
    root = os.path.dirname(GRAPHS[:-len("/graphs")])
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                           "-o", exe, os.path.join(root, "oscen_b200", "tools", "graphc.cpp")])
    return exe


def _run(exe, args, text=None):
    import subprocess
    return subprocess.run([exe] + args, input=text, capture_output=True, text=True, timeout=60)


def test_graphc_cli_matches_the_library_and_is_sanitizer_clean(graphc):
    for files in (["voice_fm3.graph"], ["synth.graph"], ["twin_peaks.graph"], ["sat_4x.graph"], ["voice_fm3.graph", "poly_fm8.graph"]):
        paths = [os.path.join(GRAPHS, f) for f in files]
        r = _run(graphc, ["--plan"] + paths)
        assert r.returncode == 0 and r.stderr == "", r.stderr
        lib = abi.Graph("".join(text(f) + "\n" for f in files))
        assert r.stdout == lib.describe()
        c = _run(graphc, ["--cuda"] + paths)
        assert c.returncode == 0 and c.stderr == "" and c.stdout == lib.codegen()
    bad = _run(graphc, ["--check", "-"], "output stream out; node x = Foo::new(); connection x.output -> out;")
    assert bad.returncode == 1 and bad.stderr.startswith("1:25: unknown node type `Foo`")
    assert _run(graphc, []).returncode == 2


def test_front_end_survives_mutated_graphs(graphc):
    """token-level fuzzing of every test graph (delete / duplicate / swap / replace tokens, truncate): the front end must
    answer with a plan or a `line:col:` diagnostic — never crash, hang, or trip ASan / UBSan"""
    import random
    import re as _re
    rng = random.Random(20261017)
    corpus = [text(f) for f in ("voice_fm3.graph", "synth_adsr.graph", "twin_peaks.graph", "sat_4x.graph")]
    corpus.append(text("voice_fm3.graph") + text("poly_fm8.graph"))
    corpus.append(FEEDBACK)
    corpus.append(text("player.graph"))                       # `external` slots, SamplePlayer
    corpus.append("input value f = 1.0; input gate: event; output stream out; nodes { o = [PolyBlepOscillator::saw(440.0, 0.6); 4] * 2; "
                  "e = [AdsrEnvelope::new(0.01, 0.1, 0.5, 0.2); 4]; v = [Gain::new(1.0); 3] * 2; } connections { gate -> e.gate; f -> o.frequency_mod; "
                  "o.output -> v.input; [linear] e[1].output * -0.5 -> v[0].gain; [sinc] v.output -> out; }")      # node arrays, indices, negative literals
    junk = ["->", "[", "]", "{", "}", "(", ")", ";", ":", "::", "..", ",", "=", "*", "/", "+", "-", ".", "<", ">", "!", "graph", "input", "output",
            "node", "nodes", "connection", "connections", "value", "stream", "event", "name", "ramp", "0", "8", "1e39", "-0.0", "4294967296",
            "99999999999999999999", "x", "gate", "Delay::new(1.0, 0.0)", "[sinc]", "[x]", "\"s", "/*", "//", "\x00", "é"]
    n_ok = n_err = 0
    for it in range(400):
        src = rng.choice(corpus)
        toks = _re.findall(r"\s+|[A-Za-z_][A-Za-z_0-9]*|[0-9][0-9_.eE]*|->|::|\.\.|.", src)
        for _ in range(rng.randint(1, 6)):
            k = rng.randrange(len(toks))
            op = rng.random()
            if op < 0.3:
                del toks[k]
            elif op < 0.5:
                toks.insert(k, toks[k])
            elif op < 0.7:
                j = rng.randrange(len(toks)); toks[k], toks[j] = toks[j], toks[k]
            elif op < 0.95:
                toks[k] = rng.choice(junk)
            else:
                toks = toks[:k]
            if not toks:
                toks = [";"]
        mutated = "".join(toks).replace("\x00", "")
        r = _run(graphc, ["--cuda", "-"], mutated)
        assert r.returncode in (0, 1), (r.returncode, r.stderr[-2000:], mutated[:400])
        if r.returncode == 1:
            assert _re.match(r"^\d+:\d+: ", r.stderr), r.stderr[:300]
            n_err += 1
        else:
            n_ok += 1
    assert n_err > 100 and n_ok > 10, (n_ok, n_err)


# ------------------------------------------------------------------ nested graphs (a graph used as a scalar node)
NESTED = """
graph! {
    name: EventVoice;
    input event gate;
    input value freq = 440.0;
    input stream fm;
    output stream audio;
    nodes {
        osc = PolyBlepOscillator::sine(440.0, 0.5);
        envelope = AdsrEnvelope::new(0.01, 0.1, 0.7, 0.3);
    }
    connections {
        gate -> envelope.gate;
        freq -> osc.frequency;
        fm -> osc.frequency_mod;
        osc.output * envelope.output -> audio;
    }
}
graph! {
    name: DualVoiceSynth;
    input event gate;
    input value detune = 1.5;
    output stream out;
    nodes {
        lfo = PolyBlepOscillator::sine(3.0, 0.01);
        voice1 = EventVoice;
        voice2 = EventVoice::new();
    }
    connections {
        gate -> voice1.gate;
        gate -> voice2.gate;
        220.0 -> voice1.freq;
        lfo.output -> voice2.fm;
        lfo.output * detune -> voice2.fm;
        voice1.audio + voice2.audio * 0.5 -> out;
    }
}
"""

FLAT = """
    name: DualVoiceSynth;
    input event gate;
    input value detune = 1.5;
    output stream out;
    nodes {
        lfo = PolyBlepOscillator::sine(3.0, 0.01);
        v1osc = PolyBlepOscillator::sine(440.0, 0.5);
        v1env = AdsrEnvelope::new(0.01, 0.1, 0.7, 0.3);
        v2osc = PolyBlepOscillator::sine(440.0, 0.5);
        v2env = AdsrEnvelope::new(0.01, 0.1, 0.7, 0.3);
    }
    connections {
        v1osc.output * v1env.output + v2osc.output * v2env.output * 0.5 -> out;      // the inliner keeps the parent's statements first
        gate -> v1env.gate;
        220.0 -> v1osc.frequency;
        0.0 -> v1osc.frequency_mod;
        gate -> v2env.gate;
        440.0 -> v2osc.frequency;
        lfo.output + lfo.output * detune -> v2osc.frequency_mod;
    }
"""


def test_nested_graphs_are_inlined_like_a_hand_flattened_graph():
    """oscen-lib/tests/nested_graph_test.rs: a graph used as a node (both `v = G;` and `v = G::new();`), with event,
    value and stream inputs; unconnected inputs keep their declared defaults; two sources into a nested stream input sum"""
    nested, flat = abi.Graph(NESTED), abi.Graph(FLAT)
    norm = lambda s: s.replace("voice1.osc", "v1osc").replace("voice1.envelope", "v1env").replace("voice2.osc", "v2osc").replace("voice2.envelope", "v2env")
    assert norm(nested.describe()) == flat.describe()
    strip = lambda c: re.sub(r"/\*.*?\*/", "", c)                 # the generated kernels differ in comments (node names) only
    assert strip(nested.codegen()) == strip(flat.codegen())
    assert nested.jit_check() > 10000


@pytest.mark.parametrize("mutation,frag", [
    (("voice1.audio +", "voice1.sound +"), "graph `EventVoice` has no output `sound`"),
    (("220.0 -> voice1.freq;", "220.0 -> voice1.pitch;"), "graph `EventVoice` has no input `pitch`"),
    (("voice2 = EventVoice::new();", "voice2 = EventVoice::new() * 2;"), "a rate on a nested graph instance"),
    (("voice2 = EventVoice::new();", "voice2 = EventVoice::with(1.0);"), "a graph is instantiated with"),
    (("lfo.output -> voice2.fm;", "lfo.output -> voice2.freq;"), "source is Stream but destination expects Value"),
])
def test_nested_graph_errors(mutation, frag):
    assert mutation[0] in NESTED
    with pytest.raises(abi.GraphCompileError) as ei:
        abi.Graph(NESTED.replace(mutation[0], mutation[1]))
    assert frag in str(ei.value), str(ei.value)


def test_self_nesting_is_reported():
    with pytest.raises(abi.GraphCompileError) as ei:
        abi.Graph("graph! { name: A; output stream o; node b = B; connection b.o -> o; } graph! { name: B; output stream o; node a = A; connection a.o -> o; }")
    assert "recursion deeper than 16 levels" in str(ei.value)


def test_oversample_variants_macro_and_named_graph_selection():
    src = text("sat_variants.graph")
    g4 = abi.Graph(src)                                        # the last variant
    assert "graph SatGraph_4x" in g4.describe() and "[sinc down x4]" in g4.describe()
    g1 = abi.Graph(src, name="SatGraph_1x")
    assert g1.launches() == ["0: PolyBlepOscillator::saw [osc]", "1: HardClip [clip]"]
    assert g4.n_inputs == 0 and g4.jit_check() > 10000
    with pytest.raises(abi.GraphCompileError) as ei:
        abi.Graph(src, name="SatGraph_2x")
    assert "no graph named `SatGraph_2x`" in str(ei.value) and "SatGraph_1x, SatGraph_4x" in str(ei.value)
    for broken, frag in ((src.replace("factors: [1, 4];", ""), "needs base_name, factors and body"),
                         (src.replace("factors: [1, 4];", "factors: [1, 3];"), "rate factor must be in {1, 2, 4, 8}"),
                         (src.replace("base_name: SatGraph;", "base: SatGraph;"), "unknown key `base`")):
        with pytest.raises(abi.GraphCompileError) as ei:
            abi.Graph(broken)
        assert frag in str(ei.value), str(ei.value)


# ------------------------------------------------------------------ random graphs (tests/random_graphs.py)
def test_random_graphs_compile_and_plan_in_topological_order():
    """300 random acyclic graphs over the whole node library: every one compiles, and in the launch plan every node
    comes after the nodes it reads (the walk's only ordering requirement)"""
    import numpy as np
    import random_graphs as R
    rng = np.random.default_rng(2026)
    n_jit = 0
    for it in range(300):
        rg = R.make_graph(rng)
        src = rg.dsl()
        try:
            g = abi.Graph(src)
        except abi.GraphCompileError as ex:
            raise AssertionError(f"graph {it} does not compile: {ex}\n{src}")
        flat = [n for grp in groups(g.launches()) for n in grp]
        pos = {}
        for k, name in enumerate(flat):
            pos.setdefault(name, k)
        for i, nd in enumerate(rg.nodes):
            if f"n{i}" not in pos:
                continue                                     # dead node: dropped
            for ep, srcs in nd["conns"].items():
                for s in srcs:
                    stack = [s]
                    while stack:
                        e = stack.pop()
                        if e.kind == "node":
                            assert pos[f"n{e.args[0]}"] < pos[f"n{i}"], (it, i, e.args[0], src)
                        elif e.kind in "+-*/":
                            stack += [e.args[0], e.args[1]]
                        elif e.kind not in ("value", "lit", "input"):
                            stack.append(e.args[0])
        if it % 60 == 0:                                     # a few of the generated kernels through NVRTC
            assert g.jit_check() > 10000
            n_jit += 1
    assert n_jit == 5


def test_graphc_lists_the_node_library(graphc):
    r = _run(graphc, ["--nodes"])
    assert r.returncode == 0
    for frag in ("AdsrEnvelope::new(attack, decay, sustain, release)", "input  gate: event", "FmOperator::new()",
                 "input  envelope: stream (unconnected: 1.000000)", "Delay::new(delay_samples, feedback) | from_seconds",
                 "input  pulse_width: value (unconnected: 0.500000)", "HardClip::new()", "SamplePlayer::new()", "input  buf: asset"):
        assert frag in r.stdout, frag
    # every type the listing names is one the front end accepts, with the listed endpoints
    for type_name in re.findall(r"^(\w+)::", r.stdout, re.M):
        ctor = {"AdsrEnvelope": "new(0.1, 0.1, 0.5, 0.1)", "FmOperator": "new()", "TptFilter": "new(1000.0, 0.7)", "PolyBlepOscillator": "sine(1.0, 1.0)",
                "LP18Filter": "new(1000.0, 0.5)", "IirLowpass": "new(1000.0, 0.7)", "Oscillator": "sine(1.0, 1.0)", "Delay": "new(1.0, 0.0)",
                "Gain": "new(1.0)", "AddValue": "new(0.0)", "Mixer": "new()", "Crossfade": "new()", "HardClip": "new()", "SamplePlayer": "new()"}[type_name]
        out = "output_a" if type_name == "Crossfade" else "output"
        abi.Graph(f"output stream o; node n = {type_name}::{ctor}; connection n.{out} -> o;")


def test_random_feedback_and_multirate_graphs_compile_and_generate_valid_cuda():
    """the two harder generators of the differential fuzzer (tests/random_graphs.py): every graph compiles; cycles are
    accepted exactly because they run through vias; multirate graphs get one resampler per cross-rate edge; a sample of
    the generated kernels goes through NVRTC (no GPU needed)"""
    import numpy as np
    import random_graphs as R
    for it in range(120):
        rg = R.make_graph(np.random.default_rng(5000 + it), feedback=True)
        g = abi.Graph(rg.dsl())
        n_delays = sum(1 for nd in rg.nodes if nd["type"] == "delay") + len(rg.feedbacks)
        assert len(g.launches()) >= 1 and g.describe().count("-> [") <= len(rg.feedbacks) and g.describe().count("Delay [") <= n_delays
        if it % 40 == 0:
            assert g.jit_check() > 10000
    for it in range(120):
        rg = R.make_graph(np.random.default_rng(9000 + it), multirate=True)
        g = abi.Graph(rg.dsl())
        d = g.describe()
        live_cross = d.count("UpEdge [") + d.count("DownEdge [")
        assert live_cross <= len(rg.cross)                        # dead edges are dropped with their nodes, none is invented
        if any(r for r in rg.rates):
            with pytest.raises(abi.EngineError):                   # multirate graphs never run as node plans
                if "Edge [" in d:
                    abi.Engine.from_graph(g, 8, plan=abi.PLAN_NODES)
                else:
                    raise abi.EngineError("no live cross-rate edge")
        if it % 40 == 0:
            assert g.jit_check() > 10000


def test_c_abi_header_is_valid_c_and_cxx(tmp_path):
    """include/oscen_b200.h is the drop-in boundary: it must compile as plain C (for cgo / Rust bindgen / ctypes-style
    consumers) and as C++"""
    import subprocess
    root = os.path.dirname(GRAPHS[:-len("/graphs")])
    hdr = os.path.join(root, "include", "oscen_b200.h")
    for lang, std in (("c", "-std=c99"), ("c", "-std=c11"), ("c++", "-std=c++17")):
        subprocess.check_call(["gcc", "-x", lang, std, "-pedantic", "-Wall", "-Werror", "-fsyntax-only", hdr])


def test_generated_kernel_compiles_from_a_relocated_library(tmp_path):
    """the device headers a generated kernel includes are embedded in the shared library (build.py -> jit_headers.inc ->
    jit.hpp): a library copied away from its source tree still compiles a generated kernel, cross-rate edges included"""
    import shutil
    import subprocess
    import sys
    from oscen_b200 import abi
    lib = tmp_path / "liboscen_b200.so"
    shutil.copy(abi.LIB, lib)
    code = ("from oscen_b200 import abi\n"
            "import sys\n"
            "for p in sys.argv[1:]:\n"
            "    n = abi.Graph(open(p).read()).jit_check()\n"
            "    assert n > 10000, (p, n)\n"
            "    print(p, n)\n")
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, OSCEN_B200_LIB=str(lib), PYTHONPATH=os.path.dirname(here))
    env.pop("OSCEN_B200_CSRC", None)
    r = subprocess.run([sys.executable, "-c", code, os.path.join(here, "graphs", "voice_fm3.graph"), os.path.join(here, "graphs", "sat_4x.graph")],
                       env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "voice_fm3" in r.stdout and "sat_4x" in r.stdout


def test_param_specs_with_commas_and_the_three_setters():
    """oscen-lib/tests/param_spec_test.rs: `[log, ramp: 100]` parses (commas inside the spec), and the ramped input has the three
    generated setters — set_freq (the declared 100-frame ramp), set_freq_with_ramp (a custom length), set_freq_immediate — after
    which `freq.current == 1000.0`.  (The reference graph declares no output; a bank has nothing to render without one, so
    that text itself is refused with a diagnostic and the mirror adds `filter.output -> out`.)"""
    from host_kernel import HostEngine
    ref_text = """graph! { name: TestGraph; input value freq = 440.0 [log, ramp: 100];
        nodes { osc = PolyBlepOscillator::saw(440.0, 0.6); filter = TptFilter::new(1000.0, 0.707); }
        connections { freq -> osc.frequency_mod; osc.output -> filter.input; } }"""
    with pytest.raises(abi.GraphCompileError, match="declares no stream output"):
        abi.Graph(ref_text)
    g = abi.Graph(ref_text.replace("nodes {", "output stream out; nodes {").replace("osc.output -> filter.input;", "osc.output -> filter.input; filter.output -> out;"))
    info = g.input_info(g.input_id("freq"))
    assert info["ramp"] == 100 and info["log"] and info["range"] is None and info["default"] == 440.0
    eng = HostEngine(g, 3, sample_rate=44100.0)
    fid = g.input_id("freq")
    eng.set("freq", 880.0)                       # set_freq: the declared ramp
    eng.process_block(40)
    assert 440.0 < float(eng.ramps[fid].current) < 880.0 and eng.ramps[fid].frames_remaining == 60
    eng.set_value_ramp(fid, 440.0, 50)           # set_freq_with_ramp
    eng.process_block(50)
    assert float(eng.ramps[fid].current) == 440.0 and eng.ramps[fid].frames_remaining == 0
    eng.set_value(fid, 1000.0)                   # set_freq_immediate
    eng.process_block(8)
    assert float(eng.ramps[fid].current) == 1000.0 and float(eng.pv[fid, 0]) == 1000.0
    import numpy as np
    assert np.isfinite(eng.read_voice_out()).all()


def test_multirate_parse_smoke_graphs_of_the_reference_compile():
    """oscen-lib/tests/multirate_parse.rs: `* 4` on a node and a `[sinc]` connection policy compile; init at 44.1 and 48 kHz"""
    for body, factor in (("osc = PolyBlepOscillator::saw(440.0, 0.6) * 4; } connections { osc.output -> audio_out;", 4),
                         ("osc = PolyBlepOscillator::saw(440.0, 0.6) * 2; } connections { [sinc] osc.output -> audio_out;", 2)):
        g = abi.Graph("graph! { name: Smoke; output stream audio_out; nodes { " + body + " } }")
        assert g.launches()[1] == "1: DownEdge [osc.output -> audio_out [sinc down x%d]]" % factor
        for sr in (44100.0, 48000.0):
            _, words, info = g.kernel_info(sr)
            assert info.inner_factor == factor and len(words) == info.n_state
        assert g.jit_check() > 10000


def test_type_validation_cases_of_the_reference():
    """oscen-lib/tests/type_validation_test.rs: the valid graph's connections are accepted (value input -> a stream input,
    stream -> stream); an OUTPUT as a destination and an INPUT as a source are compile errors (the two commented-out graphs)"""
    head = "graph! { input value freq = 440.0; output stream out; nodes { osc = PolyBlepOscillator::saw(440.0, 0.6); filter = TptFilter::new(1000.0, 0.707); } connections { "
    abi.Graph(head + "freq -> osc.frequency_mod; osc.output -> filter.input; filter.output -> out; } }")
    with pytest.raises(abi.GraphCompileError, match="`osc.output` is an output: it cannot be a connection destination"):
        abi.Graph(head + "freq -> osc.output; filter.output -> out; } }")
    with pytest.raises(abi.GraphCompileError, match="`filter.input` is an input: it cannot be a connection source"):
        abi.Graph(head + "filter.input -> osc.frequency; osc.output -> out; } }")
