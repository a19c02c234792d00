"""CPU parity tests of the GENERATED kernels, executed on the host (tests/host_kernel.py): the CUDA text codegen.hpp
emits for a graph — the text an engine hands to NVRTC — is compiled by g++ against a stand-in CUDA runtime header and
the engine's own device headers, run as 64-thread CTAs of OS threads, and compared bit for bit with the oracle.  The
DSL parity tests proper live in tests/test_dsl_gpu.py, whose every test also has a host-executed variant in the CPU suite;
here are the cases that exist only on the host: the differential fuzzer at host sizes (generated kernels, and node plans
for the graphs that have one — tests/csrc/host_nodes_runner.cpp), the reference's TptFilter known
answer, the SamplePlayer graphs, and the HAND-WRITTEN fused FMVoice kernel (the one bench.py times).  A code-generator
or node-body regression shows up here before any GPU time is spent.  The GPU tests remain the parity tests proper."""
import os

import numpy as np
import pytest

import oracle_lib as O
from host_kernel import HostEngine, HostFmEngine, HostNodePlanEngine
from parity_util import lcg_noise, lcg_uniform, seq_sum_f32
from test_parity_gpu import _fm_bank_script

GATE = O.EV_GATE
GRAPHS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "graphs")


def text(name):
    with open(os.path.join(GRAPHS, name)) as f:
        return f.read()


@pytest.mark.parametrize("kind", ["plain", "feedback", "multirate"])
@pytest.mark.parametrize("seed", range(int(os.environ.get("OSCEN_HOST_FUZZ_SEEDS", "6"))))
def test_host_executed_random_graph_kernels_match_the_oracle_walk(kind, seed):
    """the differential fuzzer of tests/test_dsl_gpu.py (same generator, same seeds: tests/random_graphs.py) on the
    host-executed kernel: plain graphs over the whole node library, graphs with `-> [N] ->` feedback cycles, graphs with
    an oversampled section — against the per-sample walk over the oracle's node primitives, over two blocks"""
    import random_graphs as R
    from oscen_b200 import abi
    rng = np.random.default_rng({"plain": 1000, "feedback": 5000, "multirate": 9000}[kind] + seed)
    rg = R.make_graph(rng, feedback=(kind == "feedback"), multirate=(kind == "multirate"))
    src = rg.dsl()
    n, frames = 3, 80
    g = abi.Graph(src)
    vals = {name: (d * (1.0 + 0.1 * np.arange(n))).astype(np.float32) for name, d in rg.values.items()}
    x = np.stack([lcg_noise(frames, 300 + seed * 8 + v) for v in range(n)], axis=1) if rg.has_input else None
    gates = [(v, 2 + 3 * v, 0.5 + 0.1 * v) for v in range(n)] + [(v, 50 + v, 0.0) for v in range(n)] if rg.has_gate else []
    want = rg.walk(n, frames, vals, gates, x)
    if not np.isfinite(want).all():
        pytest.skip("this random graph diverges: nothing to compare")
    engines = [("generated kernel", HostEngine(g, n))]
    if kind == "plain":                                       # same-rate graphs without feedback also run as node plans
        engines.append(("node plan", HostNodePlanEngine(src, n)))
    for label, eng in engines:
        for name, arr in vals.items():
            eng.set_value(g.input_id(name), arr)
        got = []
        for lo, hi in ((0, 33), (33, 80)):
            if x is not None:
                eng.stream_in(0, x[lo:hi])
            ev = [(v, f - lo, GATE, s) for v, f, s in gates if lo <= f < hi]
            if ev:
                eng.push_events(ev)
            eng.process_block(hi - lo)
            got.append(eng.read_voice_out())
        got = np.concatenate(got)
        assert np.array_equal(got, want), (kind, seed, label, float(np.nanmax(np.abs(got - want))), src)
        eng.close()


def test_host_executed_tpt_filter_matches_the_reference_impulse_response():
    """the reference's one golden waveform (filters/tpt/mod.rs:152-161,227-264, committed in tests/golden/reference_kats.json)
    against the generated kernel's TptFilter body, at the reference's 44.1 kHz"""
    import json
    k = json.load(open(os.path.join(os.path.dirname(GRAPHS), "golden", "reference_kats.json")))["tpt_impulse_response"]
    src = f"input audio_in: stream; output stream out; node f = TptFilter::new({k['cutoff']!r}, {k['q']!r}); connection audio_in -> f.input; connection f.output -> out;"
    x = np.zeros((8, 4), dtype=np.float32); x[0, :] = 1.0
    eng = HostEngine(src, 4, 1, k["sample_rate"])
    eng.stream_in(0, x); eng.process_block(8)
    y = eng.read_voice_out()
    for v in range(4):
        assert np.abs(y[:, v] - np.array(k["output"], dtype=np.float32)).max() <= k["epsilon"], y[:, v]


# ------------------------------------------------------------ `external` asset slots and SamplePlayer
@pytest.mark.parametrize("engine", [HostEngine, HostNodePlanEngine], ids=["generated_kernel", "node_plan"])
def test_host_executed_sample_player_graph_reproduces_the_reference_graph_test(engine):
    """oscen-lib/tests/sample_player_graph.rs:64-124 on the generated kernel and on the node plan: silent before any load,
    buffer A looped (across a block boundary), then a load of B restarts the playhead and B loops — every voice alike"""
    a = np.array([0.1, -0.2, 0.3, -0.4, 0.5], dtype=np.float32)
    b = np.array([0.9, 0.8, 0.7], dtype=np.float32)
    eng = engine(text("player.graph"), 70, 1, 44100.0)
    eng.process_block(8)
    assert (eng.read_voice_out() == 0.0).all()
    eng.load_asset("clip", a)
    eng.process_block(7); y1 = eng.read_voice_out()
    eng.process_block(3); y2 = eng.read_voice_out()
    want = np.concatenate([a, a])
    assert np.array_equal(np.concatenate([y1, y2]), np.repeat(want[:, None], 70, axis=1))
    eng.load_asset("clip", np.stack([b, -b], axis=1))           # a stereo source: a mono player takes channel 0
    eng.process_block(6)
    assert np.array_equal(eng.read_voice_out(), np.repeat(np.concatenate([b, b])[:, None], 70, axis=1))
    eng.load_asset("clip", np.zeros(0, dtype=np.float32))       # an empty buffer: silence again
    eng.process_block(4)
    assert (eng.read_voice_out() == 0.0).all()


PLAYERS_DSL = """
    input level: value = 0.5;
    input cutoff: value = 900.0;
    output stream out;
    external loop_a: AudioAsset;
    external loop_b: AudioAsset;
    node pa = SamplePlayer::new();
    node pb = SamplePlayer::new();
    node idle = SamplePlayer::new();
    node filt = TptFilter::new(900.0, 0.8);
    connection loop_a -> pa.buf;
    connection loop_b -> pb.buf;
    connection pa.output * level + idle.output -> filt.input;
    connection cutoff -> filt.cutoff;
    connection filt.output + pb.output * 0.25 -> out;
"""


@pytest.mark.parametrize("engine", [HostEngine, HostNodePlanEngine], ids=["generated_kernel", "node_plan"])
def test_host_executed_two_sample_players_through_a_filter_match_the_oracle_walk(engine):
    """two slots loaded at different times (one reloaded with a SHORTER buffer mid-run), a player whose `buf` is not
    connected (silence), per-voice level and cutoff — against a walk over the oracle's SamplePlayer and TptFilter"""
    import graph_walk as W
    n, sr = 5, 48000.0
    f32 = np.float32
    rng = np.random.default_rng(7)
    bufs = {k: rng.uniform(-1, 1, size).astype(np.float32) for k, size in (("a1", 37), ("b1", 11), ("a2", 5))}
    level = (0.3 + 0.1 * np.arange(n)).astype(np.float32)
    cutoff = (500.0 + 300.0 * np.arange(n)).astype(np.float32)
    script = [(20, {}), (50, {"loop_a": "a1"}), (41, {"loop_b": "b1"}), (64, {"loop_a": "a2"})]      # (frames, loads before the block)

    eng = engine(PLAYERS_DSL, n, 1, sr)
    eng.set_value(eng.graph.input_id("level"), level)
    eng.set_value(eng.graph.input_id("cutoff"), cutoff)
    got = []
    for frames, loads in script:
        for slot, key in loads.items():
            eng.load_asset(slot, bufs[key])
        eng.process_block(frames); got.append(eng.read_voice_out())
    got = np.concatenate(got)

    want = np.zeros_like(got)
    for v in range(n):
        pa, pb, idle, filt = W.sample_player(), W.sample_player(), W.sample_player(), W.tpt(900.0, 0.8, sr)
        slots = {"loop_a": pa, "loop_b": pb}
        f = 0
        for frames, loads in script:
            for slot, key in loads.items():
                slots[slot].publish(bufs[key])
            for _ in range(frames):
                a, b, i = pa.process(), pb.process(), idle.process()
                filt.set(input=f32(f32(a * level[v]) + i), cutoff=cutoff[v])
                want[f, v] = f32(filt.process() + f32(b * f32(0.25)))
                f += 1
    assert np.abs(want).max() > 0.05
    assert np.array_equal(got, want), float(np.abs(got - want).max())


# ------------------------------------------------------------ the hand-written fused FMVoice kernel (the one bench.py times)
@pytest.mark.parametrize("force", [None, "legacy", "general"], ids=["product_choice", "legacy_kernel", "general_kernel"])
def test_host_executed_hand_written_fm_kernel_matches_oracle_bank_both_instantiations(force):
    """kernel_fm.cuh / kernel_fm_seg.cuh on the host, 136 voices = 17 instances of 8: (1) BASELINE config 2's steady state —
    per-voice frequency, everything else uniform, gates only: the segmented kernel (product choice), the round-1
    GENERAL=false kernel ("legacy"), or fm_voice_kernel<true>; (2) the events-and-ramps script of the GPU parity test: the
    segmented kernel's ramp variant, or GENERAL=true.  Per-voice output and instance fold bit-identical to the oracle; the
    kernel generated from the DSL text agrees"""
    from parity_util import lcg_ints
    n, vpi = 136, 8
    notes = 36 + lcg_ints(n, 1, 60)
    freq = (440.0 * np.exp2((notes - 69) / 12.0)).astype(np.float32)
    vel = (0.25 + 0.75 * lcg_uniform(n, 2)).astype(np.float32)

    def steady(b):
        b.set_value(0, freq)
        b.set_value(3, 0.3)                                       # op3_feedback, uniform
        out = []
        for blk in range(3):
            if blk == 0:
                b.push_events([(v, int(3 + v % 90), GATE, float(vel[v])) for v in range(n)])
            if blk == 2:
                b.push_events([(v, int(v % 64), GATE, 0.0) for v in range(0, n, 2)])
                b.push_events([(v, 5, 0, float(freq[v] * 0.5)) for v in range(1, n, 3)])     # a frequency event keeps GENERAL=false
            b.process_block(96)
            out.append((b.read_voice_out(), b.read_instance_out()))
        return np.concatenate([o[0] for o in out]), np.concatenate([o[1] for o in out])

    ref, hand, gen = O.OracleBank(O.K_FM, n, vpi), HostFmEngine(n, vpi), HostEngine(text("voice_fm3.graph"), n, vpi)
    hand.force_variant = force
    (vr, ir), (vh, ih), (vg, ig) = steady(ref), steady(hand), steady(gen)
    assert hand.last_general is False
    assert hand.variants == [{None: "seg", "legacy": "legacy", "general": "general"}[force]] * 3
    assert np.abs(vr).max() > 0.05
    assert np.array_equal(vh, vr) and np.array_equal(ih, ir), float(np.abs(vh - vr).max())
    assert np.array_equal(vg, vr) and np.array_equal(ig, ir)

    ref, hand = O.OracleBank(O.K_FM, n, vpi), HostFmEngine(n, vpi)
    hand.force_variant = force
    (vr, ir, mr), (vh, ih, mh) = [_fm_bank_script(b, n, blocks=3, frames=128) for b in (ref, hand)]
    # block 0: nothing moves; blocks 1-2: ramps on op3_level and filter_env_amount (block 2 also retunes voices: `frequency` events)
    assert hand.variants == {None: ["seg", "seg_ramp", "seg_ramp"], "legacy": ["legacy", "general", "general"], "general": ["general"] * 3}[force]
    assert np.array_equal(vh, vr) and np.array_equal(ih, ir), float(np.abs(vh - vr).max())
    scale = np.abs(ir.astype(np.float64)).sum(axis=1)
    assert (np.abs(mh.astype(np.float64) - mr.astype(np.float64)) <= 1e-6 + 2e-6 * scale).all()


def test_host_executed_generated_kernel_without_segments_matches_too(monkeypatch):
    """OSCEN_B200_JIT_SEGMENTS=0: the generator's per-frame loop of round 1 (kept as the A/B and fallback form) renders the same
    bits as the segmented default on a script with gates, a ramp and per-voice value events"""
    n, vpi = 24, 8
    outs = []
    for seg in ("1", "0"):
        monkeypatch.setenv("OSCEN_B200_JIT_SEGMENTS", seg)
        eng = HostEngine(text("voice_fm3.graph"), n, vpi)
        src = eng.graph.codegen()
        assert ("__reduce_min_sync(" in src) == (seg == "1")
        outs.append(_fm_bank_script(eng, n, blocks=3, frames=128))
    ref = _fm_bank_script(O.OracleBank(O.K_FM, n, vpi), n, blocks=3, frames=128)
    for got in outs:
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])


@pytest.mark.parametrize("seed", [1047, 1131])
def test_host_fuzzer_findings_stay_fixed(seed):
    """the two multirate graphs of tools/fuzz_host_kernels.py (seeds 777000 + 1047 / 1131) that exposed the inner-rate
    FmOperator / AdsrEnvelope re-derivation at the wrong sample rate (see tests/test_zz_regressions.py)"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GRAPHS), "..", "tools"))
    import fuzz_host_kernels as F
    assert F.run("multirate", seed) == "ok"


# ------------------------------------------------------------ every other hand-written fused kernel (built-in graph kinds)
@pytest.mark.parametrize("kind", [O.K_SYNTH, O.K_SYNTH_ADSR, O.K_SAT1, O.K_SAT4, O.K_TWIN], ids=["synth", "synth_adsr", "sat_1x", "sat_4x", "twin_peaks"])
def test_host_executed_hand_written_kernels_match_oracle_banks(kind):
    """kernel_others.cuh on the host (tests/csrc/host_builtin_runner.cpp restates launch_fused()), 70 voices over two CTAs:
    per-voice values, gates, a ramp across a block boundary, per-voice value events — bit-identical to the oracle banks"""
    from host_kernel import HostBuiltinEngine
    n = 70
    ref, host = O.OracleBank(kind, n, 1), HostBuiltinEngine(kind, n, 1)
    x = np.stack([lcg_noise(512, 77 + v) for v in range(n)], axis=1) if kind == O.K_TWIN else None

    def script(b):
        out = []
        if kind in (O.K_SYNTH, O.K_SYNTH_ADSR):
            b.set_value(2, (80.0 + 9.0 * np.arange(n)).astype(np.float32)); b.set_value(3, (400.0 + 50.0 * np.arange(n)).astype(np.float32))
        elif kind in (O.K_SAT1, O.K_SAT4):
            b.set_value(0, (2000.0 * (0.5 + lcg_uniform(n, 5))).astype(np.float32))
        for blk, frames in enumerate((200, 50, 90)):
            if kind == O.K_TWIN:
                b.stream_in(0, x[blk * 100:blk * 100 + frames])
            if kind == O.K_SYNTH_ADSR and blk < 2:
                b.push_events([(v, (v + 3 * blk) % 50, GATE, 0.9 if blk == 0 else 0.0) for v in range(n)])
            if blk == 1 and kind not in (O.K_SAT1, O.K_SAT4):
                b.set_value_ramp(0, 7.0 if kind != O.K_TWIN else 1200.0, 70)
                b.push_events([(5, 40, 1, 0.3 if kind != O.K_TWIN else 2500.0)])
            if blk == 1 and kind in (O.K_SAT1, O.K_SAT4):
                b.set_value_ramp(1, 0.9, 70)                                              # amplitude ramps ...
                b.push_events([(v, 7 + v % 30, 0, 900.0 + 10.0 * v) for v in range(n)])   # ... while every voice is retuned
            b.process_block(frames); out.append(b.read_voice_out())
        return np.concatenate(out)

    yr, yh = script(ref), script(host)
    assert np.abs(yr).max() > 0.01
    assert np.array_equal(yh, yr), float(np.abs(yh - yr).max())


def test_host_executed_electric_piano_kernel_matches_oracle_voices_and_stereo_instances():
    """BASELINE configs[2]'s graph (kernel_others.cuh epiano_kernel: 32-harmonic voices, 16 per instance -> sum -> Tremolo ->
    Frame<2>), 5 instances over two CTAs, driven by per-voice frequency + gate events (what the MIDI front emits), a value
    change and a ramp on the tremolo depth: per-voice output and the stereo instance output bit-identical to the oracle"""
    from host_kernel import HostBuiltinEngine
    n, vpi = 80, 16
    ref, host = O.OracleBank(O.K_EPIANO, n, vpi), HostBuiltinEngine(O.K_EPIANO, n, vpi)

    def script(b):
        out_v, out_i = [], []
        freqs = (110.0 * 2.0 ** ((np.arange(n) % 29) / 12.0)).astype(np.float32)
        ev = []
        for v in range(0, n, 3):
            ev.append((v, 2 + v % 40, 0, float(freqs[v])))
            ev.append((v, 2 + v % 40, GATE, 0.4 + 0.005 * v))
        b.push_events(ev)
        for _ in range(2):
            b.process_block(200); out_v.append(b.read_voice_out()); out_i.append(b.read_instance_out())
        b.push_events([(v, 5 + v % 7, GATE, 0.0) for v in range(0, n, 6)])
        b.set_value(3, 60.0)
        b.set_value_ramp(7, 0.8, 150)
        for _ in range(2):
            b.process_block(120); out_v.append(b.read_voice_out()); out_i.append(b.read_instance_out())
        return np.concatenate(out_v), np.concatenate(out_i)

    (vr, ir), (vh, ih) = script(ref), script(host)
    assert np.abs(ir).max() > 1e-3 and np.abs(ir[..., 0] - ir[..., 1]).max() > 1e-6
    assert np.array_equal(vh, vr) and np.array_equal(ih, ir)


@pytest.mark.parametrize("array_smem", ["1", "0", "2"], ids=["state_in_shared_memory", "state_walked_in_hbm_rows", "hot_words_in_shared_memory_cold_rows_in_hbm"])
def test_host_executed_electric_piano_graph_from_its_own_dsl_text_matches_oracle(array_smem, monkeypatch):
    """ElectricPianoGraph compiled from the reference's two `graph!` blocks (tests/graphs/electric_piano.graph: an array-typed
    [f32; 32] stream between AmplitudeSource and OscillatorBank, 16 voices -> sum -> Tremolo -> Frame<2>): the GENERATED
    kernel, executed on the host, against the oracle's electric-piano bank — per-voice output and stereo instance output
    bit-identical, same script as the hand-written kernel's test above.  Both homes of the array nodes' 259 state words: the
    block-resident [word][thread] slices of dynamic shared memory (default) and the SoA rows in HBM
    (OSCEN_B200_JIT_ARRAY_SMEM=0, also the fallback of graphs whose slices do not fit)"""
    monkeypatch.setenv("OSCEN_B200_JIT_ARRAY_SMEM", array_smem)
    n, vpi = 80, 16
    ref, host = O.OracleBank(O.K_EPIANO, n, vpi), HostEngine(text("electric_piano.graph"), n)
    assert host.vpi == vpi and host.channels == 2
    src, _, info = host.graph.kernel_info(48000.0)
    assert ("dyn + 130u * kCta + threadIdx.x" in src) == (array_smem == "1") and ("dyn + 66u * kCta + threadIdx.x" in src) == (array_smem == "2")
    assert info.dyn_smem_bytes == {"1": 259 * 64 * 4, "2": (66 + 65) * 64 * 4, "0": 0}[array_smem]

    def script(b):
        out_v, out_i = [], []
        freqs = (110.0 * 2.0 ** ((np.arange(n) % 29) / 12.0)).astype(np.float32)
        ev = []
        for v in range(0, n, 3):
            ev.append((v, 2 + v % 40, 0, float(freqs[v])))
            ev.append((v, 2 + v % 40, GATE, 0.4 + 0.005 * v))
        b.push_events(ev)
        for _ in range(2):
            b.process_block(200); out_v.append(b.read_voice_out()); out_i.append(b.read_instance_out())
        b.push_events([(v, 5 + v % 7, GATE, 0.0) for v in range(0, n, 6)])
        b.set_value(3, 60.0)
        b.set_value_ramp(7, 0.8, 150)
        for _ in range(2):
            b.process_block(120); out_v.append(b.read_voice_out()); out_i.append(b.read_instance_out())
        return np.concatenate(out_v), np.concatenate(out_i)

    (vr, ir), (vh, ih) = script(ref), script(host)
    assert np.abs(ir).max() > 1e-3 and np.abs(ir[..., 0] - ir[..., 1]).max() > 1e-6
    assert np.array_equal(vh, vr) and np.array_equal(ih, ir)


@pytest.mark.parametrize("kind", range(7, 20))
def test_host_executed_clip_kernels_match_oracle(kind):
    """kernel_resample.cuh: ClipRef and the 12 ClipOversampled kinds (4 policies x {2, 4, 8}), state handed across blocks"""
    from host_kernel import HostBuiltinEngine
    n = 6
    ref, host = O.OracleBank(kind, n, 1), HostBuiltinEngine(kind, n, 1)
    x = np.stack([lcg_noise(300, 900 + v) for v in range(n)], axis=1) * np.float32(1.4)
    pos, yr, yh = 0, [], []
    for frames in (128, 22, 150):
        for b, outs in ((ref, yr), (host, yh)):
            b.stream_in(0, x[pos:pos + frames]); b.process_block(frames); outs.append(b.read_voice_out())
        pos += frames
    assert np.array_equal(np.concatenate(yh), np.concatenate(yr))


@pytest.mark.parametrize("engine", [HostEngine, HostNodePlanEngine], ids=["generated_kernel", "node_plan"])
def test_gate_off_on_the_frame_the_release_time_grows_uses_the_cached_length(engine):
    """found by tools/fuzz_host_builtin.py.  AdsrEnvelope::handle_gate_event (adsr.rs:252-273): a gate-OFF calls
    set_stage(Release), which counts down from `release_samples` as cached by the last process() — so when `release` was
    raised since then (a host write between blocks with the note-off at frame 0, or a ramp / value event on the note-off's
    frame) the reference releases over min(old, new) samples.  The kernels used to start from the new length; the cached
    one now travels in the upper bits of the ADSR stage word (nodes.cuh AdsrState::r_gate, state_io.cuh)."""
    import graph_walk as W
    src = ("input r: value = 0.004; input gate: event; output stream out; node env = AdsrEnvelope::new(0.002, 0.003, 0.6, 0.004); "
           "connection gate -> env.gate; connection r -> env.release; connection env.output -> out;")
    eng = engine(src, 1, 1)
    env = W.adsr(0.002, 0.003, 0.6, 0.004, 48000.0)
    got, want = [], []
    for frames, rel, gates in ((200, 0.004, [(0, 0.9)]), (100, 0.008, [(0, 0.0)])):
        eng.set_value(0, rel)
        eng.push_events([(0, f, GATE, g) for f, g in gates])
        eng.process_block(frames); got.append(eng.read_voice_out()[:, 0])
        for f in range(frames):
            env.set(release=rel)
            for gf, g in gates:
                if gf == f:
                    env.on_gate(g)
            want.append(env.process())
    assert np.array_equal(np.concatenate(got), np.array(want, dtype=np.float32))


@pytest.mark.parametrize("case", ["release grows while releasing", "release shrinks while releasing", "gate-off at frame 0 after the release shrank",
                                  "gate-off at frame 5 after the release grew", "gate-on at frame 0 after the attack changed", "decay grows while decaying",
                                  "attack shrinks while attacking", "sustain moves while sustaining"])
@pytest.mark.parametrize("engine", [HostEngine, HostNodePlanEngine], ids=["generated_kernel", "node_plan"])
def test_adsr_parameter_changes_mid_stage_match_the_reference(case, engine):
    """the siblings of the case above: every other way a time or level can change under a running envelope
    (update_sustain_level's countdown clamp, adsr.rs:84-115) is bit-identical"""
    import graph_walk as W
    src = ("input a: value = 0.002; input d: value = 0.003; input s: value = 0.6; input r: value = 0.004; input gate: event; output stream out; "
           "node env = AdsrEnvelope::new(0.002, 0.003, 0.6, 0.004); connection gate -> env.gate; connection a -> env.attack; "
           "connection d -> env.decay; connection s -> env.sustain; connection r -> env.release; connection env.output -> out;")
    blocks = {
        "release grows while releasing": [(200, {}, [(0, 0.9)]), (50, {}, [(10, 0.0)]), (100, {"r": 0.008}, [])],
        "release shrinks while releasing": [(200, {}, [(0, 0.9)]), (50, {}, [(10, 0.0)]), (100, {"r": 0.001}, [])],
        "gate-off at frame 0 after the release shrank": [(200, {}, [(0, 0.9)]), (100, {"r": 0.001}, [(0, 0.0)])],
        "gate-off at frame 5 after the release grew": [(200, {}, [(0, 0.9)]), (100, {"r": 0.008}, [(5, 0.0)])],
        "gate-on at frame 0 after the attack changed": [(50, {}, []), (100, {"a": 0.004}, [(0, 0.9)])],
        "decay grows while decaying": [(120, {}, [(0, 0.9)]), (100, {"d": 0.006}, [])],
        "attack shrinks while attacking": [(40, {}, [(0, 0.9)]), (100, {"a": 0.0005}, [])],
        "sustain moves while sustaining": [(150, {}, [(0, 0.9)]), (100, {"s": 0.3}, []), (100, {"s": 0.8}, [])],
    }[case]
    field = {"a": "attack", "d": "decay", "s": "sustain", "r": "release"}
    eng = engine(src, 1, 1)
    env = W.adsr(0.002, 0.003, 0.6, 0.004, 48000.0)
    got, want = [], []
    for frames, sets, gates in blocks:
        for k, v in sets.items():
            eng.set_value(eng.graph.input_id(k), v)
        if gates:
            eng.push_events([(0, f, GATE, g) for f, g in gates])
        eng.process_block(frames); got.append(eng.read_voice_out()[:, 0])
        for f in range(frames):
            env.set(**{field[k]: v for k, v in sets.items()})
            for gf, g in gates:
                if gf == f:
                    env.on_gate(g)
            want.append(env.process())
    assert np.array_equal(np.concatenate(got), np.array(want, dtype=np.float32))


# ------------------------------------------------------------ the device-side MIDI front (midi_device.cuh)
@pytest.mark.parametrize("kind,vpi", [(O.K_FM, 8), (O.K_EPIANO, 16)], ids=["fm_graph_8_voices", "electric_piano_16_voices"])
@pytest.mark.parametrize("seed", range(4))
def test_host_executed_device_midi_front_feeds_the_voice_kernels_like_the_oracle(kind, vpi, seed):
    """raw MIDI in, audio out, entirely on host-executed device code: midi_front_kernel (parser -> VoiceAllocator<N> ->
    [MidiVoiceHandler; N], one thread per instance; tests/csrc/host_midi_runner.cpp restates midi_device_run()) writes the
    per-voice gate / frequency events, the hand-written voice kernel renders them — against the oracle bank fed the same
    messages through its own MIDI chain.  Random streams with voice stealing, velocity-0 note-ons, controller noise, short
    messages, more than 32 messages per instance per block and offsets beyond the block (_random_midi)."""
    from host_kernel import HostBuiltinEngine, HostMidiFront
    from test_parity_gpu import _random_midi
    rng = np.random.default_rng(100 * kind + seed)
    n_inst = 9 if kind == O.K_FM else 5                     # two CTAs' worth of voices, the second partly filled
    n = n_inst * vpi
    ref, host, front = O.OracleBank(kind, n, vpi), HostBuiltinEngine(kind, n, vpi), HostMidiFront(n, vpi, 0)
    vr, vh, ir, ih = [], [], [], []
    for blk, frames in enumerate((128, 60, 200)):
        msgs = _random_midi(rng, n_inst, frames, 40 if blk == 0 else 8, 2 * vpi)
        for inst, frame, data in msgs:
            ref.push_midi(inst, frame, data)
        host.push_events(front.run(msgs, frames))
        ref.process_block(frames); host.process_block(frames)
        vr.append(ref.read_voice_out()); vh.append(host.read_voice_out())
        ir.append(ref.read_instance_out()); ih.append(host.read_instance_out())
    assert np.abs(np.concatenate(vr)).max() > 0.01
    assert np.array_equal(np.concatenate(vh), np.concatenate(vr))
    assert np.array_equal(np.concatenate(ih), np.concatenate(ir))
    front.close()


@pytest.mark.parametrize("seed", range(3))
def test_host_executed_device_midi_front_takes_data_bytes_with_the_high_bit_set(seed):
    """note / velocity are any u8 in the reference (midi.rs:8-35): the device front must index its frequency table with the
    full byte and keep note 255 distinguishable from "no note" in the allocator word"""
    from host_kernel import HostBuiltinEngine, HostMidiFront
    from test_parity_gpu import _random_midi
    rng = np.random.default_rng(4200 + seed)
    kind, vpi, n_inst = O.K_FM, 8, 5
    n = n_inst * vpi
    ref, host, front = O.OracleBank(kind, n, vpi), HostBuiltinEngine(kind, n, vpi), HostMidiFront(n, vpi, 0)
    vr, vh = [], []
    for blk, frames in enumerate((96, 64)):
        msgs = _random_midi(rng, n_inst, frames, 30 if blk == 0 else 8, 2 * vpi, malformed=0.5)
        assert any(d[1] >= 0x80 for _, _, d in msgs if len(d) > 1)
        for inst, frame, data in msgs:
            ref.push_midi(inst, frame, data)
        host.push_events(front.run(msgs, frames))
        ref.process_block(frames); host.process_block(frames)
        vr.append(ref.read_voice_out()); vh.append(host.read_voice_out())
    assert np.abs(np.concatenate(vr)).max() > 0.01
    assert np.array_equal(np.concatenate(vh), np.concatenate(vr))
    front.close()


def test_host_executed_mix_reduce_kernel_folds_the_cta_partials_in_its_fixed_shape():
    """mix_reduce_kernel (kernels_common.cuh) on the host against its restatement in tests/host_kernel.py (which the mix-bus
    assertions of the other host tests use) and, within the reduction-order tolerance, against the oracle's sequential sum:
    mono (FM bank over three CTAs) and stereo (ElectricPiano)"""
    from host_kernel import HostBuiltinEngine
    for kind, n, vpi in ((O.K_FM, 136, 8), (O.K_EPIANO, 80, 16)):
        ref, host = O.OracleBank(kind, n, vpi), HostBuiltinEngine(kind, n, vpi)
        for b in (ref, host):
            b.set_value(0, (110.0 * (1 + np.arange(n) % 24)).astype(np.float32))
            b.push_events([(v, v % 20, GATE, 0.7) for v in range(n)])
            b.process_block(48)
        m_kernel, m_py, m_ref = host.read_mix_device_kernel(), host.read_mix(), ref.read_mix()
        assert np.array_equal(m_kernel, m_py)
        scale = np.abs(ref.read_instance_out().astype(np.float64)).sum(axis=1)
        assert np.abs(m_ref).max() > 1e-3
        assert (np.abs(m_kernel.astype(np.float64) - m_ref.astype(np.float64)) <= 1e-6 + 2e-6 * scale).all()


def test_host_executed_twin_peaks_kernel_with_the_device_generated_input():
    """BASELINE configs[3]'s input path: audio_in generated on the device by the per-voice LCG (stream_in_lcg; the two LCG
    state words persist across blocks) — the hand-written twin_peaks_kernel against the oracle fed the same noise"""
    from host_kernel import HostBuiltinEngine
    n = 70
    ref, host = O.OracleBank(O.K_TWIN, n, 1), HostBuiltinEngine(O.K_TWIN, n, 1)
    x = np.stack([lcg_noise(300, 1000 + 7 + v) for v in range(n)], axis=1)
    host.stream_in_lcg(0, 1000, 7)
    pos, yr, yh = 0, [], []
    for frames in (128, 22, 150):
        ref.stream_in(0, x[pos:pos + frames]); ref.process_block(frames); yr.append(ref.read_voice_out())
        host.process_block(frames); yh.append(host.read_voice_out())
        pos += frames
    assert np.abs(np.concatenate(yr)).max() > 0.1
    assert np.array_equal(np.concatenate(yh), np.concatenate(yr))
