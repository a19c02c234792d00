"""GPU parity tests: the CUDA engine (through the C ABI) against the CPU oracle on the same
seeded inputs.  Tolerance: |gpu - ref| <= 1e-6 + 1e-5*|ref| (BASELINE.json north_star)."""
import numpy as np
import pytest

import oracle_lib as O
from parity_util import assert_close, lcg_ints, lcg_noise, lcg_uniform, seq_sum_f32

pytestmark = pytest.mark.gpu

GATE = O.EV_GATE


def engine(kind, n, vpi=1, plan=0, keep=3, sr=48000.0):
    from oscen_b200 import abi
    return abi.Engine(kind, n, vpi, sr, 0, plan, keep)


def both(kind, n, vpi=1, plan=0, sr=48000.0):
    return O.OracleBank(kind, n, vpi, sr), engine(kind, n, vpi, plan, 3, sr)


def drive(banks, fn):
    outs = []
    for b in banks:
        outs.append(fn(b))
    return outs


def compare_banks(ref, gpu, what, check_inst=True):
    stats = assert_close(gpu.read_voice_out(), ref.read_voice_out(), what + " voice_out")
    if check_inst:
        assert_close(gpu.read_instance_out(), ref.read_instance_out(), what + " instance_out")
    mref = ref.read_mix().astype(np.float64)
    mgpu = gpu.read_mix().astype(np.float64)
    # the mix bus is summed in a different (tree) order: tolerance scaled by sum |instance|
    scale = np.abs(ref.read_instance_out().astype(np.float64)).sum(axis=1)
    assert (np.abs(mgpu - mref) <= 1e-6 + 2e-6 * scale).all(), what + " mix"
    return stats


# --------------------------------------------------------------- config 1: the anchor
@pytest.mark.parametrize("chunks", [[512, 512], [256, 256, 256, 256], [100, 412, 512]])
@pytest.mark.parametrize("variant", ["defaults", "feedback"])
def test_fm_voice_anchor_1_voice_1024_frames(chunks, variant):
    ref, gpu = both(O.K_FM, 1, 1)
    for b in (ref, gpu):
        b.set_value(0, 440.0)
        if variant == "feedback":
            b.set_value(3, 0.3); b.set_value(10, 0.2); b.set_value(20, 0.5); b.set_value(27, 1500.0)
    pos = 0
    outs = {id(ref): [], id(gpu): []}
    for frames in chunks:
        for b in (ref, gpu):
            ev = []
            if pos <= 0 < pos + frames:
                ev.append((0, 0 - pos, GATE, 1.0))
            if pos <= 768 < pos + frames:
                ev.append((0, 768 - pos, GATE, 0.0))
            b.push_events(ev)
            b.process_block(frames)
            outs[id(b)].append(b.read_voice_out())
        pos += frames
    yr, yg = np.concatenate(outs[id(ref)]), np.concatenate(outs[id(gpu)])
    assert np.abs(yr).max() > 0.05
    max_err, exact = assert_close(yg, yr, f"FM anchor {variant} {chunks}")
    print(f"anchor {variant} {chunks}: max_err={max_err:.3e} bit_exact={exact:.4f}")


def _fm_bank_script(b, n, blocks=4, frames=256):
    notes = 36 + lcg_ints(n, 1, 60)
    freq = (440.0 * np.exp2((notes - 69) / 12.0)).astype(np.float32)
    vel = (0.25 + 0.75 * lcg_uniform(n, 2)).astype(np.float32)
    on = lcg_ints(n, 3, frames)
    off = lcg_ints(n, 4, frames)
    b.set_value(0, freq)
    outs_v, outs_i, outs_m = [], [], []
    for bi in range(blocks):
        if bi == 0:
            b.push_events([(v, int(on[v]), GATE, float(vel[v])) for v in range(n)])
        if bi == 1:
            b.set_value_ramp(2, 0.8, 300)       # op3_level ramp spanning two blocks
            b.set_value_ramp(27, 900.0, 100)    # filter_env_amount: tanf every sample while it moves
        if bi == 2:
            b.push_events([(v, int(off[v]), GATE, 0.0) for v in range(0, n, 2)])
            b.push_events([(v, 7, 0, float(freq[v] * 1.5)) for v in range(1, n, 4)])   # frame-accurate retune
        b.process_block(frames)
        outs_v.append(b.read_voice_out()); outs_i.append(b.read_instance_out()); outs_m.append(b.read_mix())
    return np.concatenate(outs_v), np.concatenate(outs_i), np.concatenate(outs_m)


@pytest.mark.parametrize("plan", [0, 1])
def test_fm_bank_1024_voices_events_ramps(plan):
    n, vpi = 1024, 8
    ref, gpu = both(O.K_FM, n, vpi, plan)
    (vr, ir, mr), (vg, ig, mg) = drive((ref, gpu), lambda b: _fm_bank_script(b, n))
    assert np.abs(vr).max() > 0.05
    max_err, exact = assert_close(vg, vr, f"FM bank plan={plan} voice_out")
    assert_close(ig, ir, f"FM bank plan={plan} instance_out")
    scale = np.abs(ir.astype(np.float64)).sum(axis=1)
    assert (np.abs(mg.astype(np.float64) - mr.astype(np.float64)) <= 1e-6 + 2e-6 * scale).all()
    # instance_out is the in-order f32 fold of its voices (codegen/emit_node.rs:463-466): exact
    fold = seq_sum_f32(vg.reshape(vg.shape[0], n // vpi, vpi), axis=2)
    assert np.array_equal(fold, ig[:, :, 0])
    print(f"FM bank plan={plan}: max_err={max_err:.3e} bit_exact={exact:.4f}")


def test_fm_node_plan_equals_fused_plan_bitwise():
    n, vpi = 512, 8
    a, b = engine(O.K_FM, n, vpi, plan=0), engine(O.K_FM, n, vpi, plan=1)
    (va, ia, ma), (vb, ib, mb) = drive((a, b), lambda e: _fm_bank_script(e, n, blocks=3))
    assert np.array_equal(va, vb) and np.array_equal(ia, ib)
    assert np.allclose(ma, mb, rtol=1e-5, atol=1e-4)         # mix bus: different (fixed) tree shapes per plan


def test_fm_per_voice_params_general_kernel():
    n, vpi = 256, 8
    ref, gpu = both(O.K_FM, n, vpi)
    rng = np.random.default_rng(5)
    ratios = rng.choice([0.5, 1.0, 2.0, 3.0, 7.0], size=n).astype(np.float32)
    attacks = rng.uniform(0.0, 0.05, size=n).astype(np.float32)
    attacks[::7] = 0.0                                       # attack <= 1e-5 -> straight to Decay
    rel = rng.uniform(0.0, 0.02, size=n).astype(np.float32)
    rel[::5] = 0.0                                           # release <= 1e-5 -> straight to Idle
    fb = rng.uniform(0.0, 0.4, size=n).astype(np.float32)

    def script(b):
        b.set_value(1, ratios); b.set_value(4, attacks); b.set_value(16, attacks)
        b.set_value(19, rel); b.set_value(7, rel); b.set_value(3, fb)
        b.set_value(0, (110.0 * (1 + np.arange(n) % 12)).astype(np.float32))
        out = []
        b.push_events([(v, v % 64, GATE, 0.9) for v in range(n)])
        b.process_block(384); out.append(b.read_voice_out())
        b.push_events([(v, (3 * v) % 200, GATE, 0.0) for v in range(n)])
        b.process_block(512); out.append(b.read_voice_out())
        b.push_events([(v, 0, GATE, 0.4) for v in range(0, n, 3)])       # retrigger at frame 0
        b.process_block(128); out.append(b.read_voice_out())
        return np.concatenate(out)

    yr, yg = drive((ref, gpu), script)
    assert_close(yg, yr, "FM per-voice params")


def test_fm_graph_midi_front_matches():
    n, vpi = 64, 8
    ref, gpu = both(O.K_FM, n, vpi)

    def script(b):
        out = []
        for inst in range(n // vpi):
            b.push_midi(inst, 3 + inst, [0x90, 48 + inst, 100])
            b.push_midi(inst, 40, [0x90, 55 + inst, 64])
            b.push_midi(inst, 40, [0x90, 60, 1])
        b.process_block(256); out.append(b.read_instance_out())
        for inst in range(n // vpi):
            b.push_midi(inst, 10, [0x80, 48 + inst, 0])
            b.push_midi(inst, 200, [0x90, 60, 0])             # note-on vel 0 == note-off
            for k in range(9):                                 # steal voices: 9 more notes on an 8-voice instance
                b.push_midi(inst, 100 + k, [0x90, 70 + k, 90])
        b.process_block(256); out.append(b.read_instance_out())
        b.process_block(256); out.append(b.read_instance_out())
        return np.concatenate(out)

    yr, yg = drive((ref, gpu), script)
    assert np.abs(yr).max() > 0.01
    assert_close(yg, yr, "FM MIDI front")


def test_events_beyond_block_are_dropped_and_queue_cap():
    ref, gpu = both(O.K_FM, 8, 8)

    def script(b):
        b.push_events([(0, 300, GATE, 1.0)])                   # beyond a 256-frame block: dropped
        b.push_events([(1, 5, GATE, 1.0)])
        b.process_block(256)
        return b.read_voice_out()

    yr, yg = drive((ref, gpu), script)
    assert np.abs(yr[:, 0]).max() == 0.0 and np.abs(yg[:, 0]).max() == 0.0
    assert_close(yg, yr, "dropped events")


# --------------------------------------------------------------- the other reference graphs
def test_twin_peaks_host_stream():
    n = 300
    ref, gpu = both(O.K_TWIN, n, 1)
    x = np.stack([lcg_noise(512, 1000 + v) for v in range(n)], axis=1)

    def script(b):
        out = []
        b.stream_in(0, x[:256]); b.process_block(256); out.append(b.read_voice_out())
        b.set_value(0, (800.0 + 10.0 * np.arange(n)).astype(np.float32))
        b.set_value_ramp(2, 0.8, 64)
        b.stream_in(0, x[256:]); b.process_block(256); out.append(b.read_voice_out())
        return np.concatenate(out)

    yr, yg = drive((ref, gpu), script)
    assert np.abs(yr).max() > 0.01
    max_err, exact = assert_close(yg, yr, "TwinPeaks")
    print(f"twin peaks: max_err={max_err:.3e} bit_exact={exact:.4f}")


def test_twin_peaks_device_lcg_equals_host_noise():
    n = 130
    ref, gpu = both(O.K_TWIN, n, 1)
    x = np.stack([lcg_noise(256, 1000 + 7 + v) for v in range(n)], axis=1)
    ref.stream_in(0, x); ref.process_block(256)
    gpu.stream_in_lcg(0, 1000, 7); gpu.process_block(256)
    assert_close(gpu.read_voice_out(), ref.read_voice_out(), "TwinPeaks LCG")


@pytest.mark.parametrize("kind", [O.K_SAT1, O.K_SAT4])
def test_saturator(kind):
    n = 200
    ref, gpu = both(kind, n, 1)
    f = (2000.0 * (0.5 + lcg_uniform(n, 5))).astype(np.float32)

    def script(b):
        b.set_value(0, f)
        out = []
        for frames in (256, 100, 156):
            b.process_block(frames); out.append(b.read_voice_out())
        return np.concatenate(out)

    yr, yg = drive((ref, gpu), script)
    assert np.abs(yr).max() > 0.3
    max_err, exact = assert_close(yg, yr, f"Sat kind={kind}")
    assert gpu.latency_samples() == (8 if kind == O.K_SAT4 else 0)
    print(f"sat {kind}: max_err={max_err:.3e} bit_exact={exact:.4f}")


@pytest.mark.parametrize("kind", [O.K_SYNTH, O.K_SYNTH_ADSR])
def test_readme_synth(kind):
    n = 96
    ref, gpu = both(kind, n, 1)

    def script(b):
        b.set_value(2, (55.0 * (1 + np.arange(n))).astype(np.float32))      # up to 5.3 kHz carriers
        b.set_value(3, (300.0 + 200.0 * np.arange(n)).astype(np.float32))  # cutoffs 300 .. 19 kHz+
        out = []
        if kind == O.K_SYNTH_ADSR:
            b.push_events([(v, v, GATE, 0.8) for v in range(n)])
        b.process_block(512); out.append(b.read_voice_out())
        b.set_value_ramp(0, 9.0, 200)
        if kind == O.K_SYNTH_ADSR:
            b.push_events([(v, 2 * v, GATE, 0.0) for v in range(n)])
        b.process_block(512); out.append(b.read_voice_out())
        return np.concatenate(out)

    yr, yg = drive((ref, gpu), script)
    assert np.abs(yr).max() > 0.05
    assert_close(yg, yr, f"Synth kind={kind}")


def test_electric_piano_16_voice_instances():
    n, vpi = 64, 16
    ref, gpu = both(O.K_EPIANO, n, vpi)

    def script(b):
        out_v, out_i = [], []
        for inst in range(n // vpi):
            for k in range(5):
                b.push_midi(inst, 2 * k + inst, [0x90, 40 + 7 * k + inst, 60 + 10 * k])
        for _ in range(4):
            b.process_block(512); out_v.append(b.read_voice_out()); out_i.append(b.read_instance_out())
        for inst in range(n // vpi):
            b.push_midi(inst, 5, [0x80, 40 + inst, 0])
        b.set_value(3, 60.0)
        for _ in range(2):
            b.process_block(512); out_v.append(b.read_voice_out()); out_i.append(b.read_instance_out())
        return np.concatenate(out_v), np.concatenate(out_i)

    (vr, ir), (vg, ig) = drive((ref, gpu), script)
    assert np.abs(ir).max() > 1e-3 and np.abs(ir[..., 0] - ir[..., 1]).max() > 1e-6
    assert_close(vg, vr, "EPiano voice_out")
    max_err, exact = assert_close(ig, ir, "EPiano instance_out (stereo)")
    print(f"epiano: max_err={max_err:.3e} bit_exact={exact:.4f}")


# --------------------------------------------------------------- invariants at full size
def test_block_split_invariance_on_gpu():
    n, vpi = 256, 8
    a, b = engine(O.K_FM, n, vpi), engine(O.K_FM, n, vpi)
    ev_on = [(v, (5 * v) % 200, GATE, 0.7) for v in range(n)]
    a.push_events(ev_on); a.process_block(256); ya = a.read_voice_out()
    y = []
    b.push_events([e for e in ev_on if e[1] < 100]); b.process_block(100); y.append(b.read_voice_out())
    b.push_events([(v, f - 100, i, s) for (v, f, i, s) in ev_on if f >= 100]); b.process_block(156); y.append(b.read_voice_out())
    assert np.array_equal(ya, np.concatenate(y))


# --------------------------------------------------------------- edge cases
@pytest.mark.parametrize("n,vpi", [(8, 8), (72, 8), (130, 1), (96, 2), (48, 16), (4, 4)])
@pytest.mark.parametrize("frames", [1, 7, 512])
def test_ragged_sizes_and_block_lengths(n, vpi, frames):
    """voice counts that do not fill a CTA, every supported instance size, 1-frame and max blocks"""
    ref, gpu = both(O.K_FM, n, vpi)

    def script(b):
        b.set_value(0, (100.0 + 37.0 * np.arange(n)).astype(np.float32))
        out = []
        b.push_events([(v, min(frames - 1, v % 5), GATE, 0.6) for v in range(n)])
        for _ in range(3):
            b.process_block(frames)
            out.append((b.read_voice_out(), b.read_instance_out(), b.read_mix()))
        return out

    ro, go = drive((ref, gpu), script)
    for (vr, ir, mr), (vg, ig, mg) in zip(ro, go):
        assert np.array_equal(vg, vr) and np.array_equal(ig, ir)
        assert np.allclose(mg, mr, rtol=1e-5, atol=1e-5)


def test_gate_queue_overflow_drops_like_try_push():
    """ArrayVec<_, 32>::try_push: the 33rd gate event of a voice in one block is dropped"""
    ref, gpu = both(O.K_FM, 8, 8)

    def script(b):
        ev = [(0, k, GATE, 0.9 if k % 2 == 0 else 0.0) for k in range(40)]      # 40 alternating gates on voice 0
        b.push_events(ev)
        b.process_block(64)
        return b.read_voice_out()

    yr, yg = drive((ref, gpu), script)
    assert np.array_equal(yg, yr)
    assert gpu.stats()["events_dropped"] >= 8


def test_value_input_api_errors_and_state():
    from oscen_b200 import abi
    e = engine(O.K_FM, 16, 8)
    with pytest.raises(abi.EngineError):
        e.set_value(99, 1.0)                                  # bad input id
    with pytest.raises(abi.EngineError):
        e.push_events([(16, 0, GATE, 1.0)])                    # voice out of range
    with pytest.raises(abi.EngineError):
        e.process_block(513)                                   # > MAX_BLOCK
    e.set_value(1, np.linspace(1.0, 2.0, 16).astype(np.float32))   # op3_ratio now per voice
    with pytest.raises(abi.EngineError):
        e.set_value_ramp(1, 3.0, 100)                          # a ramp is one ValueRampState: needs a uniform input
    e.set_value(1, 2.0)                                        # uniform again
    e.set_value_ramp(1, 3.0, 100)
    e.process_block(32)
    st = e.state_download()
    assert st.shape == (29, 16) and np.isfinite(st).all()
    s = e.stats()
    assert s["blocks"] == 1 and s["kernel_launches"] > 0


def test_same_frame_note_on_and_off_and_retrigger():
    ref, gpu = both(O.K_FM, 8, 8)

    def script(b):
        out = []
        b.push_midi(0, 10, [0x90, 60, 100]); b.push_midi(0, 10, [0x80, 60, 0])     # on and off in one frame
        b.push_midi(0, 20, [0x90, 64, 90]); b.push_midi(0, 21, [0x90, 64, 30])     # same note twice
        b.process_block(128); out.append(b.read_voice_out())
        b.push_midi(0, 0, [0x80, 64, 0])
        b.process_block(128); out.append(b.read_voice_out())
        return np.concatenate(out)

    yr, yg = drive((ref, gpu), script)
    assert np.abs(yr).max() > 1e-3
    assert np.array_equal(yg, yr)


# BASELINE configs 2-5 at full size: tests/test_fullsize_gpu.py (strided samples over all CTAs + the whole bank's mix bus)


# --------------------------------------------------------------- plan compiler: node plans of the other graphs
@pytest.mark.parametrize("kind", [O.K_SYNTH, O.K_SYNTH_ADSR, O.K_SAT1, O.K_TWIN])
def test_node_plans_match_fused_and_oracle(kind):
    n = 192
    ref = O.OracleBank(kind, n, 1)
    fused, nodes = engine(kind, n, 1, plan=0), engine(kind, n, 1, plan=1)
    x = np.stack([lcg_noise(512, 77 + v) for v in range(n)], axis=1) if kind == O.K_TWIN else None

    def script(b):
        out = []
        if kind in (O.K_SYNTH, O.K_SYNTH_ADSR):
            b.set_value(2, (80.0 + 9.0 * np.arange(n)).astype(np.float32))
            b.set_value(3, (400.0 + 50.0 * np.arange(n)).astype(np.float32))
        elif kind == O.K_SAT1:
            b.set_value(0, (500.0 + 20.0 * np.arange(n)).astype(np.float32))
        for blk in range(2):
            if kind == O.K_TWIN:
                b.stream_in(0, x[blk * 256:(blk + 1) * 256])
            if kind == O.K_SYNTH_ADSR:
                b.push_events([(v, (v + 3 * blk) % 200, GATE, 0.9 if blk == 0 else 0.0) for v in range(n)])
            if blk == 1 and kind != O.K_SAT1:
                b.set_value_ramp(0, 7.0 if kind != O.K_TWIN else 1200.0, 100)
                b.push_events([(5, 40, 1 if kind != O.K_TWIN else 1, 0.3 if kind != O.K_TWIN else 2500.0)])   # a value event
            b.process_block(256)
            out.append((b.read_voice_out(), b.read_mix()))
        return out

    ro, fo, no = drive((ref, fused, nodes), script)
    for (vr, mr), (vf, mf), (vn, mn) in zip(ro, fo, no):
        assert np.array_equal(vf, vr) and np.array_equal(vn, vr)
        assert np.allclose(mn, mr, rtol=1e-5, atol=1e-4)


def test_twin_peaks_node_plan_with_device_lcg_source():
    n = 130
    fused, nodes = engine(O.K_TWIN, n, 1, plan=0), engine(O.K_TWIN, n, 1, plan=1)
    for e in (fused, nodes):
        e.stream_in_lcg(0, 4242, 11)
    for _ in range(3):
        fused.process_block(200); nodes.process_block(200)
        assert np.array_equal(fused.read_voice_out(), nodes.read_voice_out())


# --------------------------------------------------------------- cross-rate clipper graphs
CLIP_KINDS = [("ref", O.K_CLIP_REF)] + [(f"{name}{f}x", O.k_clip_os(p, f))
                                        for name, p in (("sinc", O.POLICY_SINC), ("iir", O.POLICY_SINC_IIR),
                                                        ("linear", O.POLICY_LINEAR), ("latch", O.POLICY_LATCH))
                                        for f in (2, 4, 8)]


@pytest.mark.parametrize("name,kind", CLIP_KINDS)
def test_clip_oversampled_family_bit_exact(name, kind):
    """ClipRef / ClipOversampled{,Iir} and the [linear]/[latch] variants, N in {2,4,8}: every resampler
    is +,* only, so the CUDA path must be bit-identical to the oracle; three ragged blocks carry the
    FIR rings / all-pass states across block boundaries"""
    n = 200
    ref, gpu = both(kind, n, 2)
    x = (1.4 * np.stack([lcg_noise(600, 500 + v) for v in range(n)], axis=1)).astype(np.float32)
    x[300:340] *= np.float32(1e-20)                  # a quiet stretch: the IIR denormal snap path
    assert gpu.latency_samples() == O.lib().orc_bank_latency(ref.h)
    pos = 0
    for frames in (256, 37, 307):
        for b in (ref, gpu):
            b.stream_in(0, x[pos:pos + frames]); b.process_block(frames)
        assert np.array_equal(gpu.read_voice_out(), ref.read_voice_out()), (name, pos)
        assert np.array_equal(gpu.read_instance_out(), ref.read_instance_out()), (name, pos)
        pos += frames
    assert np.abs(ref.read_voice_out()).max() > 0.3


def test_clip_oversampled_device_lcg_and_silence():
    n = 130
    kind = O.k_clip_os(O.POLICY_SINC_IIR, 8)
    ref, gpu = both(kind, n, 1)
    x = np.stack([lcg_noise(256, 42 + 3 + v) for v in range(n)], axis=1)
    ref.stream_in(0, x); ref.process_block(256)
    gpu.stream_in_lcg(0, 42, 3); gpu.process_block(256)
    assert np.array_equal(gpu.read_voice_out(), ref.read_voice_out())
    # 1000+ frames of silence: the all-pass states snap to exact zero (resample_kernels.rs:299-350)
    z = np.zeros((512, n), dtype=np.float32)
    for _ in range(3):
        for b in (ref, gpu):
            b.stream_in(0, z); b.process_block(512)
    assert np.array_equal(gpu.read_voice_out(), ref.read_voice_out())
    assert (gpu.read_voice_out()[-1] == 0.0).all()


def test_clip_oversampled_full_size_1m_voices_round_trip_properties():
    """1 048 576 voices of ClipOversampled 4x [sinc]: slice parity against the oracle plus
    size-independent properties (|out| bounded by the clip level times the FIR's gain, linearity of
    the un-clipped path: a small input passes through both FIRs scaled, no clipping)"""
    n, frames = 1048576, 256
    kind = O.k_clip_os(O.POLICY_SINC, 4)
    gpu = engine(kind, n, 1, keep=1)
    gpu.stream_in_lcg(0, 9000, 0)
    outs = []
    for _ in range(2):
        gpu.process_block(frames)
        outs.append((gpu.read_voice_out(0, 64), gpu.read_voice_out(n - 64, 64)))
    for first, idx in ((0, 0), (n - 64, 1)):
        ref = O.OracleBank(kind, 64, 1)
        x = np.stack([lcg_noise(2 * frames, 9000 + first + v) for v in range(64)], axis=1)
        for b in range(2):
            ref.stream_in(0, x[b * frames:(b + 1) * frames]); ref.process_block(frames)
            assert np.array_equal(outs[b][idx], ref.read_voice_out()), (first, b)
    assert np.abs(outs[1][0]).max() < 1.2


# --------------------------------------------------------------- device-side MIDI front
def _random_midi(rng, n_inst, frames, per_inst_max, n_notes, malformed=0.0):
    """[(instance, frame, bytes)] in push order: note-ons, note-offs, velocity-0 note-ons, controller noise,
    some instances above the 32-slot queue, some offsets beyond the block.  malformed > 0: that share of the messages
    carries data bytes >= 0x80 (the reference's note / velocity are any u8, midi.rs:8-35: note 255 must still match its
    note-off, a velocity byte of 200 clamps to 1.0)"""
    msgs = []
    for inst in range(n_inst):
        for _ in range(int(rng.integers(0, per_inst_max + 1))):
            frame = int(rng.integers(0, frames + 30))
            note = int(rng.integers(40, 40 + n_notes))
            if malformed > 0.0 and rng.random() < malformed:
                note = 256 - 1 - (note - 40) % 8                     # 255, 254, ...: few distinct notes, so that note-offs match
            r = rng.random()
            if r < 0.55:
                data = [0x90 | int(rng.integers(0, 16)), note, int(rng.integers(1, 256 if malformed > 0.0 else 128))]
            elif r < 0.85:
                data = [0x80, note, 0]
            elif r < 0.93:
                data = [0x90, note, 0]
            elif r < 0.97:
                data = [0xB0, 7, 100]
            else:
                data = [0x90, note]                                  # short message: ignored by the parser
            msgs.append((inst, frame, data))
    order = rng.permutation(len(msgs))                               # interleave instances in push order
    return [msgs[i] for i in order]


@pytest.mark.parametrize("kind,vpi", [(O.K_FM, 8), (O.K_EPIANO, 16), (O.K_FM, 2)])
def test_device_midi_front_matches_oracle_and_host_front(kind, vpi):
    n_inst = 96
    n = n_inst * vpi
    ref = O.OracleBank(kind, n, vpi)
    host = engine(kind, n, vpi)
    dev = engine(kind, n, vpi)
    rng = np.random.default_rng(11)
    for blk, frames in enumerate((256, 100, 512, 256)):
        msgs = _random_midi(rng, n_inst, frames, 40 if blk == 0 else 6, 2 * vpi)
        for inst, frame, data in msgs:
            ref.push_midi(inst, frame, data); host.push_midi(inst, frame, data)
        dev.push_midi_batch(msgs)
        for b in (ref, host, dev):
            b.process_block(frames)
        yr, yh, yd = ref.read_voice_out(), host.read_voice_out(), dev.read_voice_out()
        assert np.array_equal(yh, yd), (blk, "device front != host front")
        assert_close(yd, yr, f"device MIDI front block {blk}")
        assert_close(dev.read_instance_out(), ref.read_instance_out(), f"device MIDI front instance_out block {blk}")
    assert np.abs(yr).max() > 1e-3
    assert dev.stats()["events_dropped"] == host.stats()["events_dropped"] > 0


def test_device_midi_front_mode_is_exclusive():
    from oscen_b200 import abi
    e = engine(O.K_FM, 16, 8)
    e.push_midi(0, 0, [0x90, 60, 100])
    with pytest.raises(abi.EngineError):
        e.push_midi_batch([(0, 0, [0x90, 62, 100])])
    e2 = engine(O.K_FM, 16, 8)
    e2.push_midi_batch([(1, 3, [0x90, 62, 100])])
    with pytest.raises(abi.EngineError):
        e2.push_midi(0, 0, [0x90, 60, 100])
    e2.push_events([(0, 0, GATE, 1.0)])
    with pytest.raises(abi.EngineError):
        e2.process_block(64)


def test_device_midi_front_full_size_epiano_131072_instances():
    """config 3 polyphony: 131 072 ElectricPiano instances x 16 voices take a chord each through the
    device-side front (no host loop over instances); slices are checked against the oracle"""
    from oscen_b200 import abi
    n_inst, vpi, frames = 131072, 16, 256
    gpu = engine(O.K_EPIANO, n_inst * vpi, vpi, keep=2)
    rng = np.random.default_rng(5)
    k = 3
    msgs = np.zeros(n_inst * k, dtype=abi.MIDI_DTYPE)
    msgs["instance"] = np.repeat(np.arange(n_inst), k)
    msgs["frame_offset"] = rng.integers(0, frames, size=n_inst * k)
    msgs["bytes"][:, 0] = 0x90
    msgs["bytes"][:, 1] = rng.integers(36, 96, size=n_inst * k)
    msgs["bytes"][:, 2] = rng.integers(1, 128, size=n_inst * k)
    msgs["len"] = 3
    offs = msgs.copy(); offs["bytes"][:, 0] = 0x80; offs["frame_offset"] = rng.integers(0, frames, size=n_inst * k)
    gpu.push_midi_batch(msgs); gpu.process_block(frames); y0 = gpu.read_instance_out().copy()
    gpu.push_midi_batch(offs); gpu.process_block(frames); y1 = gpu.read_instance_out().copy()
    assert np.isfinite(y0).all() and np.abs(y0).max() > 1e-3
    for first in (0, n_inst - 8):
        ref = O.OracleBank(O.K_EPIANO, 8 * vpi, vpi)
        for blk, (m, want) in enumerate(((msgs, y0), (offs, y1))):
            for row in m[first * k:(first + 8) * k]:
                ref.push_midi(int(row["instance"]) - first, int(row["frame_offset"]), [int(x) for x in row["bytes"]])
            ref.process_block(frames)
            assert_close(want[:, first:first + 8], ref.read_instance_out(), f"epiano device MIDI slice {first} block {blk}")
