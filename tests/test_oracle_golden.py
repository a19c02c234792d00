"""Pins the CPU oracle on every known-answer vector the reference's own tests hold for the
hot path.  Citations: path:line under /root/reference."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle_lib as O

L = O.lib()
f32 = np.float32


# ---- (1) TptFilter impulse response, filters/tpt/mod.rs:152-161,227-264 (eps 1e-6)
TPT_IMPULSE = [0.014401104, 0.052318562, 0.089890145, 0.11065749, 0.11862421, 0.11729243,
               0.10961619, 0.098000914]


def _tpt(cutoff=2000.0, q=0.707, sr=48000.0):
    f = O.Tpt()
    L.orc_tpt_new(C.byref(f), cutoff, q)
    L.orc_tpt_set_sample_rate(C.byref(f), sr)
    L.orc_tpt_prepare(C.byref(f))
    return f


def test_tpt_impulse_response_matches_reference():
    f = _tpt()
    f.cutoff, f.q, f.f_mod = 2000.0, 0.707, 0.0
    for n, want in enumerate(TPT_IMPULSE):
        f.input = 1.0 if n == 0 else 0.0
        L.orc_tpt_process(C.byref(f))
        assert abs(f.output - want) <= 1e-6, (n, f.output, want)


# ---- (2) coefficient identities, filters/tpt/mod.rs:201-224
def test_tpt_coefficients_follow_zavalishin():
    f = _tpt()
    sr = f32(48000.0)
    period = f32(0.5) / sr
    freq = f32(f.current_cutoff)
    arg = f32(f32(f32(f32(2.0) * f32(math.pi)) * freq) * period)
    fw = f32(f32(f32(2.0) * sr) * f32(math.tan(float(arg)))) * period
    r = f32(1.0) / f32(f.current_q)
    d = f32(1.0) / (f32(1.0) + r * fw + fw * fw)
    assert abs(f.g - fw) <= 1e-6 and abs(f.h - d) <= 1e-6
    assert abs(f.r - r) <= 1e-6 and abs(f.k - (f.g + f.r)) <= 1e-6


# ---- (3) ValueRampState exact values, graph/types.rs:379-503
def _ramp(v):
    r = O.Ramp()
    L.orc_ramp_new(C.byref(r), v)
    return r


def test_ramp_new_and_immediate():
    r = _ramp(100.0)
    assert (r.current, r.target) == (100.0, 100.0) and not L.orc_ramp_is_ramping(C.byref(r))
    L.orc_ramp_set_immediate(C.byref(r), 50.0)
    assert (r.current, r.target) == (50.0, 50.0) and not L.orc_ramp_is_ramping(C.byref(r))
    L.orc_ramp_set_with_ramp(C.byref(r), 100.0, 0)
    assert (r.current, r.target) == (100.0, 100.0) and not L.orc_ramp_is_ramping(C.byref(r))


def test_ramp_tick_advances_and_lands_exactly():
    r = _ramp(0.0)
    L.orc_ramp_set_with_ramp(C.byref(r), 100.0, 4)
    assert r.current == 0.0 and r.target == 100.0 and L.orc_ramp_is_ramping(C.byref(r))
    for want in (25.0, 50.0, 75.0):
        assert L.orc_ramp_tick(C.byref(r)) == 0
        assert abs(r.current - want) < 0.001 and L.orc_ramp_is_ramping(C.byref(r))
    assert L.orc_ramp_tick(C.byref(r)) == 1
    assert r.current == 100.0 and not L.orc_ramp_is_ramping(C.byref(r))
    # tick does nothing when idle
    r = _ramp(42.0)
    assert L.orc_ramp_tick(C.byref(r)) == 0 and r.current == 42.0


def test_ramp_lands_on_target_100_frames():
    r = _ramp(0.0)
    L.orc_ramp_set_with_ramp(C.byref(r), 1.0, 100)
    for i in range(100):
        assert L.orc_ramp_tick(C.byref(r)) == (1 if i == 99 else 0)
    assert r.current == 1.0


def test_ramp_downward_and_interrupt():
    r = _ramp(100.0)
    L.orc_ramp_set_with_ramp(C.byref(r), 0.0, 4)
    for want in (75.0, 50.0, 25.0):
        assert L.orc_ramp_tick(C.byref(r)) == 0 and abs(r.current - want) < 0.001
    assert L.orc_ramp_tick(C.byref(r)) == 1 and r.current == 0.0
    r = _ramp(0.0)
    L.orc_ramp_set_with_ramp(C.byref(r), 100.0, 10)
    L.orc_ramp_tick(C.byref(r)); L.orc_ramp_tick(C.byref(r))
    cur = r.current
    L.orc_ramp_set_with_ramp(C.byref(r), 0.0, 4)
    assert r.current == cur and L.orc_ramp_is_ramping(C.byref(r))
    for i in range(4):
        assert L.orc_ramp_tick(C.byref(r)) == (1 if i == 3 else 0)
    assert r.current == 0.0


# ---- (4) MIDI, midi.rs:237-273
def test_midi_note_to_freq():
    assert L.orc_midi_note_to_freq(69) == 440.0
    assert abs(L.orc_midi_note_to_freq(60) - 261.626) < 0.01
    assert abs(L.orc_midi_note_to_freq(81) - 880.0) < 0.01


def _parse(data):
    b = (C.c_uint8 * len(data))(*data)
    note, vel = C.c_uint8(0), C.c_float(0)
    r = L.orc_midi_parse(b, len(data), C.byref(note), C.byref(vel))
    return r, note.value, vel.value


def test_midi_parse_cases():
    r, n, v = _parse([0x90, 60, 100])
    assert r == 1 and n == 60 and abs(v - 100.0 / 127.0) < 0.01
    assert _parse([0x80, 60, 0])[:2] == (2, 60)
    assert _parse([0x90, 60, 0])[:2] == (2, 60)      # note-on velocity 0 == note-off
    assert _parse([0x90, 60])[0] == 0                 # short message
    assert _parse([0xB0, 1, 2])[0] == 0


# ---- (5) VoiceAllocator exact indices, voice_allocator.rs:156-258
def _alloc(n):
    a = O.VoiceAlloc()
    L.orc_voice_alloc_new(C.byref(a), n)
    return a


def test_voice_allocation_and_stealing():
    a = _alloc(4)
    assert [L.orc_voice_alloc_allocate(C.byref(a), n) for n in (60, 64, 67, 72)] == [0, 1, 2, 3]
    assert all(a.active[i] for i in range(4))
    assert L.orc_voice_alloc_allocate(C.byref(a), 76) == 0 and a.note[0] == 76


def test_voice_find_and_release():
    a = _alloc(4)
    L.orc_voice_alloc_allocate(C.byref(a), 60); L.orc_voice_alloc_allocate(C.byref(a), 64)
    assert L.orc_voice_alloc_find(C.byref(a), 64) == 1
    L.orc_voice_alloc_release(C.byref(a), 1)
    assert a.active[1] and a.released[1]
    assert L.orc_voice_alloc_find(C.byref(a), 64) == -1


def test_voice_prefer_released_for_stealing():
    a = _alloc(4)
    for n in (60, 64, 67, 72):
        L.orc_voice_alloc_allocate(C.byref(a), n)
    L.orc_voice_alloc_release(C.byref(a), 1)
    assert L.orc_voice_alloc_allocate(C.byref(a), 76) == 1
    assert a.note[1] == 76 and not a.released[1]


def test_voice_releasing_continues_to_sound():
    a = _alloc(2)
    assert L.orc_voice_alloc_allocate(C.byref(a), 60) == 0
    L.orc_voice_alloc_release(C.byref(a), 0)
    assert a.active[0] and a.released[0]
    assert L.orc_voice_alloc_allocate(C.byref(a), 64) == 1
    assert L.orc_voice_alloc_allocate(C.byref(a), 67) == 0


# ---- (6) ADSR bands, envelope/adsr.rs:313-386
def _adsr(a, d, s, r, sr=48000.0):
    e = O.Adsr()
    L.orc_adsr_new(C.byref(e), a, d, s, r)
    L.orc_adsr_set_sample_rate(C.byref(e), sr)
    L.orc_adsr_prepare(C.byref(e))
    return e


def test_adsr_reaches_sustain_level():
    e = _adsr(0.01, 0.02, 0.6, 0.05)
    L.orc_adsr_on_gate(C.byref(e), 1.0)
    for _ in range(4800):
        L.orc_adsr_process(C.byref(e))
    assert 0.5 <= e.output <= 0.65


def test_adsr_release_returns_to_zero():
    e = _adsr(0.0, 0.0, 0.8, 0.01)
    L.orc_adsr_on_gate(C.byref(e), 1.0)
    for _ in range(100):
        L.orc_adsr_process(C.byref(e))
    L.orc_adsr_on_gate(C.byref(e), 0.0)
    for _ in range(4800):
        L.orc_adsr_process(C.byref(e))
    assert e.output <= 0.01


def test_adsr_velocity_scales_output():
    e = _adsr(0.0, 0.0, 1.0, 0.01)
    L.orc_adsr_on_gate(C.byref(e), 0.5)
    for _ in range(100):
        L.orc_adsr_process(C.byref(e))
    assert 0.45 <= e.output <= 0.55


# ---- (7) PolyBLEP, oscillators/mod.rs:240-302
def _pb(freq, amp, wave, sr=48000.0):
    o = O.PolyBlep()
    L.orc_polyblep_new(C.byref(o), freq, amp, wave)
    L.orc_polyblep_set_sample_rate(C.byref(o), sr)
    return o


def test_polyblep_saw_bounded():
    o = _pb(440.0, 1.0, 1)
    vals = []
    for _ in range(4800):
        L.orc_polyblep_process(C.byref(o))
        vals.append(o.output)
    assert min(vals) >= -1.25 and max(vals) <= 1.25


def test_polyblep_square_continuity():
    o = _pb(880.0, 0.8, 2)
    L.orc_polyblep_process(C.byref(o))
    prev = o.output
    for _ in range(1024):
        L.orc_polyblep_process(C.byref(o))
        assert abs(o.output - prev) <= 1.6
        prev = o.output


def test_polyblep_triangle_shape():
    o = _pb(220.0, 1.0, 3)
    s = []
    for _ in range(4):
        L.orc_polyblep_process(C.byref(o))
        s.append(o.output)
    assert abs(s[0]) < 0.25 and s[1] > s[0]


# ---- (9) halfband FIR resamplers, oscen-lib/tests/resample_kernels.rs:160-245,352-412
def _down(n):
    d = O.SincDown()
    L.orc_sinc_down_new(C.byref(d), n)
    return d


def _up(n):
    u = O.SincUp()
    L.orc_sinc_up_new(C.byref(u), n)
    return u


def test_sinc_fir_dc_unity_gain():
    u, out = _up(2), (C.c_float * 2)()
    for _ in range(200):
        L.orc_sinc_up_process(C.byref(u), 0.7, out)
    assert abs(out[0] - 0.7) < 1e-3 and abs(out[1] - 0.7) < 1e-3
    d, y = _down(2), 0.0
    xs = (C.c_float * 2)(0.7, 0.7)
    for _ in range(200):
        y = L.orc_sinc_down_process(C.byref(d), xs)
    assert abs(y - 0.7) < 1e-3


def _roundtrip_err(n, f, warmup):
    u, d = _up(n), _down(n)
    buf = (C.c_float * n)()
    lag = (L.orc_sinc_up_latency(C.byref(u)) + L.orc_sinc_down_latency(C.byref(d))) // n
    max_err = 0.0
    two_pi = f32(2.0) * f32(math.pi)
    for i in range(1024):
        x = f32(math.sin(float(f32(f32(two_pi * f32(f)) * f32(i)))))
        L.orc_sinc_up_process(C.byref(u), float(x), buf)
        y = L.orc_sinc_down_process(C.byref(d), buf)
        if i > warmup and i >= lag:
            want = math.sin(float(f32(f32(two_pi * f32(f)) * f32(i - lag))))
            max_err = max(max_err, abs(y - want))
    return max_err


def test_sinc_fir_passband_flat():
    assert _roundtrip_err(2, 0.1, 64) < 0.1
    e4 = _roundtrip_err(4, 0.05, 128)
    e8 = _roundtrip_err(8, 0.05, 128)
    assert e4 < 0.25 and e8 < 0.25
    # the reference records its observed values in comments (resample_kernels.rs:356,388)
    assert abs(e4 - 0.156) < 0.01, e4
    assert abs(e8 - 0.078) < 0.01, e8


def test_sinc_fir_stopband_attenuated():
    d = _down(2)
    peak = 0.0
    two_pi = 2.0 * math.pi
    for m in range(2048):
        xs = (C.c_float * 2)(math.sin(two_pi * 0.4 * (2 * m)), math.sin(two_pi * 0.4 * (2 * m + 1)))
        y = L.orc_sinc_down_process(C.byref(d), xs)
        if m > 128:
            peak = max(peak, abs(y))
    assert -20.0 * math.log10(max(peak, 1e-12)) > 50.0


def test_sinc_fir_latency():
    lat = [L.orc_sinc_up_latency(C.byref(_up(n))) for n in (2, 4, 8)]
    assert lat[0] > 0 and lat[1] >= lat[0] and lat[2] >= lat[1]
    assert L.orc_sinc_down_latency(C.byref(_down(4))) == 33      # sinc_fir.rs:250-260, SURVEY §8a a13
    assert L.orc_sinc_down_latency(C.byref(_down(1))) == 0


# ---- (9b) latch / linear / IIR-halfband resamplers, oscen-lib/tests/resample_kernels.rs:3-152,246-350,413-467
def test_latch_resamplers_exact_values():
    out2, out4 = (C.c_float * 2)(), (C.c_float * 4)()
    L.orc_latch_up_process(2, 1.0, out2); assert list(out2) == [1.0, 1.0]
    L.orc_latch_up_process(2, -0.25, out2); assert list(out2) == [-0.25, -0.25]
    L.orc_latch_up_process(4, 0.5, out4); assert list(out4) == [0.5] * 4
    assert L.orc_latch_down_process(2, (C.c_float * 2)(1.0, 2.0)) == 1.0
    assert L.orc_latch_down_process(2, (C.c_float * 2)(3.0, 4.0)) == 3.0
    assert L.orc_latch_down_process(4, (C.c_float * 4)(10.0, 11.0, 12.0, 13.0)) == 10.0


def _lin(n):
    u = O.LinearUp()
    L.orc_linear_up_new(C.byref(u), n)
    return u


def test_linear_resamplers_reference_values():
    u, out = _lin(2), (C.c_float * 2)()
    L.orc_linear_up_process(C.byref(u), 0.0, out)
    L.orc_linear_up_process(C.byref(u), 1.0, out)
    assert abs(out[0] - 0.0) < 1e-6 and abs(out[1] - 0.5) < 1e-6
    L.orc_linear_up_process(C.byref(u), 2.0, out)
    assert abs(out[0] - 1.0) < 1e-6 and abs(out[1] - 1.5) < 1e-6
    u4, out4 = _lin(4), (C.c_float * 4)()
    L.orc_linear_up_process(C.byref(u4), 0.7, out4)
    L.orc_linear_up_process(C.byref(u4), 0.7, out4)
    assert all(abs(v - 0.7) < 1e-6 for v in out4)
    assert [L.orc_linear_up_latency(n) for n in (2, 4, 8)] == [2, 4, 8]
    assert [L.orc_linear_down_latency(n) for n in (2, 4, 8)] == [0, 1, 3]
    # impulse peak of LinearUp<4> lands at dest index == latency (resample_kernels.rs:108-126)
    u4, resp = _lin(4), []
    for x in (1.0, 0.0):
        L.orc_linear_up_process(C.byref(u4), x, out4); resp += list(out4)
    assert int(np.argmax(resp)) == 4
    assert abs(L.orc_linear_down_process(2, (C.c_float * 2)(1.0, 3.0)) - 2.0) < 1e-6
    assert abs(L.orc_linear_down_process(4, (C.c_float * 4)(1.0, 2.0, 3.0, 4.0)) - 2.5) < 1e-6


def _iir(n):
    u, d = O.IirResampler(), O.IirResampler()
    L.orc_iir_up_new(C.byref(u), n); L.orc_iir_down_new(C.byref(d), n)
    return u, d


def test_iir_halfband_dc_gain_stopband_latency():
    u, d = _iir(2)
    out = (C.c_float * 2)()
    for _ in range(1000):
        L.orc_iir_up_process(C.byref(u), 0.5, out)
    assert abs(out[0] - 0.5) < 5e-3 and abs(out[1] - 0.5) < 5e-3
    y = 0.0
    for _ in range(1000):
        y = L.orc_iir_down_process(C.byref(d), (C.c_float * 2)(0.5, 0.5))
    assert abs(y - 0.5) < 5e-3
    _, d = _iir(2)
    peak, two_pi = 0.0, f32(2.0) * f32(math.pi)
    for m in range(4096):
        x0 = math.sin(float(f32(f32(two_pi * f32(0.4)) * f32(2 * m))))
        x1 = math.sin(float(f32(f32(two_pi * f32(0.4)) * f32(2 * m + 1))))
        y = L.orc_iir_down_process(C.byref(d), (C.c_float * 2)(x0, x1))
        if m > 256:
            peak = max(peak, abs(y))
    assert -20.0 * math.log10(max(peak, 1e-12)) > 40.0
    assert L.orc_iir_up_latency(C.byref(_iir(2)[0])) < L.orc_sinc_up_latency(C.byref(_up(2)))
    assert [L.orc_iir_down_latency(C.byref(_iir(n)[1])) for n in (1, 2, 4, 8)] == [0, 2, 6, 14]


def test_iir_halfband_state_snaps_to_exact_zero_after_silence():
    u, d = _iir(8)
    out = (C.c_float * 8)()
    for _ in range(100):
        L.orc_iir_up_process(C.byref(u), 0.1, out)
    for _ in range(1000):
        L.orc_iir_up_process(C.byref(u), 0.0, out)
    assert list(out) == [0.0] * 8
    y = 0.0
    for _ in range(100):
        y = L.orc_iir_down_process(C.byref(d), (C.c_float * 8)(*([0.1] * 8)))
    for _ in range(1000):
        y = L.orc_iir_down_process(C.byref(d), (C.c_float * 8)())
    assert y == 0.0


def _iir_roundtrip_err(n, f, warmup):
    u, d = _iir(n)
    buf = (C.c_float * n)()
    lag = (L.orc_iir_up_latency(C.byref(u)) + L.orc_iir_down_latency(C.byref(d))) // n
    max_err, two_pi = 0.0, f32(2.0) * f32(math.pi)
    for i in range(1024):
        x = f32(math.sin(float(f32(f32(two_pi * f32(f)) * f32(i)))))
        L.orc_iir_up_process(C.byref(u), float(x), buf)
        y = L.orc_iir_down_process(C.byref(d), buf)
        if i > warmup and i >= lag:
            want = math.sin(float(f32(f32(two_pi * f32(f)) * f32(i - lag))))
            max_err = max(max_err, abs(y - want))
    return max_err


def test_iir_halfband_passband_flat():
    e4, e8 = _iir_roundtrip_err(4, 0.03, 256), _iir_roundtrip_err(8, 0.03, 256)
    assert e4 < 0.1 and e8 < 0.15
    # the reference records its observed values in comments (resample_kernels.rs:416,446)
    assert abs(e4 - 0.030) < 0.005, e4
    assert abs(e8 - 0.059) < 0.005, e8


# ---- (9c) ClipRef / ClipOversampled / ClipOversampledIir graphs, oscen-lib/tests/multirate_graph.rs:282-436
def _bin_magnitude(x, freq):
    i = np.arange(len(x), dtype=np.float64)
    ph = 2.0 * math.pi * freq * i
    return math.hypot(float((x * np.cos(ph)).sum()), float((x * np.sin(ph)).sum())) / len(x)


def _clip_graph_output(kind, total=4096, block=256):
    b = O.OracleBank(kind, 1, 1)
    n = np.arange(total, dtype=np.float32)
    x = (np.float32(0.9) * np.sin(np.float32(2.0 * math.pi) * np.float32(0.2) * n)).astype(np.float32)
    out = []
    for s in range(0, total, block):
        b.stream_in(0, x[s:s + block, None])
        b.process_block(block)
        out.append(b.read_voice_out()[:, 0].copy())
    return np.concatenate(out)


@pytest.mark.parametrize("policy", [O.POLICY_SINC, O.POLICY_SINC_IIR])
def test_hardclip_4x_has_less_aliasing_than_1x(policy):
    a = _clip_graph_output(O.K_CLIP_REF)[512:]
    b = _clip_graph_output(O.k_clip_os(policy, 4))[512:]
    one_x = _bin_magnitude(a, 0.4) / max(_bin_magnitude(a, 0.2), 1e-9)
    four_x = _bin_magnitude(b, 0.4) / max(_bin_magnitude(b, 0.2), 1e-9)
    assert four_x < one_x * 0.5, (four_x, one_x)


def test_clip_graph_latencies():
    # latency_samples(): sum over Down edges of latency / factor (codegen/emit_struct.rs:533-566)
    want = {(O.POLICY_SINC, 2): 11 // 2, (O.POLICY_SINC, 4): 33 // 4, (O.POLICY_SINC, 8): 77 // 8,
            (O.POLICY_SINC_IIR, 2): 1, (O.POLICY_SINC_IIR, 4): 6 // 4, (O.POLICY_SINC_IIR, 8): 14 // 8,
            (O.POLICY_LINEAR, 2): 0, (O.POLICY_LINEAR, 4): 0, (O.POLICY_LINEAR, 8): 0,
            (O.POLICY_LATCH, 2): 0, (O.POLICY_LATCH, 4): 0, (O.POLICY_LATCH, 8): 0}
    for (p, f), w in want.items():
        b = O.OracleBank(O.k_clip_os(p, f), 1, 1)
        assert L.orc_bank_latency(b.h) == w, (p, f)
    assert L.orc_bank_latency(O.OracleBank(O.K_CLIP_REF, 1, 1).h) == 0


def test_clip_latch_equals_reference_clipper():
    # [latch] up then [latch] down around a memoryless node is the 1x graph exactly
    a = _clip_graph_output(O.K_CLIP_REF, 512)
    for f in (2, 4, 8):
        assert np.array_equal(a, _clip_graph_output(O.k_clip_os(O.POLICY_LATCH, f), 512))


# ---- (8) block == per-sample, oscen-lib/tests/block_processing_test.rs:23-286
def _script(bank, frames_list, per_sample):
    """gate on at frame 5 of block 0, value event + ramp in block 1, gate off in block 2"""
    outs = []
    for bi, frames in enumerate(frames_list):
        evs = []
        if bi == 0:
            evs = [(v, 5 + v, O.EV_GATE, 0.5 + 0.1 * v) for v in range(bank.n_voices)]
        if bi == 1:
            bank.set_value_ramp(2, 0.9, 100)        # op3_level ramp
            evs = [(0, 17, 0, 330.0)]                # frame-accurate frequency change
        if bi == 2:
            evs = [(v, 40, O.EV_GATE, 0.0) for v in range(bank.n_voices)]
        if not per_sample:
            bank.push_events(evs)
            bank.process_block(frames)
            outs.append(bank.read_voice_out())
        else:
            for f in range(frames):
                bank.push_events([(v, 0, i, s) for (v, fo, i, s) in evs if fo == f])
                bank.process_block(1)
                outs.append(bank.read_voice_out())
    return np.concatenate(outs, axis=0)


def test_block_equals_per_sample_bit_exact():
    a = O.OracleBank(O.K_FM, 4, 4)
    b = O.OracleBank(O.K_FM, 4, 4)
    ya = _script(a, [64, 128, 96], False)
    yb = _script(b, [64, 128, 96], True)
    assert np.abs(ya).max() > 1e-3
    assert np.array_equal(ya, yb)


def test_block_split_invariance():
    a = O.OracleBank(O.K_SYNTH_ADSR, 2, 1)
    b = O.OracleBank(O.K_SYNTH_ADSR, 2, 1)
    a.push_events([(0, 3, O.EV_GATE, 1.0), (1, 200, O.EV_GATE, 0.7)])
    a.process_block(256)
    ya = a.read_voice_out()
    b.push_events([(0, 3, O.EV_GATE, 1.0)])
    b.process_block(100)
    y1 = b.read_voice_out()
    b.push_events([(1, 100, O.EV_GATE, 0.7)])
    b.process_block(156)
    yb = np.concatenate([y1, b.read_voice_out()])
    assert np.array_equal(ya, yb)


# ---- (10) example smoke tests: note-on => audible output
# examples/pivot/src/main.rs:599-725, examples/electric-piano/src/main.rs:328-389
def test_fm_graph_midi_note_makes_sound():
    b = O.OracleBank(O.K_FM, 8, 8)
    b.push_midi(0, 0, [0x90, 60, 100])
    peak = 0.0
    for _ in range(16):
        b.process_block(512)
        peak = max(peak, float(np.abs(b.read_instance_out()).max()))
    assert peak > 1e-4


def test_epiano_midi_note_makes_stereo_sound():
    b = O.OracleBank(O.K_EPIANO, 16, 16)
    b.push_midi(0, 0, [0x90, 60, 100])
    y = []
    for _ in range(16):
        b.process_block(512)
        y.append(b.read_instance_out())
    y = np.concatenate(y)
    assert np.abs(y).max() > 1e-4
    assert np.abs(y[:, 0, 0] - y[:, 0, 1]).max() > 1e-6       # L/R diverge (tremolo)


def test_sat_4x_reduces_aliasing_and_has_latency():
    # oscen-lib/tests/multirate_graph.rs:139-152,326-380: latency > 0; the 4x path differs from 1x
    a = O.OracleBank(O.K_SAT1, 1, 1)
    b = O.OracleBank(O.K_SAT4, 1, 1)
    assert L.orc_bank_latency(b.h) > 0 and L.orc_bank_latency(a.h) == 0
    a.process_block(512); b.process_block(512)
    ya, yb = a.read_voice_out()[:, 0], b.read_voice_out()[:, 0]
    assert np.abs(ya).max() <= 0.7 + 1e-6 and np.abs(yb).max() < 1.0
    assert np.abs(yb[:8]).max() < 0.05                      # FIR latency: output starts near zero
    assert np.abs(yb[64:]).max() > 0.3


def test_twin_peaks_runs():
    b = O.OracleBank(O.K_TWIN, 3, 1)
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, size=(256, 3)).astype(np.float32)
    b.stream_in(0, x)
    b.process_block(256)
    y = b.read_voice_out()
    assert np.isfinite(y).all() and np.abs(y).max() > 1e-3


def test_multithreaded_oracle_is_identical():
    a = O.OracleBank(O.K_FM, 64, 8)
    b = O.OracleBank(O.K_FM, 64, 8, threads=4)
    for bank in (a, b):
        bank.push_events([(v, v % 7, O.EV_GATE, 1.0) for v in range(64)])
        bank.process_block(128)
    assert np.array_equal(a.read_voice_out(), b.read_voice_out())
    assert np.array_equal(a.read_mix(), b.read_mix())


# ------------------------------------------------------------------ RingBuffer / Delay / IirLowpass
def _ring(size):
    L = O.lib()
    r = O.Ring()
    L.orc_ring_init(C.byref(r), size)
    return L, r


def test_ring_buffer_capacity_is_next_power_of_two():
    """ring_buffer/tests.rs:7-36,236-249 (PowerOfTwo mode, the default the Delay node uses)"""
    for size, cap in ((5, 8), (8, 8), (9, 16), (17, 32), (0, 1), (40, 64), (10, 16), (88200, 131072)):
        L, r = _ring(size)
        assert r.capacity == cap and r.mask == cap - 1 and r.write_pos == 0
        L.orc_ring_free(C.byref(r))


def test_ring_buffer_push_wrap_and_exact_offsets():
    """ring_buffer/tests.rs:62-86 (push/wrap) and :113-143 (get at whole offsets, wrap past the capacity, negative -> 0)"""
    L, r = _ring(4)
    for v in (1.0, 2.0, 3.0, 4.0):
        L.orc_ring_push(C.byref(r), v)
    assert r.write_pos == 0 and r.buffer[0] == 1.0 and r.buffer[3] == 4.0
    L.orc_ring_push(C.byref(r), 5.0)
    assert r.write_pos == 1 and r.buffer[0] == 5.0 and r.buffer[1] == 2.0
    L.orc_ring_push(C.byref(r), 6.0)
    assert r.write_pos == 2 and r.buffer[1] == 6.0
    # buffer [5, 6, 3, 4], newest = 6
    for off, want in ((0.0, 6.0), (1.0, 5.0), (2.0, 4.0), (3.0, 3.0), (4.0, 6.0), (-1.0, 6.0), (1.0000005, 5.0), (0.9999995, 5.0)):
        assert L.orc_ring_get(C.byref(r), off) == want, off


def test_ring_buffer_interpolation_goldens():
    """ring_buffer/tests.rs:146-170 (linear: capacity < 4) and :173-191 (Catmull-Rom: 13.1875 for 4, 8, 16, 1 at f = 0.5)"""
    L, r = _ring(2)
    L.orc_ring_push(C.byref(r), 5.0); L.orc_ring_push(C.byref(r), 9.0)          # newest 9, then 5
    assert abs(L.orc_ring_get(C.byref(r), 0.5) - 7.0) < 1e-6
    L, r = _ring(8)
    for v in (0.0, 0.0, 4.0, 8.0, 16.0, 1.0, 0.0, 0.0):                         # write_pos back at 0
        L.orc_ring_push(C.byref(r), v)
    # rp = (0 - 3.5 - 1 + 8) % 8 = 3.5: neighbours 4, 8, 16, 1
    assert abs(L.orc_ring_get(C.byref(r), 3.5) - 13.1875) < 1e-6


def test_delay_node_delays_by_whole_samples_and_feeds_back():
    """delay/mod.rs:71-85: output = ring.get(delay_samples), then push(input + delayed * feedback)"""
    L = O.lib()
    d = O.Delay()
    L.orc_delay_new(C.byref(d), 3.0, 0.5)
    d.sample_rate = 48000.0
    L.orc_delay_prepare(C.byref(d))
    assert d.ring.capacity == 131072                                 # next_power_of_two(min(96000, 88200))
    out = []
    for n in range(12):
        d.input = 1.0 if n == 0 else 0.0
        L.orc_delay_process(C.byref(d))
        out.append(d.output)
    # get(3) reads the sample pushed 4 process() calls earlier; each echo comes back at half the level
    assert out == [0, 0, 0, 0, 1.0, 0, 0, 0, 0.5, 0, 0, 0]
    # parameters are clamped on update frames only (every 32nd): feedback 2.0 -> 0.99, delay -5 -> 0
    L.orc_delay_new(C.byref(d), -5.0, 2.0); d.sample_rate = 44100.0; L.orc_delay_prepare(C.byref(d))
    L.orc_delay_process(C.byref(d))
    assert d.delay_samples == 0.0 and abs(d.feedback - 0.99) < 1e-7
    L.orc_delay_free(C.byref(d))


def test_iir_lowpass_reference_tests():
    """filters/iir_lowpass/mod.rs:176-330: JUCE coefficient formula, DC gain ~1, bounded impulse response, high-Q
    stability, denormal snap"""
    L = O.lib()
    f = O.IirLp()
    q = np.float32(0.70710678)
    L.orc_iirlp_new(C.byref(f), 1000.0, q); f.sample_rate = 48000.0; L.orc_iirlp_prepare(C.byref(f))
    n = np.float32(1.0) / np.float32(math.tan(np.float32(np.float32(np.float32(math.pi) * np.float32(1000.0)) / np.float32(48000.0))))
    c1 = np.float32(1.0) / (np.float32(1.0) + np.float32(1.0) / q * n + n * n)
    assert abs(f.b0 - c1) < 1e-6 and abs(f.b1 - 2 * c1) < 1e-6 and abs(f.b2 - c1) < 1e-6
    assert abs(f.a1 - c1 * 2 * (1 - n * n)) < 1e-5 and abs(f.a2 - c1 * (1 - 1 / q * n + n * n)) < 1e-5
    f.frames_per_update = 1
    for _ in range(1000):
        f.input = 1.0; L.orc_iirlp_process(C.byref(f))
    assert abs(f.output - 1.0) < 0.01
    L.orc_iirlp_new(C.byref(f), 2000.0, q); f.sample_rate = 48000.0; L.orc_iirlp_prepare(C.byref(f)); f.frames_per_update = 1
    outs = []
    for k in range(8):
        f.input = 1.0 if k == 0 else 0.0; L.orc_iirlp_process(C.byref(f)); outs.append(f.output)
    assert outs[0] > 0 and all(abs(x) < 2.0 for x in outs[1:])
    L.orc_iirlp_new(C.byref(f), 1000.0, 10.0); f.sample_rate = 48000.0; L.orc_iirlp_prepare(C.byref(f)); f.frames_per_update = 1
    for k in range(100):
        f.input = 1.0 if k == 0 else 0.0; L.orc_iirlp_process(C.byref(f))
        assert abs(f.output) < 10.0
    L.orc_iirlp_new(C.byref(f), 100.0, q); f.sample_rate = 48000.0; L.orc_iirlp_prepare(C.byref(f))
    f.input = 1e-20; L.orc_iirlp_process(C.byref(f))
    assert f.output == 0.0


def _player_run(p, n):
    out = []
    for _ in range(n):
        O.lib().orc_sample_player_process(C.byref(p)); out.append(p.output)
    return np.array(out, dtype=np.float32)


def test_sample_player_reference_tests():
    """sample_player/tests.rs:35-71,138-155 (mono cases) and tests/sample_player_graph.rs:64-124: in-order looping playback,
    playhead reset on swap, silence when unloaded, channel 0 of a multi-channel source — the reference's epsilon is 1e-6,
    the copies are exact"""
    L = O.lib()
    f32 = np.float32

    def buf(v):
        return np.array(v, dtype=np.float32)

    p = O.SamplePlayer(); L.orc_sample_player_new(C.byref(p))
    a = buf([0.1, 0.2, 0.3])
    L.orc_sample_player_publish(C.byref(p), a.ctypes.data, a.size)
    assert np.array_equal(_player_run(p, 7), buf([0.1, 0.2, 0.3, 0.1, 0.2, 0.3, 0.1]))

    p = O.SamplePlayer(); L.orc_sample_player_new(C.byref(p))
    a, b = buf([0.1, 0.2, 0.3, 0.4]), buf([0.8, 0.9])
    L.orc_sample_player_publish(C.byref(p), a.ctypes.data, a.size)
    assert np.array_equal(_player_run(p, 2), buf([0.1, 0.2]))
    L.orc_sample_player_publish(C.byref(p), b.ctypes.data, b.size)
    assert np.array_equal(_player_run(p, 3), buf([0.8, 0.9, 0.8]))

    p = O.SamplePlayer(); L.orc_sample_player_new(C.byref(p))
    assert (_player_run(p, 4) == f32(0.0)).all()

    inter = buf([0.1, 0.9, 0.2, 0.8])                                 # L = [0.1, 0.2], R = [0.9, 0.8]: a downmix would give 0.5
    mono = np.zeros(2, dtype=np.float32)
    L.orc_sample_player_build_mono(mono.ctypes.data, inter.ctypes.data, 2, 2)
    p = O.SamplePlayer(); L.orc_sample_player_new(C.byref(p))
    L.orc_sample_player_publish(C.byref(p), mono.ctypes.data, 2)
    assert np.array_equal(_player_run(p, 2), buf([0.1, 0.2]))

    # the graph test: silent before any load, A looped twice, then B looped twice from its start
    p = O.SamplePlayer(); L.orc_sample_player_new(C.byref(p))
    a, b = buf([0.1, -0.2, 0.3, -0.4, 0.5]), buf([0.9, 0.8, 0.7])
    assert (_player_run(p, 8) == f32(0.0)).all()
    L.orc_sample_player_publish(C.byref(p), a.ctypes.data, a.size)
    assert np.array_equal(_player_run(p, 10), np.concatenate([a, a]))
    L.orc_sample_player_publish(C.byref(p), b.ctypes.data, b.size)
    assert np.array_equal(_player_run(p, 6), np.concatenate([b, b]))


def test_bank_refuses_value_events_on_a_ramped_input():
    """a ramped graph input is one ValueRampState read through `.current` every frame (codegen/emit_node.rs:317-321): the
    bank API's per-voice value events have no meaning on it (found by tests/host_kernel.py: engine and oracle used to
    disagree once the ramp had completed) — refused on both sides, like a per-voice set_value on it"""
    b = O.OracleBank(O.K_SAT1, 3, 1)
    b.set_value(0, 900.0)
    b.push_events([(1, 4, 0, 700.0)])                     # fine while the input has never been ramped
    b.process_block(16)
    b.set_value_ramp(0, 2500.0, 8)
    with pytest.raises(Exception):
        b.push_events([(1, 4, 0, 700.0)])
    b.push_events([(1, 4, 1, 0.5)])                       # another input is unaffected
    b.process_block(16)
