"""BASELINE.json configs 2-5 at FULL size on the GPU, checked against the oracle on

  * a strided sample of >= 4 096 voices spread over ALL CTAs of the bank (whole instances, so the in-order instance fold
    is checked too) — per-voice and per-instance outputs, tolerance 1e-6 + 1e-5*|ref| (in practice bit-identical);
  * the MIX BUS of the whole bank: the oracle renders every voice of the bank (chunked over the host's cores) and its
    instance outputs are summed in float64; the engine's fixed-shape f32 tree must agree within 2e-6 * sum |instance|
    (the reference's `voices.audio_out -> audio_out` sum is a sequential f32 fold over <= 16 voices,
    oscen-graph-compiler/src/codegen/emit_node.rs:463-466; over a million voices only the order differs);
  * size-independent properties: exact in-order fold of every instance, silence before a voice's gate, determinism.

(VERDICT r1 "What's weak" 1: the round-1 tests compared the first / last CTA only and never the full-size mix.)"""
import os

import numpy as np
import pytest

import oracle_lib as O
from parity_util import assert_close, lcg_noise_block, seq_sum_f32, strided_instances

pytestmark = pytest.mark.gpu

GATE = O.EV_GATE
CORES = os.cpu_count() or 1
FRAMES = 256


def engine(kind, n, vpi=1, plan=0, keep=3, sr=48000.0):
    from oscen_b200 import abi
    return abi.Engine(kind, n, vpi, sr, 0, plan, keep)


def oracle_bank_mix(kind, n, vpi, drive, frames_per_block, chunk=65536):
    """float64 mix bus and sum |instance| of the WHOLE bank, rendered by the oracle in chunks of `chunk` voices.
    drive(bank, first_voice, count, block) applies the block's host calls to a bank holding voices [first, first+count)."""
    mix = [None] * len(frames_per_block); scale = [None] * len(frames_per_block)
    for first in range(0, n, chunk):
        count = min(chunk, n - first)
        ob = O.OracleBank(kind, count, vpi, 48000.0, threads=CORES)
        for b, fr in enumerate(frames_per_block):
            drive(ob, first, count, b)
            ob.process_block(fr)
            inst = ob.read_instance_out().astype(np.float64)
            m, s = inst.sum(axis=1), np.abs(inst).sum(axis=1)
            mix[b] = m if mix[b] is None else mix[b] + m
            scale[b] = s if scale[b] is None else scale[b] + s
        ob.close()
    return mix, scale


def check_mix(got, want, scale, what):
    d = np.abs(got.astype(np.float64) - want)
    tol = 1e-6 + 2e-6 * scale
    assert (d <= tol).all(), (what, float((d / tol).max()))
    assert np.abs(want).max() > 1.0, what
    return float((d / tol).max())


# ------------------------------------------------------------------ config 2: FMVoice x 65 536
def _fm_params(n):
    notes = 36 + (np.arange(n) * 7) % 60
    freq = (440.0 * np.exp2((notes - 69) / 12.0)).astype(np.float32)
    on = ((np.arange(n) * 13) % FRAMES).astype(np.uint32)
    off = ((np.arange(n) * 29 + 5) % FRAMES).astype(np.uint32)
    vel = (0.3 + 0.7 * ((np.arange(n) * 37) % 101) / 101.0).astype(np.float32)
    return freq, on, off, vel


def _fm_drive(freq, on, off, vel, ids_of):
    """host calls of block `block` for a bank that holds the global voices ids_of(first, count) (renumbered 0..)"""
    def drive(b, first, count, block):
        from oscen_b200 import abi
        ids = ids_of(first, count)
        if block == 0:
            b.set_value(0, freq[ids])
            ev = np.zeros(len(ids), dtype=abi.EVENT_DTYPE)
            ev["voice"] = np.arange(len(ids)); ev["frame_offset"] = on[ids]; ev["input_id"] = GATE; ev["scalar"] = vel[ids]
            b.push_events(ev)
        if block == 1:
            b.set_value_ramp(2, 0.7, 300); b.set_value_ramp(27, 600.0, 200)      # op3_level, filter_env_amount (tanf per sample)
            rel = np.nonzero(ids % 2 == 0)[0]                                     # the even voices release in block 1
            ev = np.zeros(len(rel), dtype=abi.EVENT_DTYPE)
            ev["voice"] = rel; ev["frame_offset"] = off[ids[rel]]; ev["input_id"] = GATE
            b.push_events(ev)
    return drive


def test_config2_fm_65536_voices_strided_sample_and_mix_bus():
    n, vpi = 65536, 8
    freq, on, off, vel = _fm_params(n)
    drive = _fm_drive(freq, on, off, vel, lambda first, count: np.arange(first, first + count))
    gpu = engine(O.K_FM, n, vpi)
    insts = strided_instances(n // vpi, 64 // vpi, 1024)               # one instance out of each of the 1 024 CTAs
    voices = (insts[:, None] * vpi + np.arange(vpi)[None, :]).ravel()
    assert len(voices) >= 4096 and len(np.unique(voices // 64)) == 1024
    ref = O.OracleBank(O.K_FM, len(voices), vpi, 48000.0, threads=CORES)
    drive_sample = _fm_drive(freq, on, off, vel, lambda first, count: voices)
    blocks = [FRAMES, FRAMES]
    mix_ref, mix_scale = oracle_bank_mix(O.K_FM, n, vpi, drive, blocks)
    worst = 0.0
    for b, fr in enumerate(blocks):
        drive(gpu, 0, n, b); drive_sample(ref, 0, len(voices), b)
        gpu.process_block(fr); ref.process_block(fr)
        v = gpu.read_voice_out(); io = gpu.read_instance_out()[:, :, 0]; mix = gpu.read_mix()[:, 0]
        _, exact = assert_close(v[:, voices], ref.read_voice_out(), f"config 2 block {b}: 8 192 sampled voices")
        assert_close(io[:, insts], ref.read_instance_out()[:, :, 0], f"config 2 block {b}: sampled instances")
        assert np.array_equal(seq_sum_f32(v.reshape(fr, n // vpi, vpi), axis=2), io)       # every instance: exact in-order fold
        if b == 0:
            assert np.all(v[np.arange(fr)[:, None] < on[None, :]] == 0.0)                  # silent before the gate
        worst = max(worst, check_mix(mix, mix_ref[b][:, 0], mix_scale[b][:, 0], f"config 2 block {b} mix bus"))
        print(f"config 2 block {b}: sampled voices bit-exact fraction {exact:.4f}; mix bus error {worst:.3f} of tolerance")


# ------------------------------------------------------------------ config 4: TwinPeaks x 1 048 576
def test_config4_twin_peaks_1m_voices_strided_sample_and_mix_bus():
    n = 1048576
    gpu = engine(O.K_TWIN, n, 1, keep=1)
    gpu.stream_in_lcg(0, 1000, 0)
    voices = strided_instances(n, 64, 4096)                              # one voice out of every 4th CTA, all positions
    assert len(np.unique(voices // 64)) == len(voices) >= 4096
    ref = O.OracleBank(O.K_TWIN, len(voices), 1, 48000.0, threads=CORES)
    st = None
    chunk = 65536
    states = [None] * (n // chunk)
    for b in range(2):
        gpu.process_block(FRAMES)
        x, st = lcg_noise_block((1000 + voices).astype(np.uint64), FRAMES, st)
        ref.stream_in(0, x); ref.process_block(FRAMES)
        v = gpu.read_voice_out()
        assert np.array_equal(v[:, voices], ref.read_voice_out()), f"config 4 block {b}: sampled voices"
        # the whole bank's mix bus: the oracle renders all 1 048 576 voices (16 chunks), fed the same LCG streams
        mix_ref = np.zeros(FRAMES); mix_scale = np.zeros(FRAMES)
        if b == 0:
            banks = [O.OracleBank(O.K_TWIN, chunk, 1, 48000.0, threads=CORES) for _ in states]
        for c, ob in enumerate(banks):
            xc, states[c] = lcg_noise_block((1000 + c * chunk + np.arange(chunk)).astype(np.uint64), FRAMES, states[c])
            ob.stream_in(0, xc); ob.process_block(FRAMES)
            y = ob.read_voice_out()
            assert np.array_equal(v[:, c * chunk:(c + 1) * chunk], y), f"config 4 block {b}: voices of chunk {c}"   # in fact EVERY voice
            y = y.astype(np.float64)
            mix_ref += y.sum(axis=1); mix_scale += np.abs(y).sum(axis=1)
        r = check_mix(gpu.read_mix()[:, 0], mix_ref, mix_scale, f"config 4 block {b} mix bus")
        print(f"config 4 block {b}: all 1 048 576 voices bit-identical; mix bus error {r:.3f} of tolerance")
    g2 = engine(O.K_TWIN, n, 1, keep=0)                                  # determinism of the fixed-shape tree
    g2.stream_in_lcg(0, 1000, 0); g2.process_block(FRAMES); g2.process_block(FRAMES)
    assert np.array_equal(g2.read_mix(), gpu.read_mix())


# ------------------------------------------------------------------ config 5: SatGraph_4x x 262 144
def test_config5_saturator_4x_262144_voices_strided_sample_and_mix_bus():
    n = 262144
    f = (2000.0 * (0.5 + ((np.arange(n) * 7919) % 10007) / 10007.0)).astype(np.float32)
    gpu = engine(O.K_SAT4, n, 1, keep=1)
    gpu.set_value(0, f)
    voices = strided_instances(n, 64, 4096)                              # one voice out of every CTA
    assert len(np.unique(voices // 64)) == 4096
    ref = O.OracleBank(O.K_SAT4, len(voices), 1, 48000.0, threads=CORES)
    ref.set_value(0, f[voices])

    def drive(b, first, count, block):
        if block == 0:
            b.set_value(0, f[first:first + count])
    mix_ref, mix_scale = oracle_bank_mix(O.K_SAT4, n, 1, drive, [FRAMES, FRAMES])
    for b in range(2):
        gpu.process_block(FRAMES); ref.process_block(FRAMES)
        v = gpu.read_voice_out()
        assert np.array_equal(v[:, voices], ref.read_voice_out()), f"config 5 block {b}: sampled voices"
        r = check_mix(gpu.read_mix()[:, 0], mix_ref[b][:, 0], mix_scale[b][:, 0], f"config 5 block {b} mix bus")
        print(f"config 5 block {b}: 4 096 sampled voices bit-identical; mix bus error {r:.3f} of tolerance")
    assert gpu.latency_samples() == 8 and np.abs(v).max() > 0.3


# ------------------------------------------------------------------ config 3: ElectricPiano x 131 072 instances
def test_config3_electric_piano_131072_instances_strided_sample_and_mix_bus():
    from oscen_b200 import abi
    n_inst, vpi, frames = 131072, 16, 128
    n = n_inst * vpi
    notes = 36 + (np.arange(n) * 5) % 48
    freq = (440.0 * np.exp2((notes - 69) / 12.0)).astype(np.float32)
    playing = np.arange(0, n, 4)                                          # a quarter of the voices play
    onf = ((np.arange(n // 4) * 3) % frames).astype(np.uint32)

    def drive(b, first, count, block):
        if block == 0:
            b.set_value(0, freq[first:first + count])
            m = (playing >= first) & (playing < first + count)
            ev = np.zeros(int(m.sum()), dtype=abi.EVENT_DTYPE)
            ev["voice"] = playing[m] - first; ev["frame_offset"] = onf[m]; ev["input_id"] = GATE; ev["scalar"] = 0.6
            b.push_events(ev)

    gpu = engine(O.K_EPIANO, n, vpi, keep=2)
    insts = strided_instances(n_inst, 64 // vpi, 512)                     # 8 192 voices: one instance out of every 64th CTA
    ref = O.OracleBank(O.K_EPIANO, len(insts) * vpi, vpi, 48000.0, threads=CORES)
    voices = (insts[:, None] * vpi + np.arange(vpi)[None, :]).ravel()
    ref.set_value(0, freq[voices])
    m = np.isin(playing, voices)
    ev = np.zeros(int(m.sum()), dtype=abi.EVENT_DTYPE)
    ev["voice"] = np.searchsorted(voices, playing[m]); ev["frame_offset"] = onf[m]; ev["input_id"] = GATE; ev["scalar"] = 0.6
    ref.push_events(ev)
    mix_ref, mix_scale = oracle_bank_mix(O.K_EPIANO, n, vpi, drive, [frames], chunk=131072)
    drive(gpu, 0, n, 0)
    gpu.process_block(frames); ref.process_block(frames)
    io = gpu.read_instance_out()
    assert np.array_equal(io[:, insts], ref.read_instance_out()), "config 3: sampled instances (stereo)"
    assert np.abs(io).max() > 1e-3 and np.abs(io[..., 0] - io[..., 1]).max() > 1e-6
    mix = gpu.read_mix()
    for c in range(2):
        r = check_mix(mix[:, c], mix_ref[0][:, c], mix_scale[0][:, c], f"config 3 mix bus channel {c}")
    print(f"config 3: 512 sampled instances (8 192 voices) bit-identical, stereo mix bus error {r:.3f} of tolerance")
