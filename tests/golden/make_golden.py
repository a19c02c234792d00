#!/usr/bin/env python
"""Generates the golden waveform fixtures under tests/golden/ from the CPU ORACLE (oracle/oscen_oracle.c).

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

The reference is Rust (nightly) and cannot be built or imported in this environment, and its own tests
hold no end-to-end waveforms for the example graphs, so these fixtures are the oracle's output frozen at the commit
that pinned it on every known-answer vector of the reference (tests/test_oracle_golden.py, reference_kats.json).
They serve two purposes: (1) a CPU test fails if the oracle itself ever drifts (a different libm, a changed
compiler flag); (2) the GPU tests compare the CUDA path with committed bytes, not only with an oracle built on the
same box.  Each fixture stores its driving script's parameters; `cases()` below IS the script, shared with the tests.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

GATE = 0xFFFFFFFF
K_FM, K_TWIN, K_SAT1, K_SAT4, K_EPIANO, K_SYNTH, K_SYNTH_ADSR = range(7)


def lcg_noise(n, seed):
    state = (seed * 2862933555777941757 + 3037000493) & 0xFFFFFFFFFFFFFFFF
    out = np.empty(n, dtype=np.float32)
    for i in range(n):
        state = (state * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        out[i] = np.float32(np.float32(state >> 33) / np.float32(2147483648.0)) - np.float32(1.0)
    return out


def drive(bank, name):
    """the driving script of fixture `name` on any bank with the engine's method names (OracleBank or abi.Engine);
    returns the per-voice output [frames][voices] (epiano: the per-instance stereo output [frames][instances][2])"""
    out = []
    if name.startswith("fm_anchor"):
        # BASELINE configs[0]: FMVoice, 1 voice, 48 kHz, 1024 frames (two 512 chunks), gate on at 0, off at 768
        bank.set_value(0, 440.0)
        if name.endswith("feedback"):
            bank.set_value(3, 0.3); bank.set_value(10, 0.2); bank.set_value(20, 0.5); bank.set_value(27, 1500.0)
        bank.push_events([(0, 0, GATE, 1.0)]); bank.process_block(512); out.append(bank.read_voice_out())
        bank.push_events([(0, 256, GATE, 0.0)]); bank.process_block(512); out.append(bank.read_voice_out())
    elif name == "fm_bank_64":
        n = 64
        freq = (110.0 * (1 + np.arange(n) % 24)).astype(np.float32)
        bank.set_value(0, freq)
        bank.push_events([(v, v % 32, GATE, 0.5 + 0.005 * v) for v in range(n)])
        bank.process_block(256); out.append(bank.read_voice_out())
        bank.set_value_ramp(2, 0.8, 100)
        bank.push_events([(v, 3 * (v % 40), GATE, 0.0) for v in range(0, n, 2)])
        bank.process_block(256); out.append(bank.read_voice_out())
    elif name == "twin_peaks_16":
        n = 16
        x = np.stack([lcg_noise(512, 1000 + v) for v in range(n)], axis=1)
        bank.stream_in(0, x[:256]); bank.process_block(256); out.append(bank.read_voice_out())
        bank.set_value(0, (800.0 + 40.0 * np.arange(n)).astype(np.float32))
        bank.stream_in(0, x[256:]); bank.process_block(256); out.append(bank.read_voice_out())
    elif name in ("sat_1x_16", "sat_4x_16"):
        n = 16
        bank.set_value(0, (1000.0 + 150.0 * np.arange(n)).astype(np.float32))
        for frames in (256, 100, 156):
            bank.process_block(frames); out.append(bank.read_voice_out())
    elif name == "synth_adsr_16":
        n = 16
        bank.set_value(2, (55.0 * (1 + np.arange(n))).astype(np.float32))
        bank.set_value(3, (300.0 + 900.0 * np.arange(n)).astype(np.float32))
        bank.push_events([(v, v, GATE, 0.8) for v in range(n)])
        bank.process_block(512); out.append(bank.read_voice_out())
        bank.push_events([(v, 2 * v, GATE, 0.0) for v in range(n)])
        bank.process_block(256); out.append(bank.read_voice_out())
    elif name == "epiano_2x16":
        for inst in range(2):
            bank.push_midi(inst, 3 + inst, [0x90, 48 + 7 * inst, 100])
            bank.push_midi(inst, 40, [0x90, 60 + inst, 64])
        bank.process_block(256); out.append(bank.read_instance_out())
        bank.push_midi(0, 10, [0x80, 48, 0])
        bank.process_block(256); out.append(bank.read_instance_out())
    else:
        raise KeyError(name)
    return np.concatenate(out)


# fixture name -> (graph kind, voices, voices per instance)
CASES = {
    "fm_anchor_defaults": (K_FM, 1, 1), "fm_anchor_feedback": (K_FM, 1, 1), "fm_bank_64": (K_FM, 64, 8),
    "twin_peaks_16": (K_TWIN, 16, 1), "sat_1x_16": (K_SAT1, 16, 1), "sat_4x_16": (K_SAT4, 16, 1),
    "synth_adsr_16": (K_SYNTH_ADSR, 16, 1), "epiano_2x16": (K_EPIANO, 32, 16),
}


def main():
    import oracle_lib as O
    for name, (kind, n, vpi) in CASES.items():
        y = drive(O.OracleBank(kind, n, vpi, 48000.0), name)
        assert np.isfinite(y).all() and np.abs(y).max() > 1e-3, name
        np.savez_compressed(os.path.join(HERE, name + ".npz"), output=y, kind=kind, voices=n, voices_per_instance=vpi, sample_rate=48000.0)
        print(f"{name}: shape {y.shape}, peak {np.abs(y).max():.4f}")


if __name__ == "__main__":
    main()
