#!/usr/bin/env python
"""Device-timed throughput of every graph kind at its BASELINE.json config size (fused plan).
Prints one JSON line per graph: voice-samples/s, kernel ms, Model-F bytes and HBM fraction."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oscen_b200 import abi  # noqa: E402

FRAMES = 256
PEAK = 6489.9
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

CASES = [
    # name, kind, voices, vpi, per-voice stream bytes/frame (in), state words, channels
    ("fm_voice 65536x8 (config 2)", abi.GRAPH_FM_VOICE, 65536, 8, 0, 29),
    ("twin_peaks 1048576 voices, device LCG audio_in (config 4)", abi.GRAPH_TWIN_PEAKS, 1048576, 1, 0, 20),
    ("twin_peaks 1048576 voices, audio_in stream read from HBM", abi.GRAPH_TWIN_PEAKS, 1048576, 1, 4, 20),
    ("sat_4x 262144 voices (config 5)", abi.GRAPH_SAT_4X, 262144, 1, 0, 49),
    ("sat_1x 262144 voices (control)", abi.GRAPH_SAT_1X, 262144, 1, 0, 49),
    ("epiano 131072 instances x 16 voices (config 3, reference polyphony)", abi.GRAPH_EPIANO, 131072 * 16, 16, 0, 259),
    ("epiano 131072 instances x 8 voices (config 3 as worded)", abi.GRAPH_EPIANO, 131072 * 8, 8, 0, 259),
    ("synth_adsr 1048576 voices (osc+filter+env)", abi.GRAPH_SYNTH_ADSR, 1048576, 1, 0, 13),
    # cross-rate clipper graphs: per-voice input stream read from HBM, per-voice output kept (8 B/voice-sample)
    ("clip_ref 1x 1048576 voices, stream in+out", abi.GRAPH_CLIP_REF, 1048576, 1, 4, 2),
    ("clip_os sinc 4x 1048576 voices, stream in+out (ClipOversampled)", abi.graph_clip_os(abi.POLICY_SINC, 4), 1048576, 1, 4, 74),
    ("clip_os sinc_iir 4x 1048576 voices, stream in+out (ClipOversampledIir)", abi.graph_clip_os(abi.POLICY_SINC_IIR, 4), 1048576, 1, 4, 36),
    ("clip_os sinc 8x 1048576 voices, stream in+out", abi.graph_clip_os(abi.POLICY_SINC, 8), 1048576, 1, 4, 110),
    ("clip_os sinc_iir 8x 1048576 voices, stream in+out", abi.graph_clip_os(abi.POLICY_SINC_IIR, 8), 1048576, 1, 4, 53),
    ("clip_os linear 4x 1048576 voices, stream in+out", abi.graph_clip_os(abi.POLICY_LINEAR, 4), 1048576, 1, 4, 3),
    ("clip_os latch 4x 1048576 voices, stream in+out", abi.graph_clip_os(abi.POLICY_LATCH, 4), 1048576, 1, 4, 2),
]


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    only = sys.argv[2] if len(sys.argv) > 2 else None
    for name, kind, n, vpi, in_bytes, nstate in CASES:
        if only and only not in name:
            continue
        is_clip = kind >= abi.GRAPH_CLIP_REF
        keep = abi.KEEP_INSTANCE_OUT if vpi > 1 else (abi.KEEP_VOICE_OUT if is_clip else 0)
        e = abi.Engine(kind, n, vpi, 48000.0, 0, abi.PLAN_FUSED, keep)
        rng = np.random.default_rng(1)
        if kind == abi.GRAPH_TWIN_PEAKS or is_clip:
            if in_bytes:
                x = rng.uniform(-1, 1, size=(FRAMES, n)).astype(np.float32)
                e.stream_in(0, x)
            else:
                e.stream_in_lcg(0, 1000, 0)
        elif kind in (abi.GRAPH_SAT_1X, abi.GRAPH_SAT_4X):
            e.set_value(0, (2000.0 * (0.5 + rng.random(n))).astype(np.float32))
        else:
            notes = 36 + rng.integers(0, 60, size=n)
            e.set_value(0 if kind != abi.GRAPH_SYNTH_ADSR else 2, (440.0 * np.exp2((notes - 69) / 12.0)).astype(np.float32))
            ev = np.zeros(n, dtype=abi.EVENT_DTYPE)
            ev["voice"] = np.arange(n); ev["frame_offset"] = rng.integers(0, FRAMES, size=n)
            ev["input_id"] = abi.EVENT_GATE; ev["scalar"] = 0.8
            e.push_events(ev)
        e.enable_timing(True)
        for _ in range(3):
            e.l2_flush(); e.process_block(FRAMES)
        e.synchronize()
        ms = []
        for _ in range(steps):
            e.l2_flush(); e.process_block(FRAMES); ms.append(e.last_kernel_ms())
        k_ms = float(np.mean(ms))
        ch = e.channels
        bytes_f = n * (2 * nstate * 4 + 4 * abi.load().oscen_b200_n_inputs(kind)) + n * FRAMES * in_bytes \
            + ((n // vpi) * FRAMES * 4 * ch if keep & abi.KEEP_INSTANCE_OUT else 0) + (n * FRAMES * 4 if keep & abi.KEEP_VOICE_OUT else 0) \
            + ((n + 63) // 64) * FRAMES * 4 * ch
        print(json.dumps({"graph": name, "voices": n, "frames": FRAMES, "kernel_ms": k_ms,
                          "voice_samples_per_s": n * FRAMES / (k_ms * 1e-3),
                          "realtime_voices_48k": n * FRAMES / (k_ms * 1e-3) / 48000.0,
                          "model_f_bytes_per_launch": bytes_f, "hbm_gbs": bytes_f / (k_ms * 1e-3) / 1e9,
                          "hbm_frac_of_measured_peak": bytes_f / (k_ms * 1e-3) / 1e9 / PEAK}), flush=True)
        e.close()


if __name__ == "__main__":
    main()
