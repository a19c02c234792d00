"""ctypes binding to the CPU oracle (oracle/liboscen_oracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never by oscen_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ORACLE_DIR, "liboscen_oracle.so")

EV_GATE = 0xFFFFFFFF
(K_FM, K_TWIN, K_SAT1, K_SAT4, K_EPIANO, K_SYNTH, K_SYNTH_ADSR) = range(7)
K_CLIP_REF = 7
(POLICY_SINC, POLICY_SINC_IIR, POLICY_LINEAR, POLICY_LATCH) = range(4)


def k_clip_os(policy, factor):
    """graph kind of the ClipOversampled family (oscen-lib/tests/multirate_graph.rs:295-325)"""
    return 8 + policy * 3 + ({2: 1, 4: 2, 8: 3}[factor] - 1)



def build(force=False):
    src = [os.path.join(ORACLE_DIR, f) for f in ("oscen_oracle.c", "oscen_oracle.h", "Makefile")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return SO


class Ramp(C.Structure):
    _fields_ = [("current", C.c_float), ("target", C.c_float), ("increment", C.c_float),
                ("frames_remaining", C.c_uint32)]


class Adsr(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("attack", "decay", "sustain", "release", "output")] + \
        [("stage", C.c_int)] + \
        [(n, C.c_uint32) for n in ("attack_samples", "decay_samples", "release_samples", "samples_remaining")] + \
        [(n, C.c_float) for n in ("attack_coeff", "decay_coeff", "release_increment", "level",
                                  "target_level", "sustain_level", "velocity", "sample_rate")]


class Tpt(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("input", "cutoff", "q", "f_mod", "output", "current_cutoff",
                                         "current_q")] + [("z", C.c_float * 2)] + \
        [(n, C.c_float) for n in ("h", "g", "r", "k", "sample_rate")]


class PolyBlep(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("phase", "phase_mod", "frequency", "frequency_mod", "amplitude",
                                         "pulse_width", "output")] + [("waveform", C.c_int),
                                                                      ("sample_rate", C.c_float)]


class FmOp(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("phase", "prev_output", "sample_rate", "base_freq", "ratio",
                                         "phase_mod", "feedback", "envelope", "level", "output")]


class Lp18(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("input", "cutoff", "fmod", "resonance")] + [("z", C.c_float * 3)] + \
        [(n, C.c_float) for n in ("g", "h", "last_cutoff", "last_fmod", "last_resonance", "sample_rate",
                                  "output")]


class IirLp(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("input", "cutoff", "q", "output", "b0", "b1", "b2", "a1", "a2", "v1", "v2",
                                         "sample_rate")] + [("frame_counter", C.c_size_t), ("frames_per_update", C.c_size_t)]


class Ring(C.Structure):
    _fields_ = [("buffer", C.POINTER(C.c_float)), ("write_pos", C.c_size_t), ("capacity", C.c_size_t), ("mask", C.c_size_t)]


class Delay(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("input", "delay_samples", "feedback", "output")] + [("ring", Ring), ("sample_rate", C.c_float),
                                                                                         ("frames_per_update", C.c_size_t), ("frame_counter", C.c_size_t)]


class SamplePlayer(C.Structure):
    _fields_ = [("output", C.c_float), ("current", C.c_void_p), ("current_len", C.c_size_t), ("pending", C.c_void_p),
                ("pending_len", C.c_size_t), ("has_pending", C.c_int), ("playhead", C.c_size_t)]


class NaiveOsc(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("phase", "frequency", "frequency_mod", "amplitude", "output")] + [("waveform", C.c_int), ("sample_rate", C.c_float)]


class HbDown(C.Structure):
    _fields_ = [("history", C.c_float * 24), ("head", C.c_size_t)]


class HbUp(C.Structure):
    _fields_ = [("history", C.c_float * 12), ("head", C.c_size_t)]


class SincDown(C.Structure):
    _fields_ = [("stages", HbDown * 3), ("n_stages", C.c_int), ("n", C.c_int)]


class SincUp(C.Structure):
    _fields_ = [("stages", HbUp * 3), ("n_stages", C.c_int), ("n", C.c_int)]


class Allpass1(C.Structure):
    _fields_ = [("a", C.c_float), ("x_prev", C.c_float), ("y_prev", C.c_float)]


class IirHb(C.Structure):
    _fields_ = [("branch_a", Allpass1 * 2), ("branch_b", Allpass1 * 2), ("prev_odd_in", C.c_float)]


class IirResampler(C.Structure):
    _fields_ = [("stages", IirHb * 3), ("n_stages", C.c_int), ("n", C.c_int)]


class LinearUp(C.Structure):
    _fields_ = [("prev", C.c_float), ("n", C.c_int)]


class VoiceAlloc(C.Structure):
    _fields_ = [("num_voices", C.c_int), ("active", C.c_uint8 * 24), ("released", C.c_uint8 * 24),
                ("note", C.c_int * 24), ("age", C.c_uint32 * 24), ("current_age", C.c_uint32)]


class Event(C.Structure):
    _fields_ = [("voice", C.c_uint32), ("frame_offset", C.c_uint32), ("input_id", C.c_uint32),
                ("scalar", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(SO)
        L.orc_midi_note_to_freq.restype = C.c_float
        L.orc_midi_note_to_freq.argtypes = [C.c_uint8]
        L.orc_poly_blep.restype = C.c_float
        L.orc_poly_blep.argtypes = [C.c_float, C.c_float]
        L.orc_poly_blamp.restype = C.c_float
        L.orc_poly_blamp.argtypes = [C.c_float, C.c_float]
        L.orc_sinc_down_process.restype = C.c_float
        L.orc_adsr_new.argtypes = [C.c_void_p] + [C.c_float] * 4
        L.orc_adsr_set_sample_rate.argtypes = [C.c_void_p, C.c_float]
        L.orc_adsr_on_gate.argtypes = [C.c_void_p, C.c_float]
        L.orc_tpt_new.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.orc_tpt_set_sample_rate.argtypes = [C.c_void_p, C.c_float]
        L.orc_polyblep_new.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int]
        L.orc_polyblep_set_sample_rate.argtypes = [C.c_void_p, C.c_float]
        L.orc_lp18_new.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.orc_iirlp_new.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.orc_delay_new.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.orc_osc_new.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int]
        L.orc_sample_player_publish.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_sample_player_build_mono.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
        L.orc_ring_init.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_ring_push.argtypes = [C.c_void_p, C.c_float]
        L.orc_ring_get.argtypes = [C.c_void_p, C.c_float]
        L.orc_ring_get.restype = C.c_float
        L.orc_ramp_new.argtypes = [C.c_void_p, C.c_float]
        L.orc_ramp_set_immediate.argtypes = [C.c_void_p, C.c_float]
        L.orc_ramp_set_with_ramp.argtypes = [C.c_void_p, C.c_float, C.c_uint32]
        L.orc_sinc_up_process.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
        L.orc_iir_up_process.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
        L.orc_iir_down_process.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_iir_down_process.restype = C.c_float
        L.orc_linear_up_process.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
        L.orc_linear_down_process.argtypes = [C.c_int, C.c_void_p]
        L.orc_linear_down_process.restype = C.c_float
        L.orc_latch_up_process.argtypes = [C.c_int, C.c_float, C.c_void_p]
        L.orc_latch_down_process.argtypes = [C.c_int, C.c_void_p]
        L.orc_latch_down_process.restype = C.c_float
        L.orc_bank_create.restype = C.c_void_p
        L.orc_bank_create.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_float]
        L.orc_bank_destroy.argtypes = [C.c_void_p]
        L.orc_bank_set_value.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int]
        L.orc_bank_set_value_ramp.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_uint32]
        L.orc_bank_push_events.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_bank_push_midi.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
        L.orc_bank_stream_in.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.orc_bank_process_block.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_bank_read_voice_out.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_bank_read_instance_out.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_bank_read_mix.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_bank_channels.argtypes = [C.c_void_p]
        L.orc_bank_channels.restype = C.c_uint32
        L.orc_bank_latency.argtypes = [C.c_void_p]
        L.orc_bank_latency.restype = C.c_uint32
        L.orc_bank_set_threads.argtypes = [C.c_void_p, C.c_int]
        _lib = L
    return _lib


class OracleBank:
    """Same driving surface as oscen_b200.Engine, so parity tests run one script on both."""

    def __init__(self, kind, n_voices, voices_per_instance=1, sample_rate=48000.0, threads=1):
        self.L = lib()
        self.h = self.L.orc_bank_create(kind, n_voices, voices_per_instance, sample_rate)
        if not self.h:
            raise ValueError("orc_bank_create failed")
        self.kind, self.n_voices, self.vpi = kind, n_voices, voices_per_instance
        self.n_instances = n_voices // voices_per_instance
        self.channels = self.L.orc_bank_channels(self.h)
        self.frames = 0
        if threads != 1:
            self.L.orc_bank_set_threads(self.h, threads)

    def close(self):
        if self.h:
            self.L.orc_bank_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _chk(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"oracle {what} failed rc={rc}")

    def set_value(self, input_id, values, first_voice=0, count=None):
        a = np.ascontiguousarray(np.atleast_1d(np.asarray(values, dtype=np.float32)))
        if count is None:
            count = self.n_voices - first_voice if a.size == 1 else a.size
        bc = 1 if a.size == 1 else 0
        self._chk(self.L.orc_bank_set_value(self.h, input_id, first_voice, count, a.ctypes.data, bc), "set_value")

    def set_value_ramp(self, input_id, target, frames):
        self._chk(self.L.orc_bank_set_value_ramp(self.h, input_id, float(target), int(frames)), "set_value_ramp")

    def push_events(self, events):
        """events: iterable of (voice, frame_offset, input_id, scalar)"""
        if isinstance(events, np.ndarray) and events.dtype.itemsize == 16:      # same layout as orc_event
            arr = np.ascontiguousarray(events)
            if arr.size:
                self._chk(self.L.orc_bank_push_events(self.h, arr.ctypes.data, arr.size), "push_events")
            return
        ev = list(events)
        if not ev:
            return
        arr = (Event * len(ev))(*[Event(int(v), int(f), int(i), float(s)) for v, f, i, s in ev])
        self._chk(self.L.orc_bank_push_events(self.h, arr, len(ev)), "push_events")

    def push_midi(self, instance, frame_offset, data):
        b = (C.c_uint8 * len(data))(*data)
        self._chk(self.L.orc_bank_push_midi(self.h, instance, frame_offset, b, len(data)), "push_midi")

    def stream_in(self, input_id, arr):
        a = np.ascontiguousarray(arr, dtype=np.float32)
        assert a.shape[1] == self.n_voices
        self._chk(self.L.orc_bank_stream_in(self.h, input_id, a.ctypes.data, a.shape[0]), "stream_in")

    def process_block(self, frames):
        self._chk(self.L.orc_bank_process_block(self.h, frames), "process_block")
        self.frames = frames

    def read_voice_out(self, first_voice=0, count=None):
        count = self.n_voices - first_voice if count is None else count
        out = np.empty((self.frames, count), dtype=np.float32)
        self._chk(self.L.orc_bank_read_voice_out(self.h, out.ctypes.data, first_voice, count), "read_voice_out")
        return out

    def read_instance_out(self, first=0, count=None):
        count = self.n_instances - first if count is None else count
        out = np.empty((self.frames, count, self.channels), dtype=np.float32)
        self._chk(self.L.orc_bank_read_instance_out(self.h, out.ctypes.data, first, count), "read_instance_out")
        return out

    def read_mix(self):
        out = np.empty((self.frames, self.channels), dtype=np.float32)
        self._chk(self.L.orc_bank_read_mix(self.h, out.ctypes.data), "read_mix")
        return out
