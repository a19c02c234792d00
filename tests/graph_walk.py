"""Per-sample graph walks over the ORACLE's node primitives, written out by hand the way `graph!`
generates them (codegen/mod.rs:499-572: per frame, per node in topological order: assign the
edges into the node's fields, run its events, call process()).  TEST INFRASTRUCTURE: this is the
checker for graphs that the engine compiles from DSL text and for which the oracle has no bank.
It does not read the DSL — every test states its graph twice, once as text for the product and
once as a walk here, so a front-end bug cannot hide on both sides.  Pure-Python loops: small cases only."""
import ctypes as C

import numpy as np

import oracle_lib as O

f32 = np.float32
_libm = C.CDLL("libm.so.6")
_libm.tanhf.restype = C.c_float
_libm.tanhf.argtypes = [C.c_float]


for _fn in ("sinf", "cosf", "expf", "sqrtf"):
    getattr(_libm, _fn).restype = C.c_float
    getattr(_libm, _fn).argtypes = [C.c_float]


def libm32(name, x):
    """Rust f32::sin / cos / exp / sqrt -> libm sinf / cosf / expf / sqrtf"""
    return f32(getattr(_libm, name)(float(x)))


def tanh32(x):
    return f32(_libm.tanhf(float(x)))          # Rust f32::tanh -> libm tanhf


def clamp32(x, lo, hi):                         # f32::clamp
    x = f32(x)
    return f32(lo) if x < f32(lo) else (f32(hi) if x > f32(hi) else x)


def hardclip(x):                                # examples/oversampled-saturator/src/main.rs:55-62
    return clamp32(f32(x) * f32(1.5), -0.7, 0.7)


class Node:
    """one oracle node struct + its process(); fields are read / written by name like the Rust struct"""

    def __init__(self, st, process):
        self.st, self._process = st, process

    def __getattr__(self, k):
        return getattr(self.st, k)

    def set(self, **kw):
        for k, v in kw.items():
            setattr(self.st, k, float(v))
        return self

    def process(self):
        self._process(C.byref(self.st))
        return f32(self.st.output)


def polyblep(wave, freq, amp, sr):
    L = O.lib()
    st = O.PolyBlep()
    L.orc_polyblep_new(C.byref(st), freq, amp, {"sine": 0, "saw": 1, "square": 2, "triangle": 3}[wave])
    L.orc_polyblep_set_sample_rate(C.byref(st), sr)
    return Node(st, L.orc_polyblep_process)


def tpt(cutoff, q, sr):
    L = O.lib()
    st = O.Tpt()
    L.orc_tpt_new(C.byref(st), cutoff, q)
    L.orc_tpt_set_sample_rate(C.byref(st), sr)
    L.orc_tpt_prepare(C.byref(st))
    return Node(st, L.orc_tpt_process)


def adsr(a, d, s, r, sr):
    L = O.lib()
    st = O.Adsr()
    L.orc_adsr_new(C.byref(st), a, d, s, r)
    L.orc_adsr_set_sample_rate(C.byref(st), sr)
    L.orc_adsr_prepare(C.byref(st))
    n = Node(st, L.orc_adsr_process)
    n.on_gate = lambda v: L.orc_adsr_on_gate(C.byref(st), float(v))
    return n


def fmop(sr):
    L = O.lib()
    st = O.FmOp()
    L.orc_fmop_new(C.byref(st))
    st.sample_rate = sr
    return Node(st, L.orc_fmop_process)


def lp18(cutoff, res, sr):
    L = O.lib()
    st = O.Lp18()
    L.orc_lp18_new(C.byref(st), cutoff, res)
    st.sample_rate = sr
    L.orc_lp18_prepare(C.byref(st))
    return Node(st, L.orc_lp18_process)


def naive_osc(wave, freq, amp, sr):
    L = O.lib()
    st = O.NaiveOsc()
    L.orc_osc_new(C.byref(st), freq, amp, {"sine": 0, "square": 1, "saw": 2}[wave])
    st.sample_rate = sr
    return Node(st, L.orc_osc_process)


def iirlp(cutoff, q, sr):
    L = O.lib()
    st = O.IirLp()
    L.orc_iirlp_new(C.byref(st), cutoff, q)
    st.sample_rate = sr
    L.orc_iirlp_prepare(C.byref(st))
    return Node(st, L.orc_iirlp_process)


def delay(delay_samples, feedback, sr):
    L = O.lib()
    st = O.Delay()
    L.orc_delay_new(C.byref(st), delay_samples, feedback)
    st.sample_rate = sr
    L.orc_delay_prepare(C.byref(st))
    return Node(st, L.orc_delay_process)


def sample_player():
    """SamplePlayer::new() (mono); .publish(array) = the slot's load handle, picked up by the next process()"""
    L = O.lib()
    st = O.SamplePlayer()
    L.orc_sample_player_new(C.byref(st))
    n = Node(st, L.orc_sample_player_process)
    n._bufs = []

    def publish(samples):
        a = np.ascontiguousarray(samples, dtype=np.float32)
        n._bufs.append(a)                                   # the oracle borrows the memory
        L.orc_sample_player_publish(C.byref(st), a.ctypes.data, a.size)
    n.publish = publish
    return n


def render(make_voice, n_voices, frames, values=None, gates=None, stream_in=None):
    """make_voice(v) -> (frame_fn, gate_fn); frame_fn(values_of_voice: dict, x: f32) -> f32 output;
    gate_fn(scalar) handles a gate event BEFORE that frame's process (derive(Node) process_event_inputs).
    values: {name: scalar | [n_voices] array | callable(frame) -> scalar | callable(frame, voice) with .per_voice = True};
    gates: [(voice, frame, scalar)]"""
    out = np.zeros((frames, n_voices), dtype=np.float32)
    values = values or {}
    by_voice = {}
    for (v, f, s) in (gates or []):
        by_voice.setdefault(v, {}).setdefault(f, []).append(s)
    for v in range(n_voices):
        frame_fn, gate_fn = make_voice(v)
        for f in range(frames):
            vals = {}
            for k, val in values.items():
                if callable(val):
                    vals[k] = f32(val(f, v)) if getattr(val, "per_voice", False) else f32(val(f))
                elif np.ndim(val) == 0:
                    vals[k] = f32(val)
                else:
                    vals[k] = f32(val[v])
            for s in by_voice.get(v, {}).get(f, []):
                gate_fn(s)
            x = f32(stream_in[f, v]) if stream_in is not None else f32(0)
            out[f, v] = frame_fn(vals, x)
    return out
