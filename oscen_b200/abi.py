"""ctypes binding to the engine's C ABI (include/oscen_b200.h) — the host-side Python mirror used
by tests and bench.py.  It never falls back to a CPU implementation: if the shared library or a
CUDA device is missing, it raises."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.environ.get("OSCEN_B200_LIB") or os.path.join(HERE, "lib", "liboscen_b200.so")   # env override: A/B of two builds

ABI_VERSION = 1
MAX_BLOCK = 512
EVENT_GATE = 0xFFFFFFFF
(GRAPH_FM_VOICE, GRAPH_TWIN_PEAKS, GRAPH_SAT_1X, GRAPH_SAT_4X, GRAPH_EPIANO, GRAPH_SYNTH,
 GRAPH_SYNTH_ADSR) = range(7)
GRAPH_CLIP_REF = 7
POLICY_SINC, POLICY_SINC_IIR, POLICY_LINEAR, POLICY_LATCH = range(4)
GRAPH_CUSTOM = 20
GRAPH_COUNT = 21


def graph_clip_os(policy, factor):
    """OSCEN_B200_GRAPH_CLIP_OS(policy, log2 factor): ClipOversampled with `policy` on both cross-rate edges"""
    return 8 + policy * 3 + ({2: 1, 4: 2, 8: 3}[factor] - 1)


PLAN_FUSED, PLAN_NODES = 0, 1
KEEP_VOICE_OUT, KEEP_INSTANCE_OUT = 1, 2

# every symbol include/oscen_b200.h declares
SYMBOLS = [
    "oscen_b200_abi_version", "oscen_b200_device_count", "oscen_b200_create", "oscen_b200_destroy",
    "oscen_b200_last_error", "oscen_b200_set_value", "oscen_b200_set_value_ramp", "oscen_b200_push_events",
    "oscen_b200_push_midi", "oscen_b200_push_midi_batch", "oscen_b200_stream_in", "oscen_b200_stream_in_lcg", "oscen_b200_process_block",
    "oscen_b200_synchronize", "oscen_b200_read_voice_out", "oscen_b200_read_instance_out", "oscen_b200_read_mix",
    "oscen_b200_channels", "oscen_b200_latency_samples", "oscen_b200_n_inputs",
    "oscen_b200_state_bytes_per_voice", "oscen_b200_state_download", "oscen_b200_get_stats",
    "oscen_b200_last_block_ms", "oscen_b200_enable_timing", "oscen_b200_last_kernel_ms", "oscen_b200_timing_drain", "oscen_b200_last_kernel_name", "oscen_b200_l2_flush", "oscen_b200_comm_unique_id", "oscen_b200_comm_init", "oscen_b200_comm_p2p_export", "oscen_b200_comm_p2p_import", "oscen_b200_comm_p2p_close", "oscen_b200_plan_describe",
    "oscen_b200_graph_compile", "oscen_b200_graph_compile_named", "oscen_b200_graph_free", "oscen_b200_graph_describe", "oscen_b200_graph_n_inputs",
    "oscen_b200_graph_input_id", "oscen_b200_graph_input_info", "oscen_b200_create_from_graph", "oscen_b200_specialize",
    "oscen_b200_graph_codegen", "oscen_b200_graph_jit_check", "oscen_b200_graph_voices_per_instance", "oscen_b200_graph_channels", "oscen_b200_graph_has_midi",
    "oscen_b200_kind_info", "oscen_b200_graph_kernel_info", "oscen_b200_graph_kernel_source", "oscen_b200_graph_n_assets", "oscen_b200_graph_asset_id", "oscen_b200_load_asset",
]


class Config(C.Structure):
    _fields_ = [("graph_kind", C.c_uint32), ("n_voices", C.c_uint32), ("voices_per_instance", C.c_uint32),
                ("sample_rate", C.c_float), ("device", C.c_int32), ("plan", C.c_uint32), ("keep", C.c_uint32),
                ("reserved", C.c_uint32)]


class KindTables(C.Structure):
    _fields_ = [("n_inputs", C.c_uint32), ("n_state", C.c_uint32), ("channels", C.c_uint32), ("dyn_smem_bytes", C.c_uint32),
                ("has_stream_in", C.c_uint32), ("reserved", C.c_uint32)]


def kind_info(graph_kind, sample_rate=48000.0):
    """(KindTables, defaults [n_inputs] f32, initial state words of one voice [n_state] u32) of a built-in graph kind"""
    L = load()
    t = KindTables()
    d = (C.c_float * 28)()
    if L.oscen_b200_kind_info(graph_kind, sample_rate, C.byref(t), d, None, 0) != 0:
        raise EngineError(f"kind_info: bad kind {graph_kind}")
    w = (C.c_uint32 * max(t.n_state, 1))()
    L.oscen_b200_kind_info(graph_kind, sample_rate, None, None, w, t.n_state)
    return t, np.array(d[:t.n_inputs], dtype=np.float32), np.frombuffer(w, dtype=np.uint32)[:t.n_state].copy()


class KernelInfo(C.Structure):
    _fields_ = [("source_len", C.c_uint32), ("n_state", C.c_uint32), ("dyn_smem_bytes", C.c_uint32), ("inner_factor", C.c_uint32)]


class Event(C.Structure):
    _fields_ = [("voice", C.c_uint32), ("frame_offset", C.c_uint32), ("input_id", C.c_uint32),
                ("scalar", C.c_float)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("kernel_launches", "graph_launches", "h2d_bytes", "d2h_bytes",
                                          "blocks", "events_dropped")]


class InputInfo(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("default_value", C.c_float), ("has_range", C.c_uint32), ("range_min", C.c_float),
                ("range_max", C.c_float), ("logarithmic", C.c_uint32), ("ramp_frames", C.c_uint32)]


EVENT_DTYPE = np.dtype([("voice", "<u4"), ("frame_offset", "<u4"), ("input_id", "<u4"), ("scalar", "<f4")])
MIDI_DTYPE = np.dtype([("instance", "<u4"), ("frame_offset", "<u4"), ("bytes", "u1", (3,)), ("len", "u1")])   # oscen_b200_midi_msg

_lib = None


def load():
    """dlopen the in-tree library (built by oscen_b200.build); raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        raise RuntimeError(f"{LIB} is missing: run `python -m oscen_b200.build` (there is no fallback)")
    L = C.CDLL(LIB)
    vp, u32, f32p = C.c_void_p, C.c_uint32, C.c_void_p
    L.oscen_b200_abi_version.restype = C.c_int
    L.oscen_b200_device_count.restype = C.c_int
    L.oscen_b200_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.oscen_b200_destroy.argtypes = [vp]
    L.oscen_b200_destroy.restype = None
    L.oscen_b200_last_error.argtypes = [vp]
    L.oscen_b200_last_error.restype = C.c_char_p
    L.oscen_b200_set_value.argtypes = [vp, u32, u32, u32, f32p, C.c_int]
    L.oscen_b200_set_value_ramp.argtypes = [vp, u32, C.c_float, u32]
    L.oscen_b200_push_events.argtypes = [vp, vp, C.c_size_t]
    L.oscen_b200_push_midi.argtypes = [vp, u32, u32, vp, u32]
    L.oscen_b200_push_midi_batch.argtypes = [vp, vp, C.c_size_t]
    L.oscen_b200_stream_in.argtypes = [vp, u32, vp, u32, C.c_int]
    L.oscen_b200_stream_in_lcg.argtypes = [vp, u32, C.c_uint64, u32]
    L.oscen_b200_process_block.argtypes = [vp, u32]
    L.oscen_b200_synchronize.argtypes = [vp]
    L.oscen_b200_read_voice_out.argtypes = [vp, f32p, u32, u32]
    L.oscen_b200_read_instance_out.argtypes = [vp, f32p, u32, u32]
    L.oscen_b200_read_mix.argtypes = [vp, f32p]
    L.oscen_b200_channels.argtypes = [vp]
    L.oscen_b200_channels.restype = u32
    L.oscen_b200_latency_samples.argtypes = [vp]
    L.oscen_b200_latency_samples.restype = u32
    L.oscen_b200_n_inputs.argtypes = [u32]
    L.oscen_b200_n_inputs.restype = u32
    L.oscen_b200_state_bytes_per_voice.argtypes = [vp]
    L.oscen_b200_state_bytes_per_voice.restype = C.c_size_t
    L.oscen_b200_state_download.argtypes = [vp, vp, C.c_size_t]
    L.oscen_b200_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.oscen_b200_last_block_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.oscen_b200_enable_timing.argtypes = [vp, C.c_int]
    L.oscen_b200_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.oscen_b200_l2_flush.argtypes = [vp]
    L.oscen_b200_timing_drain.argtypes = [vp, vp, vp, u32]
    L.oscen_b200_last_kernel_name.argtypes = [vp]
    L.oscen_b200_last_kernel_name.restype = C.c_char_p
    L.oscen_b200_comm_unique_id.argtypes = [vp]
    L.oscen_b200_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.oscen_b200_comm_p2p_export.argtypes = [vp, C.c_int, vp]
    L.oscen_b200_plan_describe.argtypes = [u32, C.c_char_p, C.c_size_t]
    L.oscen_b200_comm_p2p_import.argtypes = [vp, C.c_int, C.c_int, vp]
    L.oscen_b200_comm_p2p_close.argtypes = [vp]
    L.oscen_b200_graph_compile.argtypes = [C.c_char_p, C.POINTER(vp), C.c_char_p, C.c_size_t]
    L.oscen_b200_graph_compile_named.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp), C.c_char_p, C.c_size_t]
    L.oscen_b200_graph_free.argtypes = [vp]
    L.oscen_b200_graph_free.restype = None
    L.oscen_b200_graph_describe.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.oscen_b200_graph_n_inputs.argtypes = [vp]
    L.oscen_b200_graph_n_inputs.restype = u32
    L.oscen_b200_graph_input_id.argtypes = [vp, C.c_char_p]
    L.oscen_b200_graph_input_info.argtypes = [vp, u32, C.POINTER(InputInfo)]
    L.oscen_b200_create_from_graph.argtypes = [vp, C.POINTER(Config), C.POINTER(vp)]
    L.oscen_b200_specialize.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.oscen_b200_graph_voices_per_instance.argtypes = [vp]
    L.oscen_b200_graph_voices_per_instance.restype = u32
    L.oscen_b200_graph_channels.argtypes = [vp]
    L.oscen_b200_graph_channels.restype = u32
    L.oscen_b200_graph_has_midi.argtypes = [vp]
    L.oscen_b200_graph_codegen.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.oscen_b200_graph_jit_check.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.oscen_b200_kind_info.argtypes = [C.c_uint32, C.c_float, C.POINTER(KindTables), C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.c_size_t]
    L.oscen_b200_graph_n_assets.argtypes = [vp]
    L.oscen_b200_graph_n_assets.restype = C.c_uint32
    L.oscen_b200_graph_asset_id.argtypes = [vp, C.c_char_p]
    L.oscen_b200_load_asset.argtypes = [vp, C.c_uint32, vp, C.c_uint32, C.c_uint32]
    L.oscen_b200_graph_kernel_source.argtypes = [vp, C.c_float, C.c_uint32, C.c_char_p, C.c_size_t]
    L.oscen_b200_graph_kernel_info.argtypes = [vp, C.c_float, C.POINTER(KernelInfo), C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32), C.c_size_t]
    _lib = L
    return L


class EngineError(RuntimeError):
    pass


class GraphCompileError(ValueError):
    """`line:col: message` from the DSL front end"""


class Graph:
    """A `graph!` description compiled to a node plan (oscen_b200_graph_compile).  Needs no GPU."""

    def __init__(self, dsl_text, name=None):
        """compiles the LAST graph of the text, or the one called `name`"""
        self.L = load()
        h = C.c_void_p()
        err = C.create_string_buffer(1024)
        rc = self.L.oscen_b200_graph_compile_named(dsl_text.encode(), name.encode() if name else None, C.byref(h), err, len(err))
        if rc != 0:
            raise GraphCompileError(err.value.decode())
        self.h = h
        self.n_inputs = self.L.oscen_b200_graph_n_inputs(h)
        self.voices_per_instance = self.L.oscen_b200_graph_voices_per_instance(h)    # N of `voices = [Voice::new(); N]`, else 1
        self.channels = self.L.oscen_b200_graph_channels(h)                          # 2 for a `stream: Frame<2>` output
        self.has_midi = bool(self.L.oscen_b200_graph_has_midi(h))

    def close(self):
        if getattr(self, "h", None):
            self.L.oscen_b200_graph_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def describe(self):
        buf = C.create_string_buffer(1 << 16)
        n = self.L.oscen_b200_graph_describe(self.h, buf, len(buf))
        if n < 0:
            raise EngineError(f"graph_describe rc={n}")
        return buf.value.decode()

    def launches(self):
        """the launch plan lines of describe(): ["0: AdsrEnvelope [env3, env2]", ...]"""
        return [l for l in self.describe().splitlines() if l[:1].isdigit()]

    def codegen(self):
        """CUDA source of the fused kernel generated for this graph (PLAN_FUSED)"""
        buf = C.create_string_buffer(1 << 18)
        n = self.L.oscen_b200_graph_codegen(self.h, buf, len(buf))
        if n < 0:
            raise EngineError(f"graph_codegen rc={n}")
        return buf.value.decode()

    def kernel_info(self, sample_rate=48000.0):
        """(generated source, initial state words of one voice [n_state] u32, KernelInfo) of an engine at this sample rate"""
        info = KernelInfo()
        self.L.oscen_b200_graph_kernel_info(self.h, sample_rate, C.byref(info), None, 0, None, 0)
        src = C.create_string_buffer(info.source_len + 1)
        words = (C.c_uint32 * info.n_state)()
        rc = self.L.oscen_b200_graph_kernel_info(self.h, sample_rate, C.byref(info), src, len(src), words, info.n_state)
        if rc != 0:
            raise EngineError(f"graph_kernel_info rc={rc}")
        return src.value.decode(), np.frombuffer(words, dtype=np.uint32).copy(), info

    def kernel_source(self, sample_rate=48000.0, uniform_mask=0):
        """the generated source of the variant Engine.specialize() compiles for `uniform_mask` (0 = the general kernel)"""
        n = self.L.oscen_b200_graph_kernel_source(self.h, sample_rate, uniform_mask, None, 0)
        if n < 0:
            raise EngineError(f"graph_kernel_source rc={n}")
        buf = C.create_string_buffer(n + 1)
        self.L.oscen_b200_graph_kernel_source(self.h, sample_rate, uniform_mask, buf, len(buf))
        return buf.value.decode()

    def jit_check(self):
        """compile the generated kernel with NVRTC for sm_100a (no GPU needed); returns the cubin size"""
        log = C.create_string_buffer(1 << 16)
        n = self.L.oscen_b200_graph_jit_check(self.h, log, len(log))
        if n < 0:
            raise EngineError(f"graph_jit_check rc={n}: {log.value.decode()}")
        return n

    def input_id(self, name):
        i = self.L.oscen_b200_graph_input_id(self.h, name.encode())
        if i < 0:
            raise KeyError(name)
        return i

    def asset_id(self, name):
        """index of the `external NAME: AudioAsset;` slot"""
        i = self.L.oscen_b200_graph_asset_id(self.h, name.encode())
        if i < 0:
            raise KeyError(name)
        return i

    def n_assets(self):
        return self.L.oscen_b200_graph_n_assets(self.h)

    def input_info(self, input_id):
        info = InputInfo()
        if self.L.oscen_b200_graph_input_info(self.h, input_id, C.byref(info)) != 0:
            raise KeyError(input_id)
        return {"name": info.name.decode(), "default": info.default_value,
                "range": (info.range_min, info.range_max) if info.has_range else None,
                "log": bool(info.logarithmic), "ramp": info.ramp_frames}


class Engine:
    """One engine = one GPU = one shard of voices.  Method names follow the C ABI."""

    def __init__(self, graph_kind, n_voices, voices_per_instance=1, sample_rate=48000.0, device=0,
                 plan=PLAN_FUSED, keep=KEEP_VOICE_OUT | KEEP_INSTANCE_OUT, graph=None):
        self.L = load()
        h = C.c_void_p()
        if graph is not None:          # a DSL-compiled graph: node plan, or a generated fused kernel (NVRTC)
            graph_kind = GRAPH_CUSTOM
            cfg = Config(graph_kind, n_voices, voices_per_instance, sample_rate, device, plan, keep, 0)
            rc = self.L.oscen_b200_create_from_graph(graph.h, C.byref(cfg), C.byref(h))
        else:
            cfg = Config(graph_kind, n_voices, voices_per_instance, sample_rate, device, plan, keep, 0)
            rc = self.L.oscen_b200_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise EngineError(f"oscen_b200_create rc={rc}: {self.L.oscen_b200_last_error(None).decode()}")
        self.h = h
        self.graph = graph             # DSL-compiled graphs: input names, declared ramps
        self.kind, self.n_voices, self.vpi = graph_kind, n_voices, voices_per_instance
        self.n_instances = n_voices // voices_per_instance
        self.channels = self.L.oscen_b200_channels(h)
        self.frames = 0

    @classmethod
    def from_graph(cls, graph, n_voices, voices_per_instance=None, sample_rate=48000.0, device=0,
                   keep=KEEP_VOICE_OUT | KEEP_INSTANCE_OUT, plan=PLAN_NODES):
        """G::new() + G::init(sample_rate) for n_voices voices of a DSL-compiled graph (text or Graph);
        plan = PLAN_NODES (node kernels, CUDA graph) or PLAN_FUSED (one generated kernel, NVRTC).
        voices_per_instance defaults to the graph's own polyphony (1 unless it holds a voice array)."""
        if isinstance(graph, str):
            graph = Graph(graph)
        if voices_per_instance is None:
            voices_per_instance = graph.voices_per_instance
        return cls(GRAPH_CUSTOM, n_voices, voices_per_instance, sample_rate, device, plan, keep, graph=graph)

    def specialize(self):
        """compile the generated kernel's variant for the value inputs that are uniform over the voices now (PLAN_FUSED engines
        of a DSL graph); returns the mask of input ids taken"""
        m = C.c_uint32(0)
        self._chk(self.L.oscen_b200_specialize(self.h, C.byref(m)), "specialize")
        return int(m.value)

    def close(self):
        if getattr(self, "h", None):
            self.L.oscen_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, what):
        if rc != 0:
            raise EngineError(f"{what} rc={rc}: {self.L.oscen_b200_last_error(self.h).decode()}")

    def set_value(self, input_id, values, first_voice=0, count=None):
        a = np.ascontiguousarray(np.atleast_1d(np.asarray(values, dtype=np.float32)))
        if count is None:
            count = self.n_voices - first_voice if a.size == 1 else a.size
        self._chk(self.L.oscen_b200_set_value(self.h, input_id, first_voice, count, a.ctypes.data,
                                              1 if a.size == 1 else 0), "set_value")

    def set(self, name, value):
        """the generated `set_<name>(value)` of a DSL graph (codegen/mod.rs:917-976): ramps over the input's declared
        `ramp: N` frames, or assigns when the input has no ramp"""
        if self.graph is None:
            raise EngineError("set(name, value) needs an engine created from a DSL graph")
        i = self.graph.input_id(name)
        ramp = self.graph.input_info(i)["ramp"]
        if ramp:
            self.set_value_ramp(i, value, ramp)
        else:
            self.set_value(i, value)

    def render(self, frames, stream_in=None):
        """BlockRender::render (oscen-lib/src/graph/offline.rs:46-90): any number of frames in chunks of at most
        MAX_BLOCK; returns the per-voice output [frames][voices] (needs KEEP_VOICE_OUT or the node plan)"""
        out = np.empty((frames, self.n_voices), dtype=np.float32)
        pos = 0
        while pos < frames:
            chunk = min(MAX_BLOCK, frames - pos)
            if stream_in is not None:
                self.stream_in(0, stream_in[pos:pos + chunk])
            self.process_block(chunk)
            out[pos:pos + chunk] = self.read_voice_out()
            pos += chunk
        return out

    def set_value_ramp(self, input_id, target, frames):
        self._chk(self.L.oscen_b200_set_value_ramp(self.h, input_id, float(target), int(frames)), "set_value_ramp")

    def push_events(self, events):
        """events: structured ndarray (EVENT_DTYPE) or iterable of (voice, frame_offset, input_id, scalar)"""
        if isinstance(events, np.ndarray) and events.dtype == EVENT_DTYPE:
            arr = np.ascontiguousarray(events)
        else:
            ev = list(events)
            if not ev:
                return
            arr = np.array([(int(v), int(f), int(i), float(s)) for v, f, i, s in ev], dtype=EVENT_DTYPE)
        if arr.size:
            self._chk(self.L.oscen_b200_push_events(self.h, arr.ctypes.data, arr.size), "push_events")

    def push_midi(self, instance, frame_offset, data):
        b = (C.c_uint8 * len(data))(*data)
        self._chk(self.L.oscen_b200_push_midi(self.h, instance, frame_offset, b, len(data)), "push_midi")

    def push_midi_batch(self, msgs):
        """device-side MIDI front: structured ndarray (MIDI_DTYPE) or iterable of (instance, frame_offset, bytes)"""
        if not (isinstance(msgs, np.ndarray) and msgs.dtype == MIDI_DTYPE):
            rows = list(msgs)
            arr = np.zeros(len(rows), dtype=MIDI_DTYPE)
            for k, (inst, frame, data) in enumerate(rows):
                data = list(data)[:3]
                arr[k]["instance"], arr[k]["frame_offset"], arr[k]["len"] = inst, frame, len(data)
                arr[k]["bytes"][:len(data)] = data
            msgs = arr
        msgs = np.ascontiguousarray(msgs)
        if msgs.size:
            self._chk(self.L.oscen_b200_push_midi_batch(self.h, msgs.ctypes.data, msgs.size), "push_midi_batch")

    def load_asset(self, slot, samples):
        """graph.<slot>.load(..): samples [frames] (mono) or [frames][channels] f32 at the engine's sample rate; every
        SamplePlayer bound to the slot restarts at the next block.  `slot`: name (DSL engines) or index."""
        if isinstance(slot, str):
            if self.graph is None:
                raise EngineError("load_asset by name needs an engine created from a DSL graph")
            slot = self.graph.asset_id(slot)
        a = np.ascontiguousarray(samples, dtype=np.float32)
        if a.ndim == 1:
            a = a[:, None]
        assert a.ndim == 2
        self._chk(self.L.oscen_b200_load_asset(self.h, slot, a.ctypes.data if a.size else None, a.shape[0], max(a.shape[1], 1)), "load_asset")

    def stream_in(self, input_id, arr):
        a = np.ascontiguousarray(arr, dtype=np.float32)
        assert a.ndim == 2 and a.shape[1] == self.n_voices
        self._chk(self.L.oscen_b200_stream_in(self.h, input_id, a.ctypes.data, a.shape[0], 0), "stream_in")

    def stream_in_lcg(self, input_id, seed, voice_offset=0):
        self._chk(self.L.oscen_b200_stream_in_lcg(self.h, input_id, seed, voice_offset), "stream_in_lcg")

    def process_block(self, frames):
        self._chk(self.L.oscen_b200_process_block(self.h, frames), "process_block")
        self.frames = frames

    def synchronize(self):
        self._chk(self.L.oscen_b200_synchronize(self.h), "synchronize")

    def read_voice_out(self, first_voice=0, count=None):
        count = self.n_voices - first_voice if count is None else count
        out = np.empty((self.frames, count), dtype=np.float32)
        self._chk(self.L.oscen_b200_read_voice_out(self.h, out.ctypes.data, first_voice, count), "read_voice_out")
        return out

    def read_instance_out(self, first=0, count=None):
        count = self.n_instances - first if count is None else count
        out = np.empty((self.frames, count, self.channels), dtype=np.float32)
        self._chk(self.L.oscen_b200_read_instance_out(self.h, out.ctypes.data, first, count), "read_instance_out")
        return out

    def read_mix(self, out=None):
        if out is None:
            out = np.empty((self.frames, self.channels), dtype=np.float32)
        self._chk(self.L.oscen_b200_read_mix(self.h, out.ctypes.data), "read_mix")
        return out

    def latency_samples(self):
        return self.L.oscen_b200_latency_samples(self.h)

    def state_download(self):
        nb = self.L.oscen_b200_state_bytes_per_voice(self.h)
        out = np.empty((nb // 4, self.n_voices), dtype=np.float32)
        self._chk(self.L.oscen_b200_state_download(self.h, out.ctypes.data, out.nbytes), "state_download")
        return out

    def stats(self):
        s = Stats()
        self._chk(self.L.oscen_b200_get_stats(self.h, C.byref(s)), "get_stats")
        return {n: getattr(s, n) for n, _ in Stats._fields_}

    def enable_timing(self, on=True):
        self._chk(self.L.oscen_b200_enable_timing(self.h, 1 if on else 0), "enable_timing")

    def last_block_ms(self):
        ms = C.c_float()
        self._chk(self.L.oscen_b200_last_block_ms(self.h, C.byref(ms)), "last_block_ms")
        return ms.value

    def last_kernel_ms(self):
        ms = C.c_float()
        self._chk(self.L.oscen_b200_last_kernel_ms(self.h, C.byref(ms)), "last_kernel_ms")
        return ms.value

    def timing_drain(self, cap=4096):
        """(block_ms, kernel_ms) float32 arrays of every block since enable_timing(True) / the previous drain; one sync"""
        b = np.empty(cap, dtype=np.float32); k = np.empty(cap, dtype=np.float32)
        n = self.L.oscen_b200_timing_drain(self.h, b.ctypes.data, k.ctypes.data, cap)
        if n < 0:
            self._chk(n, "timing_drain")
        return b[:n].copy(), k[:n].copy()

    def last_kernel_name(self):
        return (self.L.oscen_b200_last_kernel_name(self.h) or b"").decode()

    def l2_flush(self):
        self._chk(self.L.oscen_b200_l2_flush(self.h), "l2_flush")

    def comm_init(self, rank, n_ranks, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._chk(self.L.oscen_b200_comm_init(self.h, rank, n_ranks, buf), "comm_init")


    def comm_p2p_export(self, n_ranks) -> bytes:
        buf = C.create_string_buffer(64)
        self._chk(self.L.oscen_b200_comm_p2p_export(self.h, n_ranks, buf), "comm_p2p_export")
        return buf.raw

    def comm_p2p_import(self, rank, n_ranks, handles: bytes):
        assert len(handles) == 64 * n_ranks
        buf = C.create_string_buffer(handles, len(handles))
        self._chk(self.L.oscen_b200_comm_p2p_import(self.h, rank, n_ranks, buf), "comm_p2p_import")

    def comm_p2p_close(self):
        self._chk(self.L.oscen_b200_comm_p2p_close(self.h), "comm_p2p_close")


def plan_describe(graph_kind):
    """the node plan as a list of (node_type, [node names]) launches; raises for fused-only graphs"""
    L = load()
    buf = C.create_string_buffer(4096)
    rc = L.oscen_b200_plan_describe(graph_kind, buf, 4096)
    if rc <= 0:
        raise EngineError(f"plan_describe rc={rc}")
    out = []
    for line in buf.value.decode().strip().splitlines():
        head, names = line.split(" [")
        out.append((head.split(": ", 1)[1], [x.strip() for x in names.rstrip("]").split(",")]))
    return out


def comm_unique_id() -> bytes:
    L = load()
    buf = C.create_string_buffer(128)
    rc = L.oscen_b200_comm_unique_id(buf)
    if rc != 0:
        raise EngineError(f"comm_unique_id rc={rc}: {L.oscen_b200_last_error(None).decode()}")
    return buf.raw
