#!/usr/bin/env python
"""bench.py — headline benchmark of the oscen_b200 engine.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--voices V] [--plan fused|nodes] [--mix p2p|nccl]

Metric (BASELINE.json): voice-samples/sec (voices x frames rendered / second) of the fm-synth voice bank, plus the
fraction of the HBM roofline.  Headline workload = BASELINE.json configs[1]: FMVoice x 65 536 voices (8 192 FMGraph
instances x 8 voices) PER GPU (weak scaling), 48 kHz, 256-frame blocks, the synthetic voice bank of SURVEY.md §8d (LCG
seeds 1-3: note, velocity, gate-on frame in block 0; gate-off in block 150; ramped parameters from block 10).  A "step" is
one process_block(256) over the whole bank.

Timing.  Pass 1 (`value`): W warm-up steps, then exactly K steps between barrier + synchronize; every step is bracketed by
CUDA events recorded on the engine's stream (read back in one batch after the loop: no per-step host synchronisation), L2
(126 MB) is flushed between steps by a 256 MiB write outside the event bracket; value = work / sum of step times, max over
ranks.  Pass 2 (`e2e`): a FRESH engine replays the SAME block schedule through the C ABI with host buffers — events and
parameter calls in, mix bus out (read_mix: D2H + synchronise) inside the host-clock bracket of every step.

Also on the line: `roofline` + `kernels` (per kernel variant that ran in the timed window), `mix_check` (N > 1: the
exchanged mix bus against all_reduce(local mix) and bit-for-bit across ranks), `other_configs` (BASELINE configs 3-5, the
node plan and the north_star operating point; at N > 1 config 4 strong-scaled and config 5), `cpu_baseline` (N = 1).

`--impl reference` times the reference's CPU path (the oracle's C restatement — the Rust crate cannot be built here, see
DESIGN.md) on all host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SR = 48000.0
FRAMES = 256
VPI = 8
VOICES_DEFAULT = 65536
MODEL_N_BYTES = 113.0        # node-granular bytes / voice-sample at B=256
FM_STATE_WORDS = 29
CTA = 64                     # kCta of the engine: voices per CTA
EV_DT = np.dtype([("voice", "<u4"), ("frame_offset", "<u4"), ("input_id", "<u4"), ("scalar", "<f4")])
EV_GATE = 0xFFFFFFFF


def lcg_stream(seed, n):
    """64-bit LCG of oscen-lib/tests/realtime_safety.rs:21-33: the first n draws (state >> 33) of the stream `seed`.
    Vectorised exactly: state_i = A^i * s0 + C * (1 + A + .. + A^(i-1)) mod 2^64 (uint64 cumprod / cumsum wrap mod 2^64)."""
    A, C = np.uint64(6364136223846793005), np.uint64(1442695040888963407)
    s0 = np.uint64((seed * 2862933555777941757 + 3037000493) & 0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        pw = np.cumprod(np.full(n, A, dtype=np.uint64))                       # pw[i] = A^(i+1)
        geo = np.cumsum(np.concatenate([np.ones(1, dtype=np.uint64), pw[:-1]]))   # geo[i] = 1 + A + .. + A^i
        state = pw * s0 + C * geo
    return state >> np.uint64(33)


_LCG_CACHE = {}


def lcg_cached(seed, n):
    key = (seed, n)
    if key not in _LCG_CACHE:
        _LCG_CACHE[key] = lcg_stream(seed, n)
    return _LCG_CACHE[key]


class Workload:
    """SURVEY.md §8d config 2, for the voices [first, first + n) of the global bank"""

    def __init__(self, n_total, first, n):
        idx = slice(first, first + n)
        if n_total <= (1 << 24):
            notes = 36 + (lcg_cached(1, n_total) % 60).astype(np.int64)
            u2, u3, u4 = lcg_cached(2, n_total), lcg_cached(3, n_total), lcg_cached(4, n_total)
        else:                # multi-million-voice banks (north_star point): a closed-form hash instead of the serial LCG walk
            g = np.arange(n_total, dtype=np.uint64)
            with np.errstate(over="ignore"):
                h = [((g + np.uint64(k)) * np.uint64(6364136223846793005) + np.uint64(1442695040888963407)) >> np.uint64(33) for k in (11, 23, 37, 41)]
            notes = 36 + (h[0] % np.uint64(60)).astype(np.int64); u2, u3, u4 = h[1], h[2], h[3]
        self.freq = (440.0 * np.exp2((notes - 69) / 12.0)).astype(np.float32)[idx]
        self.vel = (0.25 + 0.75 * (u2 / 2147483648.0)).astype(np.float32)[idx]
        self.on = (u3 % FRAMES).astype(np.uint32)[idx]
        self.off = (u4 % FRAMES).astype(np.uint32)[idx]
        self.n = n

    def events(self, block, dtype, gate):
        if block == 0:
            frame, scalar = self.on, self.vel
        elif block == 150:
            frame, scalar = self.off, np.zeros(self.n, dtype=np.float32)
        else:
            return None
        ev = np.zeros(self.n, dtype=dtype)
        ev["voice"] = np.arange(self.n); ev["frame_offset"] = frame; ev["input_id"] = gate; ev["scalar"] = scalar
        return ev

    def apply(self, bank, block, dtype, gate):
        """what the host does before process_block of `block`"""
        if block == 0:
            bank.set_value(0, self.freq)
        ev = self.events(block, dtype, gate)
        if ev is not None:
            bank.push_events(ev)
        if block == 10:      # one set_*_with_ramp(.., 2205) on the ramped inputs (fm-synth main.rs:58-87)
            for input_id, target in ((2, 0.6), (3, 0.1), (9, 0.4), (10, 0.05), (20, 0.25), (21, 2400.0), (22, 0.9), (27, 300.0)):
                bank.set_value_ramp(input_id, target, 2205)


class ClockSampler(threading.Thread):
    """SM clock and clock-event (throttle) reasons DURING the timed region (B200_PROFILING.md's clocks line).  In-process
    NVML polling (the timed region of the default run is only tens of milliseconds; an `nvidia-smi -lms` child needs
    seconds to start on an 8-GPU box and then misses it); falls back to the nvidia-smi loop when NVML cannot be loaded.
    Only rank 0 samples (its own GPU): NVML queries take a driver-wide lock (profiles/r01_diag_8gpu_nvml_sampler.log).
    begin() runs BEFORE the barrier that opens the timed region, so that thread start + nvmlInit (milliseconds) can never
    land inside a peer's first CUDA-event bracket."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device, uuid=None, period_s=0.010):
        super().__init__(daemon=True)
        self.device, self.uuid, self.period_s = device, uuid, period_s
        self.sm, self.mx, self.reasons = [], [], set()
        self.proc = None
        self.source = None
        self._stop_flag = threading.Event()
        self._ready = threading.Event()
        self._armed = threading.Event()

    def _nvml_loop(self):
        import pynvml as nv
        nv.nvmlInit()
        h = None
        if self.uuid:
            for u in (self.uuid, self.uuid.encode()):
                try:
                    h = nv.nvmlDeviceGetHandleByUUID(u); break
                except Exception:
                    h = None
        if h is None:
            h = nv.nvmlDeviceGetHandleByIndex(self.device)
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        names = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                 ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                 ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
        self.source = "nvml"
        self._ready.set()
        while not self._stop_flag.is_set():
            if self._armed.is_set():
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))); self.mx.append(mx)
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for name, bit in names:
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(self.period_s)
            else:
                self._armed.wait(timeout=0.05)

    def _smi_loop(self):
        self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-lms", "20", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
        self.source = "nvidia-smi"
        for line in self.proc.stdout:
            self._ready.set()
            if not self._armed.is_set():
                continue
            r = [x.strip() for x in line.split(",")]
            try:
                self.sm.append(float(r[1])); self.mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            except (ValueError, IndexError):
                continue

    def run(self):
        try:
            self._nvml_loop()
        except Exception:
            try:
                self._smi_loop()
            except Exception:
                pass
        self._ready.set()

    def begin(self):
        """start the thread and wait until it can sample (NVML loaded); sampling itself starts at arm()"""
        self.start()
        self._ready.wait(timeout=10.0)

    def arm(self):
        self._armed.set()

    def disarm(self):
        self._armed.clear()

    def stop(self):
        self._stop_flag.set()
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx = list(self.sm), list(self.mx)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(self.reasons),
                "samples": len(sm), "source": self.source}


class NoSampler:
    def begin(self): pass
    def arm(self): pass
    def disarm(self): pass
    def stop(self): return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "disabled (rank != 0)"}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_capture(key):
    """the whole committed capture record of `key` (profiles/ncu_traffic.json), or {}"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(key) or {}
    except Exception:
        return {}


def ncu_traffic(key):
    """DRAM bytes (read + write) of one launch from a committed `ncu --set full` capture of THIS kernel at THIS size
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py from the .ncu-rep): (bytes, source) or (None, None)"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f).get(key)
        return (int(t["dram_bytes"]), t["source"]) if t else (None, None)
    except Exception:
        return None, None


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def dist_stats(xs):
    xs = sorted(float(x) for x in xs)
    if not xs:
        return None
    q = lambda p: xs[min(len(xs) - 1, int(p * len(xs)))]
    return {"min": xs[0], "median": q(0.5), "p95": q(0.95), "max": xs[-1], "mean": sum(xs) / len(xs)}


WORKLOAD_NAME = "fm-synth FMVoice bank, 65536 voices (8192 FMGraph instances x 8) per GPU, 48 kHz, 256-frame blocks (BASELINE configs[1])"


def headline_config(args, world, mix_mode):
    n = args.voices
    return {"workload": WORKLOAD_NAME if n == VOICES_DEFAULT else f"fm-synth FMVoice bank, {n} voices per GPU, 48 kHz, 256-frame blocks",
            "voices_per_gpu": n, "voices_total": n * world, "frames": FRAMES, "sample_rate": SR, "plan": args.plan,
            "blocks_timed": [args.warmup, args.warmup + args.steps - 1],
            "parallelism": (f"voice-sharded x{world}, mix bus: " + ("fused reduce+exchange kernel over NVLink peer memory" if mix_mode == "p2p" else "ncclAllReduce")) if world > 1 else "single GPU",
            "l2": "flushed between steps (256 MiB write on the engine stream, outside the per-step CUDA-event bracket and outside the per-step e2e host clock)",
            "timing": "sum of per-step CUDA-event times on the engine stream (read back in one batch), max over ranks"}


def run_oracle(n_voices, budget_s, threads, wl_total):
    """time the oracle port on a bounded sample of the workload: the first n_voices voices"""
    import oracle_lib as O
    wl = Workload(wl_total, 0, n_voices)
    bank = O.OracleBank(O.K_FM, n_voices, VPI, SR, threads=threads)
    wl.apply(bank, 0, EV_DT, O.EV_GATE)
    bank.process_block(FRAMES)                 # block 0 (gate-on) = warm-up
    t0 = time.perf_counter(); blocks = 0
    while True:
        bank.process_block(FRAMES); blocks += 1
        if time.perf_counter() - t0 >= budget_s:
            break
    dt = time.perf_counter() - t0
    return blocks * n_voices * FRAMES / dt, blocks, dt


def reference_arm(args, rank, world, out):
    """the reference's CPU path on the box's host cores (oracle port; all threads), same block schedule as the GPU arm, each
    step a bounded sample of the bank: the first 4 096 of the 65 536 voices"""
    if rank != 0:
        return
    import oracle_lib as O
    cores = host_cores()
    n_sample = 4096
    wl = Workload(args.voices, 0, n_sample)
    bank = O.OracleBank(O.K_FM, n_sample, VPI, SR, threads=cores)
    for b in range(args.warmup):
        wl.apply(bank, b, EV_DT, O.EV_GATE); bank.process_block(FRAMES)
    t0 = time.perf_counter()
    for b in range(args.warmup, args.warmup + args.steps):
        wl.apply(bank, b, EV_DT, O.EV_GATE); bank.process_block(FRAMES); bank.read_mix()
    dt = time.perf_counter() - t0
    value = args.steps * n_sample * FRAMES / dt
    sample = f"{n_sample} of {args.voices} voices x {FRAMES} frames per step (a rate: voice-samples rendered / s), {cores} threads (instances statically partitioned)"
    cfg = headline_config(args, args.gpus, "p2p")
    line = {
        "impl": "reference", "metric": "voice-samples/sec", "value": value, "unit": "voice-samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "voice-samples/s", "cores": cores, "kind": "port", "sample": sample, "cpu": cpu_model()},
        "e2e": {"value": value, "unit": "voice-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C restatement of the reference CPU path (oracle/); the Rust crate needs nightly cargo + un-vendored crates, absent here",
    }
    out.emit(line)


class QuietStdout:
    """fd 1 -> stderr while the run is in progress (NCCL and friends print banners to stdout);
    the one JSON line goes to the real stdout"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, obj):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def main():
    with QuietStdout() as out:
        run(out)


class Ctx:
    """rank / world plumbing shared by the headline run and the other configs"""

    def __init__(self, rank, world, local_rank):
        self.rank, self.world, self.local_rank = rank, world, local_rank

    def barrier(self, *engines):
        import torch
        import torch.distributed as dist
        for e in engines:
            e.synchronize()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        import torch
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


def device_timed_pass(ctx, eng, prepare, steps, warmup, sampler=None):
    """W untimed + K timed steps; returns (step_ms[K], kern_ms[K], kernel name per step, wall seconds of the timed loop)"""
    for b in range(warmup):
        prepare(eng, b); eng.l2_flush(); eng.process_block(FRAMES)
    ctx.barrier(eng)
    if sampler:
        sampler.arm()
    eng.enable_timing(True)
    names = []
    launches0 = eng.stats()["kernel_launches"]
    t0 = time.perf_counter()
    for b in range(warmup, warmup + steps):
        prepare(eng, b)
        eng.l2_flush()
        eng.process_block(FRAMES)
        names.append(eng.last_kernel_name())
    step_ms, kern_ms = eng.timing_drain()
    ctx.barrier(eng)
    wall = time.perf_counter() - t0
    if sampler:
        sampler.disarm()
    eng.enable_timing(False)
    assert len(step_ms) == steps, (len(step_ms), steps)
    device_timed_pass.launches = eng.stats()["kernel_launches"] - launches0
    return step_ms, kern_ms, names, wall


def e2e_pass(ctx, eng, prepare, steps, warmup, channels=1, sampler=None):
    """the same schedule through the C ABI with host buffers: per step [host calls of the block + process_block + read_mix]
    on the host clock; L2 flushed and the device idle before each step, outside the clock"""
    mix = np.empty((FRAMES, channels), dtype=np.float32)
    for b in range(warmup):
        prepare(eng, b); eng.l2_flush(); eng.process_block(FRAMES); eng.read_mix(mix)
    ctx.barrier(eng)
    if sampler:
        sampler.arm()
    st0 = eng.stats()
    total, per_step = 0.0, []
    for b in range(warmup, warmup + steps):
        eng.l2_flush(); eng.synchronize()
        t0 = time.perf_counter()
        prepare(eng, b)                        # host buffers in: events / parameter calls of this block
        eng.process_block(FRAMES)
        eng.read_mix(mix)                      # D2H of the step's result; synchronises
        dt = time.perf_counter() - t0
        total += dt; per_step.append(1e3 * dt)
    ctx.barrier(eng)
    if sampler:
        sampler.disarm()
    st1 = eng.stats()
    return total, per_step, st0, st1


def fm_alg_bytes(n):
    n_cta = (n + CTA - 1) // CTA
    return n * (2 * FM_STATE_WORDS * 4 + 4) + (n // VPI) * FRAMES * 4 + n_cta * FRAMES * 4


def kernels_report(names, kern_ms):
    """per kernel variant that ran inside the timed window: launches, mean / min / max device ms"""
    rep = {}
    for nm, ms in zip(names, kern_ms):
        r = rep.setdefault(nm, [])
        r.append(float(ms))
    return [{"kernel": nm, "launches": len(v), "mean_ms": sum(v) / len(v), "min_ms": min(v), "max_ms": max(v)} for nm, v in rep.items()]


# ------------------------------------------------------------------------------------------------ other configs
def other_config_run(ctx, abi, shard, name, kind, n_total, vpi, state_words, n_inputs, steps, setup, scaling, keep=None,
                     plan=None, mix="p2p", in_bytes=0, model_n=None, graph_file=None, specialize=False):
    """one BASELINE config at its own size: device-timed (L2 flushed), voices sharded over the ranks for N > 1"""
    world, rank = ctx.world, ctx.rank
    if scaling == "strong":
        first, n = shard.shard_range(n_total, rank, world, max(vpi, CTA))
    else:
        first, n = rank * n_total, n_total
    keep = (abi.KEEP_INSTANCE_OUT if vpi > 1 else 0) if keep is None else keep
    plan = abi.PLAN_FUSED if plan is None else plan
    if graph_file is None:
        eng = abi.Engine(kind, n, vpi, SR, ctx.local_rank, plan, keep)
    else:
        # the graph compiler's path: the same graph from its `graph!` text, ONE kernel generated for it and compiled with NVRTC
        text = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "graphs", graph_file)).read()
        eng = abi.Engine.from_graph(abi.Graph(text), n, vpi, SR, ctx.local_rank, keep, plan)
    mix_mode = shard.connect_mix_bus(eng, rank, world, mix)
    prepare = setup(eng, first, n)
    if specialize:
        # the variant for the bank-uniform inputs; `frequency` (input 0) is per voice in this workload — block 0 writes it — so it is
        # marked per-voice before the call and stays in a register
        eng.set_value(0, np.full(n, 440.0, dtype=np.float32))
        eng.specialize()
    step_ms, kern_ms, names, _ = device_timed_pass(ctx, eng, prepare, steps, 3)
    dev_ms = ctx.max_over_ranks(float(np.sum(step_ms)))
    voices_all = n_total if scaling == "strong" else n_total * world
    value = steps * voices_all * FRAMES / (dev_ms * 1e-3)
    k_ms = float(np.mean(kern_ms))
    ch = eng.channels
    alg = n * (2 * state_words * 4 + 4 * n_inputs) + n * FRAMES * in_bytes \
        + ((n // vpi) * FRAMES * 4 * ch if keep & abi.KEEP_INSTANCE_OUT else 0) + (n * FRAMES * 4 if keep & abi.KEEP_VOICE_OUT else 0) \
        + ((n + CTA - 1) // CTA) * FRAMES * 4 * ch
    if model_n is not None:
        alg = n * FRAMES * model_n
    peak, _ = measured_peaks()
    kname = max(set(names), key=names.count)
    traffic, tsrc = ncu_traffic(f"{kname}@{n}")
    res = {"config": name, "voices_total": voices_all, "voices_per_gpu": n, "scaling": scaling, "value": value, "unit": "voice-samples/s",
           "realtime_voices": value / SR, "ms_per_step": dev_ms / steps, "kernel_ms": k_ms, "steps": steps, "mix": mix_mode,
           "kernel": kname,
           "roofline": {"bound": "hbm", "achieved": alg / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg / (k_ms * 1e-3) / 1e9 / peak, "traffic": traffic, "traffic_source": tsrc,
                        "algorithmic_bytes_per_launch": alg,
                        "byte_model": "N (node-granular streams in HBM)" if model_n is not None else "F (fused, as built)"}}
    ctx.barrier(eng)
    eng.close()
    return res


def setups(abi):
    """name -> setup(eng, first_voice, n) -> prepare(eng, block): the synthetic inputs of SURVEY.md §8d"""
    def fm(eng, first, n, n_total=None):
        wl = Workload(n_total or (first + n), first, n)
        return lambda e, b: wl.apply(e, b, abi.EVENT_DTYPE, abi.EVENT_GATE)

    def twin(eng, first, n):
        eng.stream_in_lcg(0, 1000, first)                   # audio_in = U[-1,1) LCG(seed 1000 + global voice), generated on the device
        return lambda e, b: None

    def sat(eng, first, n):
        u = (lcg_cached(5, 262144)[first:first + n] / 2147483648.0) if first + n <= 262144 else np.random.default_rng(5).random(n)
        eng.set_value(0, (2000.0 * (0.5 + u)).astype(np.float32))
        return lambda e, b: None

    def epiano(eng, first, n):
        g = np.arange(first, first + n, dtype=np.int64)
        notes = 36 + (g * 5) % 48
        eng.set_value(0, (440.0 * np.exp2((notes - 69) / 12.0)).astype(np.float32))
        ev = np.zeros(n // 2, dtype=abi.EVENT_DTYPE)                 # half of each instance's 16 voices play
        ev["voice"] = np.arange(0, n, 2); ev["frame_offset"] = (g[::2] * 3) % FRAMES
        ev["input_id"] = abi.EVENT_GATE; ev["scalar"] = 0.25 + 0.75 * ((g[::2] * 37) % 101) / 101.0

        def prepare(e, b):
            if b == 0:
                e.push_events(ev)
        return prepare
    return {"fm": fm, "twin": twin, "sat": sat, "epiano": epiano}


def other_configs(ctx, args, abi, shard):
    S = setups(abi)
    out = []
    steps = args.other_steps
    single = ctx.world == 1
    scal = "weak" if single else "strong"
    # config 4: TwinPeaks x 1 048 576 voices, voice-sharded over the N GPUs, mix bus exchanged (strong scaling)
    out.append(other_config_run(ctx, abi, shard, "configs[3]: nih-twin-peaks (2 x LP18Filter, a - b) x 1 048 576 voices, device-generated LCG audio_in"
                                + ("" if single else f", voice-sharded over {ctx.world} GPUs with the mix-bus exchange"),
                                abi.GRAPH_TWIN_PEAKS, 1048576, 1, 20, 3, steps, S["twin"], scal, mix=args.mix))
    # config 5: SatGraph_4x x 262 144 voices
    out.append(other_config_run(ctx, abi, shard, "configs[4]: oversampled-saturator SatGraph_4x x 262 144 voices"
                                + ("" if single else f", voice-sharded over {ctx.world} GPUs"),
                                abi.GRAPH_SAT_4X, 262144, 1, 49, 2, steps, S["sat"], scal, mix=args.mix))
    if single:
        # config 3: ElectricPiano, 16-voice polyphony x 131 072 instances
        out.append(other_config_run(ctx, abi, shard, "configs[2]: electric-piano (32-harmonic additive voice, 16-voice polyphony) x 131 072 instances = 2 097 152 voices, stereo",
                                    abi.GRAPH_EPIANO, 131072 * 16, 16, 259, 9, max(3, steps // 2), S["epiano"], "weak"))
        out.append(other_config_run(ctx, abi, shard, "fm-synth FMVoice x 1 048 576 voices (fused plan)", abi.GRAPH_FM_VOICE, 1048576, VPI,
                                    FM_STATE_WORDS, 1, steps, lambda e, f, n: S["fm"](e, f, n, 1048576), "weak"))
        out.append(other_config_run(ctx, abi, shard, "fm-synth FMVoice x 1 048 576 voices, NODE plan (one kernel per node type, streams in HBM, CUDA graph)",
                                    abi.GRAPH_FM_VOICE, 1048576, VPI, FM_STATE_WORDS, 1, max(3, steps // 2), lambda e, f, n: S["fm"](e, f, n, 1048576), "weak",
                                    plan=abi.PLAN_NODES, model_n=MODEL_N_BYTES))
        out.append(other_config_run(ctx, abi, shard, "fm-synth FMVoice x 65 536 voices, NODE plan", abi.GRAPH_FM_VOICE, 65536, VPI, FM_STATE_WORDS, 1,
                                    steps, lambda e, f, n: S["fm"](e, f, n, 65536), "weak", plan=abi.PLAN_NODES, model_n=MODEL_N_BYTES))
    if single:
        # the graph compiler at bench scale: the two reference graphs from their own `graph!` text
        for nm, kw in (("graph compiler: FMVoice from tests/graphs/voice_fm3.graph, generated kernel (NVRTC) + oscen_b200_specialize, 65 536 voices, blocks 3.. of the bench schedule (decay stages; from block 10 the ramp blocks, which a generated kernel renders on its per-frame path)",
                        dict(kind=abi.GRAPH_FM_VOICE, n_total=65536, vpi=VPI, state_words=FM_STATE_WORDS, n_inputs=1, steps=steps,
                             setup=lambda e, f, n: S["fm"](e, f, n, 65536), graph_file="voice_fm3.graph", specialize=True)),
                       ("graph compiler: ElectricPiano from tests/graphs/electric_piano.graph (array-typed streams, Tremolo -> Frame<2>), generated kernel, 16 384 instances x 16 voices",
                        dict(kind=abi.GRAPH_EPIANO, n_total=16384 * 16, vpi=16, state_words=259, n_inputs=9, steps=max(3, steps // 2),
                             setup=S["epiano"], graph_file="electric_piano.graph"))):
            try:
                out.append(other_config_run(ctx, abi, shard, nm, kw.pop("kind"), kw.pop("n_total"), kw.pop("vpi"), kw.pop("state_words"),
                                            kw.pop("n_inputs"), kw.pop("steps"), kw.pop("setup"), "weak", **kw))
            except Exception as ex:              # a box without libnvrtc: reported, never silently replaced by another path
                out.append({"config": nm, "error": repr(ex)[:300]})
    # north_star operating point: 1 250 000-class FM bank PER GPU (10 M voices on 8 GPUs), weak scaling
    n_ns = 1250048                      # a multiple of 64
    r = other_config_run(ctx, abi, shard, f"north_star point: fm-synth FMVoice x {n_ns} voices per GPU ({n_ns * ctx.world} voices on {ctx.world} GPU(s))",
                         abi.GRAPH_FM_VOICE, n_ns, VPI, FM_STATE_WORDS, 1, steps, lambda e, f, n: S["fm"](e, f, n, n_ns * ctx.world), "weak", mix=args.mix)
    r["north_star"] = {"target_realtime_voices_8gpu": 10_000_000, "realtime_voices_this_run": r["realtime_voices"],
                       "per_gpu_realtime_voices": r["realtime_voices"] / ctx.world,
                       "limiter": "graph kernel (instruction issue): exchange + host are ms_per_step - kernel_ms",
                       "exchange_and_host_ms_per_step": r["ms_per_step"] - r["kernel_ms"]}
    out.append(r)
    return out


def mix_check(ctx, args, abi, shard, eng_factory, wl):
    """N > 1: one untimed run of 3 blocks — the exchanged mix bus of the benchmarked configuration against
    all_reduce(sum) of the ranks' local mix buses, and bit-for-bit across ranks"""
    import torch
    import torch.distributed as dist
    local = eng_factory()
    bus = eng_factory()
    mode = shard.connect_mix_bus(bus, ctx.rank, ctx.world, args.mix)
    worst, identical = 0.0, True
    for b in range(3):
        for e in (local, bus):
            wl.apply(e, b, abi.EVENT_DTYPE, abi.EVENT_GATE); e.process_block(FRAMES)
        want = torch.from_numpy(local.read_mix().astype(np.float64)).cuda()
        dist.all_reduce(want, op=dist.ReduceOp.SUM)
        got = bus.read_mix()
        worst = max(worst, float(np.abs(got.astype(np.float64) - want.cpu().numpy()).max()))
        g = torch.from_numpy(got.view(np.int32).copy()).cuda()
        lo, hi = g.clone(), g.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        identical &= bool(torch.equal(lo, hi))
    peak = float(np.abs(got).max())
    ctx.barrier(local, bus)
    local.close(); bus.close()
    return {"max_abs_diff": ctx.max_over_ranks(worst), "ranks_identical": identical, "mix_peak": peak, "mode": mode, "blocks": 3,
            "against": "all_reduce(sum) of the ranks' local mix buses (float64)"}


def run(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--voices", type=int, default=VOICES_DEFAULT, help="voices PER GPU (weak scaling)")
    ap.add_argument("--plan", default="fused", choices=["fused", "nodes"])
    ap.add_argument("--mix", default="p2p", choices=["p2p", "nccl"],
                    help="N>1 mix-bus exchange: fused reduce+exchange kernel over NVLink peer memory, or ncclAllReduce")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--other-steps", type=int, default=10)
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world, out)
        return

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout for the one JSON line
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    from oscen_b200 import abi, build, shard
    build.build()
    ctx = Ctx(rank, world, local_rank)
    n = args.voices
    plan = abi.PLAN_FUSED if args.plan == "fused" else abi.PLAN_NODES
    wl = Workload(n * world, rank * n, n)

    def new_engine(connect=True):
        e = abi.Engine(abi.GRAPH_FM_VOICE, n, VPI, SR, local_rank, plan, abi.KEEP_INSTANCE_OUT)
        mode = shard.connect_mix_bus(e, rank, world, args.mix) if connect else "local"
        return e, mode

    def prepare(e, b):
        wl.apply(e, b, abi.EVENT_DTYPE, abi.EVENT_GATE)

    # the clock sampler is up (thread started, NVML loaded) BEFORE any barrier that opens a timed region
    if rank == 0 and not os.environ.get("OSCEN_BENCH_NO_SAMPLER"):
        try:
            gpu_uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
        except Exception:
            gpu_uuid = None
        # 10 ms: NVML queries take a driver-wide lock; polled faster, a query now and then stalls a launch for milliseconds
        # (seen at 2 ms: one timed step in ~100 took 10-18 ms)
        sampler = ClockSampler(local_rank, gpu_uuid, 0.010)
    else:
        sampler = NoSampler()
    sampler.begin()

    # ---------------- pass 1: device-timed (inputs resident), L2 flushed between steps
    eng, mix_mode = new_engine()
    step_ms, kern_ms, names, t_wall = device_timed_pass(ctx, eng, prepare, args.steps, args.warmup, sampler)
    launches = int(device_timed_pass.launches)
    if os.environ.get("OSCEN_BENCH_DUMP"):                # diagnostics: per-step device times of every rank
        with open(f"{os.environ['OSCEN_BENCH_DUMP']}.rank{rank}.json", "w") as f:
            json.dump({"step_ms": [float(x) for x in step_ms], "kern_ms": [float(x) for x in kern_ms], "kernels": names, "wall_s": t_wall}, f)
    dev_ms = ctx.max_over_ranks(float(np.sum(step_ms)))
    work = args.steps * n * world * FRAMES
    value = work / (dev_ms * 1e-3)
    # `launches`: kernels launched inside the timed window — graph kernel + mix-bus reduction / exchange per step, plus the
    # row refills a set_value_ramp ends with (the L2-flush fill is a benchmarking aid the engine does not count)
    ctx.barrier(eng)
    eng.close()

    # ---------------- pass 2: end to end through the C ABI with host buffers, same block schedule, fresh engine
    eng2, _ = new_engine()
    e2e_local, e2e_steps, st2, st3 = e2e_pass(ctx, eng2, prepare, args.steps, args.warmup, 1, sampler)
    e2e_s = ctx.max_over_ranks(e2e_local)
    ctx.barrier(eng2)
    eng2.close()
    clocks = sampler.stop()                       # sampled across both timed regions (device-timed pass and e2e pass)
    e2e = {"value": work / e2e_s, "unit": "voice-samples/s",
           "h2d_bytes_per_step": (st3["h2d_bytes"] - st2["h2d_bytes"]) / args.steps,
           "d2h_bytes_per_step": (st3["d2h_bytes"] - st2["d2h_bytes"]) / args.steps,
           "timed": "sum over K steps of the host clock around (this block's host calls: push_events / set_value_ramp + process_block + read_mix); "
                    "the same blocks as `value`; L2 flushed + device idle before each step, outside the clock; max over ranks",
           "ms_per_step": 1e3 * e2e_s / args.steps, "step_ms": dist_stats(e2e_steps)}

    # ---------------- roofline of the dominant kernel (this rank)
    peak, peak_src = measured_peaks()
    k_ms = float(np.mean(kern_ms))
    kernels = kernels_report(names, kern_ms)
    dominant = max(kernels, key=lambda k: k["launches"] * k["mean_ms"])["kernel"]
    if args.plan == "fused":
        alg_bytes = fm_alg_bytes(n)
        model = "F (fused, as built): state in+out + frequency + kept instance outputs + mix partials"
    else:
        alg_bytes = n * FRAMES * MODEL_N_BYTES
        model = "N (node-granular): 4*(14 W + 14 R) + (2*116 + P)/256 bytes per voice-sample"
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(f"{dominant}@{n}")
    cap = ncu_capture(f"{dominant}@{n}")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                # what does bound the fused kernel, from the same committed capture: share of the SM's issue slots in use
                "issue_active_frac": (cap.get("issue_active_pct") or 0.0) / 100.0 if cap else None,
                "warp_instructions_per_launch": cap.get("warp_instructions") if cap else None,
                "peak_source": peak_src, "kernel": dominant, "kernel_ms": k_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "byte_model": model,
                "model_n_equivalent_frac": (MODEL_N_BYTES * n * FRAMES / (k_ms * 1e-3) / 1e9) / peak,
                "note": "the fused kernel is FP32/FP64-issue bound, not HBM bound (see DESIGN.md and profiles/); kernel_ms is the mean over "
                        "ALL timed steps, `kernels` splits it by the variant that ran"}

    line = {
        "metric": "voice-samples/sec", "value": value, "unit": "voice-samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": headline_config(args, world, mix_mode),
        "realtime_voices": value / SR,
        "wall_ms_per_step_incl_host": 1e3 * t_wall / args.steps,
        "step_ms": dist_stats(step_ms), "kern_ms": dist_stats(kern_ms), "kernels": kernels,
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }
    if world > 1:
        line["mix_check"] = mix_check(ctx, args, abi, shard, lambda: new_engine(False)[0], wl)
    if not args.no_other_configs:
        t0 = time.perf_counter()
        line["other_configs"] = other_configs(ctx, args, abi, shard)
        line["other_configs_wall_s"] = time.perf_counter() - t0
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        v_all, blocks, dt = run_oracle(2048, args.cpu_budget, cores, n)
        v_one, _, _ = run_oracle(256, 3.0, 1, n)
        line["cpu_baseline"] = {"value": v_all, "unit": "voice-samples/s", "cores": cores, "kind": "port",
                                "sample": f"first 2048 voices of the bank x {blocks} blocks of 256 frames ({dt:.1f} s), {cores} threads",
                                "single_thread": v_one, "cpu": cpu_model()}
    if rank == 0:
        out.emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
