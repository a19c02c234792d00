#!/usr/bin/env python
"""bench.py — headline benchmark of the oscen_b200 engine.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--voices V]

Metric (BASELINE.json): voice-samples/sec (voices x frames rendered / second) of the fm-synth
voice bank, plus the fraction of the HBM roofline.  Workload = BASELINE.json configs[1]:
FMVoice x 65 536 voices (8 192 FMGraph instances x 8 voices), 48 kHz, 256-frame blocks, synthetic
voice bank of SURVEY.md §8d (LCG seeds 1-3: note, velocity, gate-on frame; gate-off in block 150;
ramped parameters in block 10).  A "step" is one process_block(256) over the whole bank.

Timing: W warm-up steps, then exactly K steps between barrier + synchronize; every step is timed
on the device with CUDA events recorded on the engine's stream, L2 (126 MB) is flushed between
steps by a 256 MiB write that is outside the event bracket; `value` = work / sum of step times,
max over ranks.  `e2e` is the same K steps driven through the C ABI with host buffers (events
in, mix bus out) and timed with the host clock around the whole loop.

`--impl reference` times the reference's CPU path (the oracle's C restatement — the Rust crate
cannot be built here) on all host cores on a bounded sample of the same workload.
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


def lcg_stream(seed, n):
    """64-bit LCG of oscen-lib/tests/realtime_safety.rs:21-33, vectorised: returns state >> 33"""
    out = np.empty(n, dtype=np.uint64)
    state = (seed * 2862933555777941757 + 3037000493) & 0xFFFFFFFFFFFFFFFF
    for i in range(n):
        state = (state * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        out[i] = state >> 33
    return out


class Workload:
    """SURVEY.md §8d config 2, for the voices [first, first + n) of the global bank"""

    def __init__(self, n_total, first, n):
        idx = slice(first, first + n)
        notes = 36 + (lcg_stream(1, n_total) % 60).astype(np.int64)
        self.freq = (440.0 * np.exp2((notes - 69) / 12.0)).astype(np.float32)[idx]
        self.vel = (0.25 + 0.75 * (lcg_stream(2, n_total) / 2147483648.0)).astype(np.float32)[idx]
        self.on = (lcg_stream(3, n_total) % FRAMES).astype(np.uint32)[idx]
        self.off = (lcg_stream(4, n_total) % FRAMES).astype(np.uint32)[idx]
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
    """SM clock and clock-event (throttle) reasons DURING the timed region (B200_PROFILING.md's clocks
    line).  In-process NVML polling every ~10 ms (the timed region of the default run is only tens of
    milliseconds; an `nvidia-smi -lms` child needs seconds to start on an 8-GPU box and then misses
    it); falls back to the nvidia-smi loop when NVML cannot be loaded.  Only rank 0 samples (its own
    GPU): NVML queries take a driver-wide lock, and with 8 ranks polling at 250 Hz each a few steps per
    run stalled for 5-15 ms (profiles/r01_diag_8gpu_nvml_sampler.log: 0.41 ms/step with all ranks
    sampling, 0.29 without)."""
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
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))); self.mx.append(mx)
            try:
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            for name, bit in names:
                if mask & bit:
                    self.reasons.add(name)
            time.sleep(self.period_s)

    def _smi_loop(self):
        self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-lms", "20", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
        self.source = "nvidia-smi"
        for line in self.proc.stdout:
            self._ready.set()
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
        """start sampling and wait until the first sample can arrive, then drop what was seen while idle"""
        self.start()
        self._ready.wait(timeout=10.0)
        self.sm.clear(); self.mx.clear(); self.reasons.clear()

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


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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


def run_oracle(n_voices, budget_s, threads, wl_total):
    """time the oracle port on a bounded sample of the workload: the first n_voices voices"""
    import oracle_lib as O
    wl = Workload(wl_total, 0, n_voices)
    bank = O.OracleBank(O.K_FM, n_voices, VPI, SR, threads=threads)
    wl.apply(bank, 0, np.dtype([("voice", "<u4"), ("frame_offset", "<u4"), ("input_id", "<u4"), ("scalar", "<f4")]), O.EV_GATE)
    bank.process_block(FRAMES)                 # block 0 (gate-on) = warm-up
    t0 = time.perf_counter(); blocks = 0
    while True:
        bank.process_block(FRAMES); blocks += 1
        if time.perf_counter() - t0 >= budget_s:
            break
    dt = time.perf_counter() - t0
    return blocks * n_voices * FRAMES / dt, blocks, dt


def reference_arm(args, rank, world, out):
    if rank != 0:
        return
    import oracle_lib as O
    cores = host_cores()
    n_sample = 4096                                   # bounded sample: 1/16 of the 65 536-voice bank per step
    wl = Workload(args.voices, 0, n_sample)
    dt_ev = np.dtype([("voice", "<u4"), ("frame_offset", "<u4"), ("input_id", "<u4"), ("scalar", "<f4")])
    bank = O.OracleBank(O.K_FM, n_sample, VPI, SR, threads=cores)
    for b in range(args.warmup):
        wl.apply(bank, b, dt_ev, O.EV_GATE); bank.process_block(FRAMES)
    t0 = time.perf_counter()
    for b in range(args.warmup, args.warmup + args.steps):
        wl.apply(bank, b, dt_ev, O.EV_GATE); bank.process_block(FRAMES); bank.read_mix()
    dt = time.perf_counter() - t0
    value = args.steps * n_sample * FRAMES / dt
    sample = f"{n_sample} of {args.voices} voices x {FRAMES} frames per step, {cores} threads (instances statically partitioned)"
    line = {
        "impl": "reference", "metric": "voice-samples/sec", "value": value, "unit": "voice-samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "fm-synth FMVoice bank, 65536 voices (8192 x 8), 48 kHz, 256-frame blocks (BASELINE configs[1])",
                   "voices": args.voices, "frames": FRAMES, "sample_rate": SR, "sample": sample, "cpu": cpu_model()},
        "cpu_baseline": {"value": value, "unit": "voice-samples/s", "cores": cores, "kind": "port", "sample": sample},
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

    from oscen_b200 import abi, build
    build.build()
    n = args.voices
    plan = abi.PLAN_FUSED if args.plan == "fused" else abi.PLAN_NODES
    eng = abi.Engine(abi.GRAPH_FM_VOICE, n, VPI, SR, local_rank, plan, abi.KEEP_INSTANCE_OUT)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(abi.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        eng.comm_init(rank, world, bytes(idt.cpu().numpy().tobytes()))
        mix_mode = "nccl"
        if args.mix == "p2p" and args.plan == "fused":
            try:
                mine = torch.frombuffer(bytearray(eng.comm_p2p_export(world)), dtype=torch.uint8).cuda()
                allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
                dist.all_gather(allh, mine)
                eng.comm_p2p_import(rank, world, b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
                mix_mode = "p2p"
            except abi.EngineError as ex:          # IPC not available: NCCL stays in place
                print(f"[rank {rank}] p2p mix exchange unavailable ({ex}); using NCCL", file=sys.stderr)
            ok = torch.tensor([1 if mix_mode == "p2p" else 0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0 and mix_mode == "p2p":      # some rank could not map a peer: all ranks use NCCL
                eng.comm_p2p_close(); mix_mode = "nccl"
    if world == 1:
        mix_mode = "local"
    wl = Workload(n * world, rank * n, n)
    mix = np.empty((FRAMES, 1), dtype=np.float32)

    def barrier():
        eng.synchronize(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        eng.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- pass 1: device-timed (inputs resident), L2 flushed between steps
    eng.enable_timing(True)
    for b in range(args.warmup):
        wl.apply(eng, b, abi.EVENT_DTYPE, abi.EVENT_GATE)
        eng.l2_flush(); eng.process_block(FRAMES)
    barrier()
    try:
        gpu_uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(local_rank, gpu_uuid, 0.010 if world == 1 else 0.025)   # fewer driver-lock stalls when 8 ranks share a box
    if rank != 0 or os.environ.get("OSCEN_BENCH_NO_SAMPLER"):   # rank 0 reports the clocks line (env: diagnostics only)
        sampler.stop = lambda: {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "disabled"}
    else:
        sampler.begin()
    st0 = eng.stats()
    step_ms, kern_ms = [], []
    t_wall0 = time.perf_counter()
    for b in range(args.warmup, args.warmup + args.steps):
        wl.apply(eng, b, abi.EVENT_DTYPE, abi.EVENT_GATE)
        eng.l2_flush()
        eng.process_block(FRAMES)
        step_ms.append(eng.last_block_ms()); kern_ms.append(eng.last_kernel_ms())
    barrier()
    t_wall = time.perf_counter() - t_wall0
    if os.environ.get("OSCEN_BENCH_DUMP"):                # diagnostics: per-step device times of every rank
        with open(f"{os.environ['OSCEN_BENCH_DUMP']}.rank{rank}.json", "w") as f:
            json.dump({"step_ms": step_ms, "kern_ms": kern_ms, "wall_s": t_wall}, f)
    st1 = eng.stats()
    dev_ms = max_over_ranks(sum(step_ms))
    work = args.steps * n * world * FRAMES
    value = work / (dev_ms * 1e-3)
    # graph kernel + mix-bus reduction per step (the L2-flush fill is a benchmarking aid and is not counted by the engine)
    launches = st1["kernel_launches"] - st0["kernel_launches"]

    # ---------------- pass 2: end to end through the C ABI with host buffers (events in, mix out)
    eng.enable_timing(False)
    b0 = args.warmup + args.steps
    for b in range(b0, b0 + 3):
        wl.apply(eng, b, abi.EVENT_DTYPE, abi.EVENT_GATE); eng.l2_flush(); eng.process_block(FRAMES); eng.read_mix(mix)
    barrier()
    st2 = eng.stats()
    e2e_local = 0.0
    for k in range(args.steps):
        # the same event schedule as pass 1, shifted: gate-on burst at step 0, gate-off burst at step 130
        blk = {0: 0, 130: 150}.get(k, 1000 + k)
        ev = wl.events(blk, abi.EVENT_DTYPE, abi.EVENT_GATE)
        eng.l2_flush(); eng.synchronize()      # cold L2 for every step; the flush itself is not part of the step
        t0 = time.perf_counter()
        if ev is not None:
            eng.push_events(ev)                # H2D of the step's inputs (host buffers)
        eng.process_block(FRAMES)
        eng.read_mix(mix)                      # D2H of the step's result; synchronises
        e2e_local += time.perf_counter() - t0
    barrier()
    e2e_s = max_over_ranks(e2e_local)
    clocks = sampler.stop()                       # sampled across both timed regions (device-timed pass and e2e pass)
    st3 = eng.stats()
    e2e = {"value": work / e2e_s, "unit": "voice-samples/s",
           "h2d_bytes_per_step": (st3["h2d_bytes"] - st2["h2d_bytes"]) / args.steps,
           "d2h_bytes_per_step": (st3["d2h_bytes"] - st2["d2h_bytes"]) / args.steps,
           "timed": "sum over K steps of the host clock around (push_events + process_block + read_mix); L2 flushed + synchronised before each step, outside the clock; max over ranks",
           "ms_per_step": 1e3 * e2e_s / args.steps}

    # ---------------- roofline of the dominant kernel (this rank)
    peak, peak_src = measured_peaks()
    k_ms = statistics.mean(kern_ms)
    n_cta = (n + 127) // 128
    if args.plan == "fused":
        kernel = "fm_voice_kernel<false> (whole FMVoice fused; falls to <true> in blocks with ramps)"
        alg_bytes = n * (2 * FM_STATE_WORDS * 4 + 4) + (n // VPI) * FRAMES * 4 + n_cta * FRAMES * 4
        model = "F (fused, as built): state in+out + frequency + kept instance outputs + mix partials"
    else:
        kernel = "FMVoice node plan: 14 node kernels + instance mix, CUDA graph"
        alg_bytes = n * FRAMES * MODEL_N_BYTES
        model = "N (node-granular): 4*(14 W + 14 R) + (2*116 + P)/256 bytes per voice-sample"
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    # DRAM bytes of one launch from the committed `ncu --set full` capture of this kernel at this
    # size (profiles/r01_ncu_fm_fused_v2.md: dram__bytes_read.sum 7.92 MB + dram__bytes_write.sum 0:
    # the outputs and the written-back state of a 24 MB working set are still in the 126 MB L2 when the
    # kernel ends); null for configurations that were not captured
    traffic = 7920896 if (args.plan == "fused" and n == 65536) else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": "profiles/r01_ncu_fm_fused_v2.md" if traffic else None,
                "peak_source": peak_src, "kernel": kernel, "kernel_ms": k_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "byte_model": model,
                "model_n_equivalent_frac": (MODEL_N_BYTES * n * FRAMES / (k_ms * 1e-3) / 1e9) / peak,
                "note": "the fused kernel is FP32/FP64-issue bound, not HBM bound; see DESIGN.md and profiles/"}

    line = {
        "metric": "voice-samples/sec", "value": value, "unit": "voice-samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "fm-synth FMVoice bank, 65536 voices (8192 FMGraph instances x 8) per GPU, 48 kHz, 256-frame blocks (BASELINE configs[1])",
                   "voices_per_gpu": n, "voices_total": n * world, "frames": FRAMES, "sample_rate": SR, "plan": args.plan,
                   "parallelism": (f"voice-sharded x{world}, mix bus: " + ("fused reduce+exchange kernel over NVLink peer memory" if mix_mode == "p2p" else "ncclAllReduce")) if world > 1 else "single GPU",
                   "l2": "flushed between steps (256 MiB write on the engine stream, outside the per-step CUDA-event bracket and outside the per-step e2e host clock)",
                   "timing": "sum of per-step CUDA-event times on the engine stream, max over ranks"},
        "realtime_voices": value / SR,
        "wall_ms_per_step_incl_host": 1e3 * t_wall / args.steps,
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        v_all, blocks, dt = run_oracle(2048, args.cpu_budget, cores, n)
        v_one, _, _ = run_oracle(256, 3.0, 1, n)
        line["cpu_baseline"] = {"value": v_all, "unit": "voice-samples/s", "cores": cores, "kind": "port",
                                "sample": f"first 2048 voices of the bank x {blocks} blocks of 256 frames ({dt:.1f} s), {cores} threads",
                                "single_thread": v_one, "cpu": cpu_model()}
    if rank == 0:
        out.emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
