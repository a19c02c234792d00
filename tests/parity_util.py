"""Shared helpers of the GPU parity tests: one driving script runs on the oracle bank and on the
CUDA engine (same method names), outputs are compared sample by sample."""
import numpy as np

RTOL, ATOL = 1e-5, 1e-6          # BASELINE.json north_star: +-1e-5 relative (abs floor 1e-6)


def lcg(seed):
    """64-bit LCG of oscen-lib/tests/realtime_safety.rs:21-33 (unit-interval variant for inputs)"""
    state = (seed * 2862933555777941757 + 3037000493) & 0xFFFFFFFFFFFFFFFF
    while True:
        state = (state * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        yield state >> 33


def lcg_noise(n, seed):
    g = lcg(seed)
    return np.array([np.float32(np.float32(next(g)) / np.float32(2147483648.0)) - np.float32(1.0) for _ in range(n)],
                    dtype=np.float32)


def lcg_uniform(n, seed):
    g = lcg(seed)
    return np.array([next(g) / 2147483648.0 for _ in range(n)], dtype=np.float64)


def lcg_ints(n, seed, mod):
    g = lcg(seed)
    return np.array([next(g) % mod for _ in range(n)], dtype=np.int64)


def assert_close(got, want, what="", rtol=RTOL, atol=ATOL):
    got = np.asarray(got); want = np.asarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert np.isfinite(got).all(), what
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    tol = atol + rtol * np.abs(want.astype(np.float64))
    bad = err > tol
    if bad.any():
        idx = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.size} out of tolerance; worst at {idx}: "
                             f"got {got[idx]!r} want {want[idx]!r} err {err[idx]:.3e}")
    return float(err.max()), float((got == want).mean())


def seq_sum_f32(x, axis):
    """sequential left fold in f32 (== Rust `.iter().sum::<f32>()`)"""
    x = np.moveaxis(np.asarray(x, dtype=np.float32), axis, 0)
    acc = np.zeros(x.shape[1:], dtype=np.float32)
    for i in range(x.shape[0]):
        acc = (acc + x[i]).astype(np.float32)
    return acc


def lcg_noise_block(seeds, frames, state=None):
    """lcg_noise() for a whole bank at once: seeds [n] (uint64) -> (x [frames, n] f32, state) — `state` continues the streams
    across blocks.  Same values as lcg_noise(frames, seed) per column (the engine's stream_in_lcg / lcg_next)."""
    with np.errstate(over="ignore"):
        if state is None:
            state = np.asarray(seeds, dtype=np.uint64) * np.uint64(2862933555777941757) + np.uint64(3037000493)
        x = np.empty((frames, state.size), dtype=np.float32)
        for f in range(frames):
            state = state * np.uint64(6364136223846793005) + np.uint64(1442695040888963407)
            x[f] = (state >> np.uint64(33)).astype(np.float32) / np.float32(2147483648.0) - np.float32(1.0)
    return x, state


def strided_instances(n_inst, inst_per_cta, count):
    """`count` instance indices spread over ALL CTAs of the bank (one thread = one voice, 64-thread CTAs): consecutive picks
    land in different CTAs and rotate through the positions inside a CTA, first and last instance included"""
    n_cta = (n_inst + inst_per_cta - 1) // inst_per_cta
    count = min(count, n_inst)
    ctas = (np.arange(count, dtype=np.int64) * n_cta) // count
    idx = np.minimum(ctas * inst_per_cta + (np.arange(count) % inst_per_cta), n_inst - 1)
    idx[0] = 0; idx[-1] = n_inst - 1
    return np.unique(idx)
