/*
 * kernel_resample.cuh — the cross-rate ("multirate") clipper graphs: a per-voice stream input goes
 * through an Up edge, a `* N` inner-rate HardClip and a Down edge.
 *
 *   ClipRef / ClipOversampled / ClipOversampledIir      oscen-lib/tests/multirate_graph.rs:282-325
 *   multirate frame body (warm-up, N inner ticks, finalize)   codegen/emit_frame.rs:114-176,254-514
 *   resampler kernels per connection policy:
 *     [sinc]      SincUpFir<N> / SincDownFir<N>          resample/sinc_fir.rs:146-266
 *     [sinc_iir]  IirHalfbandUp<N> / IirHalfbandDown<N>  resample/halfband_iir.rs:146-262
 *     [linear]    LinearUp<N> / LinearDown<N>            resample/linear.rs:11-74
 *     [latch]     LatchUp<N> / LatchDown<N>              resample/latch.rs:6-54
 *
 * One thread = one voice, serial in time, like every other fused kernel.  FIR histories are shift registers in
 * REGISTERS (nodes.cuh HbUpRegs / HbDownRegs; shared-memory rings until round 2), like the IIR all-pass states.
 * State in HBM is [field][voice]; FIR histories are stored chronologically (oldest first).
 */
#pragma once
#include "kernel_others.cuh"

namespace oscen {

enum : int { POL_SINC = 0, POL_SINC_IIR = 1, POL_LINEAR = 2, POL_LATCH = 3 };

/* state fields: lcg(2) then, per policy:
     sinc      L up stages x 12 taps, L down stages x 24 taps      (2 + 36 L)
     sinc_iir  L up stages x 8 words, L down stages x 9 words      (2 + 17 L)
     linear    LinearUp.prev                                       (3)
     latch     -                                                   (2)                        */
enum : uint32_t { CL_LCG = 0, CL_STATE = 2 };
OSC_HD constexpr uint32_t clip_n_state(int policy, int log2f) {
    return log2f == 0 ? 2u
         : policy == POL_SINC ? 2u + 36u * (uint32_t)log2f
         : policy == POL_SINC_IIR ? 2u + 17u * (uint32_t)log2f
         : policy == POL_LINEAR ? 3u : 2u;
}
/* dynamic shared memory: isum [kTileF][kCta] (the FIR histories left shared memory in round 2) */
OSC_HD constexpr size_t clip_dyn_bytes(int /*policy*/, int /*log2f*/) {
    return (size_t)(kTileF * kCta) * sizeof(float);
}

/* the test graphs' HardClip: clamp(+-0.7), no drive (multirate_graph.rs:35-40) */
OSC_D float clip07(float x) { return clampf(x, -0.7f, 0.7f); }

template <int POLICY, int L>
__global__ void __launch_bounds__(kCta) clip_kernel(const __grid_constant__ KernelArgs a) {
    constexpr int N = 1 << L;
    constexpr int LS = (POLICY == POL_SINC && L > 0) ? L : 1;          /* array extents (>= 1) */
    constexpr int LI = (POLICY == POL_SINC_IIR && L > 0) ? L : 1;
    __shared__ float tile[kTileF * kTileW];
    extern __shared__ float dyn[];
    float *isum = dyn;
    const uint32_t n = a.n_voices;
    const uint32_t v = blockIdx.x * kCta + threadIdx.x;
    const bool active = v < n;
    const uint32_t vs = active ? v : 0u;
    const uint32_t frames = a.ctl->frames;
    const bool lcg_on = a.ctl->lcg_on != 0u;
    const uint32_t col = tile_col(threadIdx.x, a.vpi);

    unsigned long long lcg = (unsigned long long)__float_as_uint(a.state[(size_t)(CL_LCG + 0) * n + vs]) |
                             ((unsigned long long)__float_as_uint(a.state[(size_t)(CL_LCG + 1) * n + vs]) << 32);
    HbUpRegs fir_up[LS];
    HbDownRegs fir_down[LS];
    IirHb iir_up[LI], iir_down[LI];
    float lin_prev = 0.0f;
    if (POLICY == POL_SINC && L > 0) {
#pragma unroll
        for (int s = 0; s < LS; ++s) {
            hb_up_load(fir_up[s], a.state, CL_STATE + 12u * (uint32_t)s, n, vs);
            hb_down_load(fir_down[s], a.state, CL_STATE + 12u * (uint32_t)L + 24u * (uint32_t)s, n, vs);
        }
    }
    if (POLICY == POL_SINC_IIR && L > 0) {
#pragma unroll
        for (int s = 0; s < LI; ++s) {
            const float *p = a.state + (size_t)(CL_STATE + 8u * s) * n + vs;
            iir_up[s].a0 = {p[0], p[n]}; iir_up[s].a1 = {p[2 * (size_t)n], p[3 * (size_t)n]};
            iir_up[s].b0 = {p[4 * (size_t)n], p[5 * (size_t)n]}; iir_up[s].b1 = {p[6 * (size_t)n], p[7 * (size_t)n]};
            iir_up[s].prev_odd_in = 0.0f;
            const float *q = a.state + (size_t)(CL_STATE + 8u * L + 9u * s) * n + vs;
            iir_down[s].a0 = {q[0], q[n]}; iir_down[s].a1 = {q[2 * (size_t)n], q[3 * (size_t)n]};
            iir_down[s].b0 = {q[4 * (size_t)n], q[5 * (size_t)n]}; iir_down[s].b1 = {q[6 * (size_t)n], q[7 * (size_t)n]};
            iir_down[s].prev_odd_in = q[8 * (size_t)n];
        }
    }
    if (POLICY == POL_LINEAR && L > 0) lin_prev = a.state[(size_t)CL_STATE * n + vs];

    for (uint32_t f0 = 0; f0 < frames; f0 += kTileF) {
        const uint32_t fe = min(kTileF, frames - f0);
        /* this tile's input column: up to 32 independent coalesced streaming loads in flight */
        if (active && !lcg_on && a.stream_in != nullptr) {
#pragma unroll 8
            for (uint32_t ff = 0; ff < fe; ++ff) isum[ff * kCta + threadIdx.x] = __ldcs(a.stream_in + (size_t)(f0 + ff) * n + v);
        }
#pragma unroll 1
        for (uint32_t ff = 0; ff < fe; ++ff) {
            float y = 0.0f;
            if (active) {
                const float x = lcg_on ? lcg_next(lcg) : ((a.stream_in != nullptr) ? isum[ff * kCta + threadIdx.x] : 0.0f);
                if (L == 0) {
                    y = clip07(x);
                } else {
                    float buf[N];
                    /* ---- Up edge warm-up: one outer sample -> N inner samples (emit_frame.rs:254-307) */
                    if (POLICY == POL_SINC) {
                        buf[0] = x;
#pragma unroll
                        for (int s = 0; s < L; ++s) {                  /* stage s: 2^s steps, in index order */
                            float nx[N];
#pragma unroll
                            for (int i = 0; i < N / 2; ++i) if (i < (1 << s))
                                hb_up_step_r(fir_up[s], buf[i], nx[2 * i], nx[2 * i + 1]);
#pragma unroll
                            for (int i = 0; i < N; ++i) if (i < (2 << s)) buf[i] = nx[i];
                        }
                    } else if (POLICY == POL_SINC_IIR) {
                        buf[0] = x;
#pragma unroll
                        for (int s = 0; s < L; ++s) {
                            float nx[N];
#pragma unroll
                            for (int i = 0; i < N / 2; ++i) if (i < (1 << s)) iir_hb_step_up(iir_up[s], buf[i], nx[2 * i], nx[2 * i + 1]);
#pragma unroll
                            for (int i = 0; i < N; ++i) if (i < (2 << s)) buf[i] = nx[i];
                        }
                    } else if (POLICY == POL_LINEAR) {                 /* linear.rs:23-32 */
                        const float n_inv = 1.0f / (float)N;
                        const float delta = x - lin_prev;
#pragma unroll
                        for (int i = 0; i < N; ++i) buf[i] = lin_prev + delta * ((float)i * n_inv);
                        lin_prev = x;
                    } else {                                           /* latch.rs:17-23 */
#pragma unroll
                        for (int i = 0; i < N; ++i) buf[i] = x;
                    }
                    /* ---- N inner ticks: write clip.input, process, capture (emit_frame.rs:388-461) */
#pragma unroll
                    for (int i = 0; i < N; ++i) buf[i] = clip07(buf[i]);
                    /* ---- Down edge finalize (emit_frame.rs:474-514) */
                    if (POLICY == POL_SINC) {
#pragma unroll
                        for (int s = 0; s < L; ++s) {                  /* stage s consumes pairs in order */
#pragma unroll
                            for (int i = 0; i < N / 2; ++i) if (i < (N >> (s + 1)))
                                buf[i] = hb_down_step_r(fir_down[s], buf[2 * i], buf[2 * i + 1]);
                        }
                        y = buf[0];
                    } else if (POLICY == POL_SINC_IIR) {
#pragma unroll
                        for (int s = 0; s < L; ++s) {
#pragma unroll
                            for (int i = 0; i < N / 2; ++i) if (i < (N >> (s + 1))) buf[i] = iir_hb_step_down(iir_down[s], buf[2 * i], buf[2 * i + 1]);
                        }
                        y = buf[0];
                    } else if (POLICY == POL_LINEAR) {                 /* linear.rs:59-66 */
                        float acc = 0.0f;
#pragma unroll
                        for (int i = 0; i < N; ++i) acc = acc + buf[i];
                        y = acc * (1.0f / (float)N);
                    } else {
                        y = buf[0];                                    /* latch.rs:45-49 */
                    }
                }
                if (a.voice_out != nullptr) a.voice_out[(size_t)(f0 + ff) * n + v] = y;
            }
            tile[ff * kTileW + col] = y;
        }
        tile_flush_mono(a, tile, isum, f0, fe);
    }
    if (active) {
        a.state[(size_t)(CL_LCG + 0) * n + v] = __uint_as_float((uint32_t)(lcg & 0xffffffffull));
        a.state[(size_t)(CL_LCG + 1) * n + v] = __uint_as_float((uint32_t)(lcg >> 32));
        if (POLICY == POL_SINC && L > 0) {
#pragma unroll
            for (int s = 0; s < LS; ++s) {
                hb_up_store(fir_up[s], a.state, CL_STATE + 12u * (uint32_t)s, n, v);
                hb_down_store(fir_down[s], a.state, CL_STATE + 12u * (uint32_t)L + 24u * (uint32_t)s, n, v);
            }
        }
        if (POLICY == POL_SINC_IIR && L > 0) {
#pragma unroll
            for (int s = 0; s < LI; ++s) {
                float *p = a.state + (size_t)(CL_STATE + 8u * s) * n + v;
                p[0] = iir_up[s].a0.x_prev; p[n] = iir_up[s].a0.y_prev; p[2 * (size_t)n] = iir_up[s].a1.x_prev; p[3 * (size_t)n] = iir_up[s].a1.y_prev;
                p[4 * (size_t)n] = iir_up[s].b0.x_prev; p[5 * (size_t)n] = iir_up[s].b0.y_prev; p[6 * (size_t)n] = iir_up[s].b1.x_prev; p[7 * (size_t)n] = iir_up[s].b1.y_prev;
                float *q = a.state + (size_t)(CL_STATE + 8u * L + 9u * s) * n + v;
                q[0] = iir_down[s].a0.x_prev; q[n] = iir_down[s].a0.y_prev; q[2 * (size_t)n] = iir_down[s].a1.x_prev; q[3 * (size_t)n] = iir_down[s].a1.y_prev;
                q[4 * (size_t)n] = iir_down[s].b0.x_prev; q[5 * (size_t)n] = iir_down[s].b0.y_prev; q[6 * (size_t)n] = iir_down[s].b1.x_prev; q[7 * (size_t)n] = iir_down[s].b1.y_prev;
                q[8 * (size_t)n] = iir_down[s].prev_odd_in;
            }
        }
        if (POLICY == POL_LINEAR && L > 0) a.state[(size_t)CL_STATE * n + v] = lin_prev;
    }
}

}  // namespace oscen
