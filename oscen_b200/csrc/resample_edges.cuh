/*
 * resample_edges.cuh — cross-rate edges as per-thread objects, for the kernels codegen.hpp generates
 * for multirate graphs (`node = T::new() * N`): one UpEdge / DownEdge per cross-rate stream edge,
 * with the connection policy's resampler as its state.
 *
 *   StreamUpsampler::upsample(x, &mut [F; N]) / StreamDownsampler::downsample(&[F; N]) -> F
 *                                                              oscen-lib/src/resample/mod.rs:22-43
 *   [sinc]      SincUpFir<N> / SincDownFir<N>                  resample/sinc_fir.rs:146-266
 *   [sinc_iir]  IirHalfbandUp<N> / IirHalfbandDown<N>          resample/halfband_iir.rs:146-262
 *   [linear]    LinearUp<N> / LinearDown<N>                    resample/linear.rs:11-74
 *   [latch]     LatchUp<N> / LatchDown<N>                      resample/latch.rs:6-54
 *
 * Same arithmetic and the same HBM state layout as the hand-written clip_kernel (kernel_resample.cuh):
 * FIR histories are per-thread shift registers IN REGISTERS (nodes.cuh HbUpRegs / HbDownRegs; until round 2 they were
 * shared-memory rings stored twice, which capped these kernels at 5 CTAs per SM), chronological (oldest first) in HBM;
 * IIR all-pass sections and the linear `prev` live in registers as well.
 * Device functions only (the generated translation unit includes this file).
 */
#pragma once
#include "state_io.cuh"

namespace oscen {

enum : int { EDGE_SINC = 0, EDGE_SINC_IIR = 1, EDGE_LINEAR = 2, EDGE_LATCH = 3 };

/* SoA state words per voice / shared-memory floats per thread of one edge */
OSC_HD constexpr uint32_t up_edge_state_words(int policy, int l) {
    return policy == EDGE_SINC ? 12u * (uint32_t)l : policy == EDGE_SINC_IIR ? 8u * (uint32_t)l : policy == EDGE_LINEAR ? 1u : 0u;
}
OSC_HD constexpr uint32_t down_edge_state_words(int policy, int l) {
    return policy == EDGE_SINC ? 24u * (uint32_t)l : policy == EDGE_SINC_IIR ? 9u * (uint32_t)l : 0u;
}
OSC_HD constexpr uint32_t up_edge_smem_floats(int /*policy*/, int /*l*/) { return 0u; }      /* no edge needs shared memory any more */
OSC_HD constexpr uint32_t down_edge_smem_floats(int /*policy*/, int /*l*/) { return 0u; }

OSC_D void iir_hb_load(IirHb &h, const float *p, size_t n, bool with_prev_odd) {
    h.a0 = {p[0], p[n]}; h.a1 = {p[2 * n], p[3 * n]};
    h.b0 = {p[4 * n], p[5 * n]}; h.b1 = {p[6 * n], p[7 * n]};
    h.prev_odd_in = with_prev_odd ? p[8 * n] : 0.0f;
}
OSC_D void iir_hb_store(const IirHb &h, float *p, size_t n, bool with_prev_odd) {
    p[0] = h.a0.x_prev; p[n] = h.a0.y_prev; p[2 * n] = h.a1.x_prev; p[3 * n] = h.a1.y_prev;
    p[4 * n] = h.b0.x_prev; p[5 * n] = h.b0.y_prev; p[6 * n] = h.b1.x_prev; p[7 * n] = h.b1.y_prev;
    if (with_prev_odd) p[8 * n] = h.prev_odd_in;
}

/* one outer-rate sample -> N = 2^L inner-rate samples (codegen/emit_frame.rs:254-307) */
template <int POLICY, int L>
struct UpEdge {
    static constexpr int N = 1 << L;
    static constexpr int LS = L > 0 ? L : 1;
    HbUpRegs fir[LS];                    /* [sinc] one 12-sample history per halfband stage */
    IirHb iir[LS];
    float prev;

    OSC_D void load(const float *state, uint32_t base, uint32_t n, uint32_t vs, float * /*smem: unused since the histories moved to registers*/) {
        prev = 0.0f;
        if (POLICY == EDGE_SINC) {
#pragma unroll
            for (int s = 0; s < L; ++s) hb_up_load(fir[s], state, base + 12u * (uint32_t)s, n, vs);
        } else if (POLICY == EDGE_SINC_IIR) {
#pragma unroll
            for (int s = 0; s < L; ++s) iir_hb_load(iir[s], state + (size_t)(base + 8u * s) * n + vs, n, false);
        } else if (POLICY == EDGE_LINEAR) {
            prev = state[(size_t)base * n + vs];
        }
    }
    OSC_D void store(float *state, uint32_t base, uint32_t n, uint32_t v) const {
        if (POLICY == EDGE_SINC) {
#pragma unroll
            for (int s = 0; s < L; ++s) hb_up_store(fir[s], state, base + 12u * (uint32_t)s, n, v);
        } else if (POLICY == EDGE_SINC_IIR) {
#pragma unroll
            for (int s = 0; s < L; ++s) iir_hb_store(iir[s], state + (size_t)(base + 8u * s) * n + v, n, false);
        } else if (POLICY == EDGE_LINEAR) {
            state[(size_t)base * n + v] = prev;
        }
    }
    OSC_D void up(float x, float (&buf)[N]) {
        if (POLICY == EDGE_SINC) {
            buf[0] = x;
#pragma unroll
            for (int s = 0; s < L; ++s) {                  /* stage s: 2^s steps, in index order */
                float nx[N];
#pragma unroll
                for (int i = 0; i < N / 2; ++i) if (i < (1 << s))
                    hb_up_step_r(fir[s], buf[i], nx[2 * i], nx[2 * i + 1]);
#pragma unroll
                for (int i = 0; i < N; ++i) if (i < (2 << s)) buf[i] = nx[i];
            }
        } else if (POLICY == EDGE_SINC_IIR) {
            buf[0] = x;
#pragma unroll
            for (int s = 0; s < L; ++s) {
                float nx[N];
#pragma unroll
                for (int i = 0; i < N / 2; ++i) if (i < (1 << s)) iir_hb_step_up(iir[s], buf[i], nx[2 * i], nx[2 * i + 1]);
#pragma unroll
                for (int i = 0; i < N; ++i) if (i < (2 << s)) buf[i] = nx[i];
            }
        } else if (POLICY == EDGE_LINEAR) {                /* linear.rs:23-32 */
            const float n_inv = 1.0f / (float)N;
            const float delta = x - prev;
#pragma unroll
            for (int i = 0; i < N; ++i) buf[i] = prev + delta * ((float)i * n_inv);
            prev = x;
        } else {                                           /* latch.rs:17-23 */
#pragma unroll
            for (int i = 0; i < N; ++i) buf[i] = x;
        }
    }
};

/* N = 2^L inner-rate samples -> one outer-rate sample (codegen/emit_frame.rs:474-514) */
template <int POLICY, int L>
struct DownEdge {
    static constexpr int N = 1 << L;
    static constexpr int LS = L > 0 ? L : 1;
    HbDownRegs fir[LS];                  /* [sinc] one polyphase history (12 + 7 samples) per halfband stage */
    IirHb iir[LS];

    OSC_D void load(const float *state, uint32_t base, uint32_t n, uint32_t vs, float * /*smem: unused*/) {
        if (POLICY == EDGE_SINC) {
#pragma unroll
            for (int s = 0; s < L; ++s) hb_down_load(fir[s], state, base + 24u * (uint32_t)s, n, vs);
        } else if (POLICY == EDGE_SINC_IIR) {
#pragma unroll
            for (int s = 0; s < L; ++s) iir_hb_load(iir[s], state + (size_t)(base + 9u * s) * n + vs, n, true);
        }
    }
    OSC_D void store(float *state, uint32_t base, uint32_t n, uint32_t v) const {
        if (POLICY == EDGE_SINC) {
#pragma unroll
            for (int s = 0; s < L; ++s) hb_down_store(fir[s], state, base + 24u * (uint32_t)s, n, v);
        } else if (POLICY == EDGE_SINC_IIR) {
#pragma unroll
            for (int s = 0; s < L; ++s) iir_hb_store(iir[s], state + (size_t)(base + 9u * s) * n + v, n, true);
        }
    }
    OSC_D float down(float (&buf)[N]) {
        if (POLICY == EDGE_SINC) {
#pragma unroll
            for (int s = 0; s < L; ++s) {                  /* stage s consumes pairs in order */
#pragma unroll
                for (int i = 0; i < N / 2; ++i) if (i < (N >> (s + 1)))
                    buf[i] = hb_down_step_r(fir[s], buf[2 * i], buf[2 * i + 1]);
            }
            return buf[0];
        } else if (POLICY == EDGE_SINC_IIR) {
#pragma unroll
            for (int s = 0; s < L; ++s) {
#pragma unroll
                for (int i = 0; i < N / 2; ++i) if (i < (N >> (s + 1))) buf[i] = iir_hb_step_down(iir[s], buf[2 * i], buf[2 * i + 1]);
            }
            return buf[0];
        } else if (POLICY == EDGE_LINEAR) {                /* linear.rs:59-66 */
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < N; ++i) acc = acc + buf[i];
            return acc * (1.0f / (float)N);
        }
        return buf[0];                                     /* latch.rs:45-49 */
    }
};

}  // namespace oscen
