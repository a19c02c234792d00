/*
 * kernel_nodes.cuh — the node-granular plan: ONE kernel per node type over
 * [voices x frames_per_block], inter-node streams materialised in HBM as [frame][voice] f32,
 * per-node state SoA in HBM, launched in the graph compiler's topological order
 * (oscen-graph-compiler/src/ir/lower.rs:1015-1085) and replayed as a CUDA graph.
 * This is the HBM-bound formulation BASELINE.md's byte model N describes.
 *
 * Every node kernel is "one thread = one voice, serial in time" (each stateful node is a
 * nonlinear / time-varying recurrence; a parallel scan would change the rounding order), with
 * stream inputs prefetched 8 frames deep so a warp keeps 8 x 128 B loads in flight per input.
 */
#pragma once
#include "kernels_common.cuh"
#include "kernel_fm.cuh"

namespace oscen {

constexpr uint32_t kNoInput = 0xFFFFFFFFu;
constexpr uint32_t kPF = 8;                 /* prefetch depth (frames) */

struct NodeArgs {
    uint32_t n_voices;
    uint32_t state_base;                    /* first state field of this node */
    const BlockCtl *ctl;
    const Event *events;
    const float *ptab;
    const float *pv;
    float *state;
    const float *in0, *in1;
    float *out0, *out1;
    uint32_t pid[4];                        /* value-input ids feeding this node (kNoInput: use cst) */
    float cst[4];
    float sr;
};

template <uint32_t NP>
struct NodeParams {
    float P[NP];
    OSC_D void load(const NodeArgs &a, uint32_t vs) {
#pragma unroll
        for (uint32_t k = 0; k < NP; ++k)
            P[k] = a.pid[k] == kNoInput ? a.cst[k] : a.pv[(size_t)a.pid[k] * a.n_voices + vs];
    }
    OSC_D void ramps(const NodeArgs &a, uint32_t ramp_mask, uint32_t f) {
        if (ramp_mask) {
#pragma unroll
            for (uint32_t k = 0; k < NP; ++k)
                if (a.pid[k] != kNoInput && ((ramp_mask >> a.pid[k]) & 1u)) P[k] = a.ptab[(size_t)f * kMaxInputs + a.pid[k]];
        }
    }
    OSC_D void value_event(const NodeArgs &a, const Event &e) {
#pragma unroll
        for (uint32_t k = 0; k < NP; ++k) if (e.input_id == a.pid[k]) P[k] = e.scalar;
    }
    /* apply this frame's value events; returns the [i, j) event range of the frame */
    OSC_D uint32_t events(const NodeArgs &a, EventCursor &ec, uint32_t f) {
        const uint32_t j = ec.end_of_frame(f);
        for (uint32_t q = ec.i; q < j; ++q) if (ec.ev[q].input_id != kEventGate) value_event(a, ec.ev[q]);
        return j;
    }
};

#define OSC_NODE_PROLOGUE(NP)                                                             \
    const uint32_t n = a.n_voices;                                                        \
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;                             \
    if (v >= n) return;                                                                   \
    const uint32_t frames = a.ctl->frames, ramp_mask = a.ctl->ramp_mask;                  \
    NodeParams<NP> prm; prm.load(a, v);                                                   \
    EventCursor ec; ec.init(a.events, a.ctl->n_events, v);

/* AdsrEnvelope node type: pid = {attack, decay, sustain, release}; out0 = output.
   One launch renders every AdsrEnvelope node of the graph (blockIdx.y = node): the nodes are
   mutually independent sources, and 4x the warps hide the serial recurrence's latency. */
struct NodeArgs4 { NodeArgs n[4]; };
__global__ void __launch_bounds__(kCta) adsr_node_kernel(const __grid_constant__ NodeArgs4 args) {
    const NodeArgs &a = args.n[blockIdx.y];
    OSC_NODE_PROLOGUE(4)
    AdsrState s; adsr_load(s, a.state, a.state_base, n, v);
    AdsrDerived d; adsr_derive(d, prm.P[0], prm.P[1], prm.P[3], a.sr);
    for (uint32_t f = 0; f < frames; ++f) {
        prm.ramps(a, ramp_mask, f);
        const bool due = ec.due(f);
        uint32_t j = ec.i;
        if (due) j = prm.events(a, ec, f);
        if (adsr_derived_stale(d, prm.P[0], prm.P[1], prm.P[3])) adsr_derive(d, prm.P[0], prm.P[1], prm.P[3], a.sr);
        if (due) {
            for (uint32_t q = ec.i; q < j; ++q) if (ec.ev[q].input_id == kEventGate)
                adsr_on_gate(s, ec.ev[q].scalar, prm.P[0], prm.P[3], d);
            ec.advance(j);
        }
        __stcs(a.out0 + (size_t)f * n + v, adsr_process(s, prm.P[2], d));
    }
    adsr_store(s, a.state, a.state_base, n, v);
}

/* FmOperator node: in0 = phase_mod (nullable), in1 = envelope; pid = {base_freq, ratio, feedback, level} */
__global__ void __launch_bounds__(kCta) fmop_node_kernel(const __grid_constant__ NodeArgs a) {
    OSC_NODE_PROLOGUE(4)
    FmOpState s;
    s.phase = a.state[(size_t)(a.state_base + 0) * n + v]; s.prev = a.state[(size_t)(a.state_base + 1) * n + v];
    s.inc = fmop_phase_inc(prm.P[0], prm.P[1], a.sr);
    /* software pipeline: the loads of chunk k+1 are in flight while chunk k is computed */
    float pm[kPF], en[kPF], pm_n[kPF], en_n[kPF];
#define OSC_FMOP_LOAD(PM, EN, F0)                                                             \
    _Pragma("unroll") for (uint32_t k = 0; k < kPF; ++k) {                                    \
        const uint32_t f = (F0) + k;                                                          \
        PM[k] = (a.in0 != nullptr && f < frames) ? __ldcs(a.in0 + (size_t)f * n + v) : 0.0f;  \
        EN[k] = (f < frames) ? __ldcs(a.in1 + (size_t)f * n + v) : 0.0f;                      \
    }
    OSC_FMOP_LOAD(pm, en, 0u)
    for (uint32_t f0 = 0; f0 < frames; f0 += kPF) {
        OSC_FMOP_LOAD(pm_n, en_n, f0 + kPF)
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) {
            const uint32_t f = f0 + k;
            if (f < frames) {
                if (ramp_mask) { prm.ramps(a, ramp_mask, f); s.inc = fmop_phase_inc(prm.P[0], prm.P[1], a.sr); }
                if (ec.due(f)) { ec.advance(prm.events(a, ec, f)); s.inc = fmop_phase_inc(prm.P[0], prm.P[1], a.sr); }
                __stcs(a.out0 + (size_t)f * n + v, fmop_process(s, pm[k], prm.P[2], en[k], prm.P[3]));
            }
        }
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) { pm[k] = pm_n[k]; en[k] = en_n[k]; }
    }
#undef OSC_FMOP_LOAD
    a.state[(size_t)(a.state_base + 0) * n + v] = s.phase; a.state[(size_t)(a.state_base + 1) * n + v] = s.prev;
}

/* TptFilter node: in0 = input, in1 = cutoff stream; pid = {q} */
__global__ void __launch_bounds__(kCta) tpt_node_kernel(const __grid_constant__ NodeArgs a) {
    OSC_NODE_PROLOGUE(1)
    TptState s; tpt_load(s, a.state, a.state_base, n, v);
    float x[kPF], c[kPF], x_n[kPF], c_n[kPF];
#define OSC_TPT_LOAD(X, C, F0)                                                                \
    _Pragma("unroll") for (uint32_t k = 0; k < kPF; ++k) {                                    \
        const uint32_t f = (F0) + k;                                                          \
        X[k] = (f < frames) ? __ldcs(a.in0 + (size_t)f * n + v) : 0.0f;                       \
        C[k] = (f < frames) ? __ldcs(a.in1 + (size_t)f * n + v) : 0.0f;                       \
    }
    OSC_TPT_LOAD(x, c, 0u)
    for (uint32_t f0 = 0; f0 < frames; f0 += kPF) {
        OSC_TPT_LOAD(x_n, c_n, f0 + kPF)
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) {
            const uint32_t f = f0 + k;
            if (f < frames) {
                prm.ramps(a, ramp_mask, f);
                if (ec.due(f)) ec.advance(prm.events(a, ec, f));
                __stcs(a.out0 + (size_t)f * n + v, tpt_process(s, x[k], c[k], prm.P[0], 0.0f, a.sr));
            }
        }
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) { x[k] = x_n[k]; c[k] = c_n[k]; }
    }
#undef OSC_TPT_LOAD
    tpt_store(s, a.state, a.state_base, n, v);
}

/* stateless nodes: Gain (in*param | in*const), AddValue (in+param), Mixer (a+b), Crossfade */
enum : int { EW_GAIN = 0, EW_ADD_VALUE = 1, EW_MIXER = 2, EW_CROSSFADE = 3 };
template <int OP>
__global__ void __launch_bounds__(kCta) ew_node_kernel(const __grid_constant__ NodeArgs a) {
    OSC_NODE_PROLOGUE(1)
    for (uint32_t f0 = 0; f0 < frames; f0 += kPF) {
        float x[kPF], y[kPF];
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) {
            const uint32_t f = f0 + k;
            x[k] = (f < frames) ? __ldg(a.in0 + (size_t)f * n + v) : 0.0f;
            y[k] = (OP == EW_MIXER && f < frames) ? __ldg(a.in1 + (size_t)f * n + v) : 0.0f;
        }
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) {
            const uint32_t f = f0 + k;
            if (f < frames) {
                prm.ramps(a, ramp_mask, f);
                if (ec.due(f)) ec.advance(prm.events(a, ec, f));
                if (OP == EW_GAIN) a.out0[(size_t)f * n + v] = x[k] * prm.P[0];          /* gain/mod.rs:30-35 */
                else if (OP == EW_ADD_VALUE) a.out0[(size_t)f * n + v] = x[k] + prm.P[0]; /* add_value.rs:33-35 */
                else if (OP == EW_MIXER) a.out0[(size_t)f * n + v] = x[k] + y[k];        /* mixer.rs:32-34 */
                else {                                                                    /* crossfade.rs:39-44 */
                    const float mix = clampf(prm.P[0], 0.0f, 1.0f);
                    a.out0[(size_t)f * n + v] = x[k] * (1.0f - mix);
                    a.out1[(size_t)f * n + v] = x[k] * mix;
                }
            }
        }
    }
}

/* Stateless nodes, vectorised: when the block carries no value events every element (frame, voice)
   is independent, so the kernel is a float4 grid-stride stream over [frames x voices] — the
   HBM-bound form (two 16-byte loads, one or two 16-byte stores per thread-iteration).  Requires
   n_voices % 4 == 0.  The value input is per voice (pv row) unless it is ramping this block
   (then one scalar per frame from ptab). */
template <int OP>
__global__ void __launch_bounds__(256) ew_node_vec_kernel(const __grid_constant__ NodeArgs a) {
    const uint32_t n4 = a.n_voices >> 2;
    const uint32_t frames = a.ctl->frames, ramp_mask = a.ctl->ramp_mask;
    const size_t total = (size_t)frames * n4;
    const bool has_p = a.pid[0] != kNoInput;
    const bool ramped = has_p && ((ramp_mask >> a.pid[0]) & 1u);
    const float4 *in0 = reinterpret_cast<const float4 *>(a.in0);
    const float4 *in1 = reinterpret_cast<const float4 *>(a.in1);
    const float4 *pv = has_p ? reinterpret_cast<const float4 *>(a.pv + (size_t)a.pid[0] * a.n_voices) : nullptr;
    float4 *out0 = reinterpret_cast<float4 *>(a.out0);
    float4 *out1 = reinterpret_cast<float4 *>(a.out1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t f = (uint32_t)(i / n4), v4 = (uint32_t)(i - (size_t)f * n4);
        const float4 x = __ldcs(in0 + i);
        float4 p;
        if (!has_p) p = make_float4(a.cst[0], a.cst[0], a.cst[0], a.cst[0]);
        else if (ramped) { const float r = a.ptab[(size_t)f * kMaxInputs + a.pid[0]]; p = make_float4(r, r, r, r); }
        else p = __ldg(pv + v4);
        if (OP == EW_GAIN) { __stcs(out0 + i, make_float4(x.x * p.x, x.y * p.y, x.z * p.z, x.w * p.w)); }
        else if (OP == EW_ADD_VALUE) { __stcs(out0 + i, make_float4(x.x + p.x, x.y + p.y, x.z + p.z, x.w + p.w)); }
        else if (OP == EW_MIXER) { const float4 y = __ldcs(in1 + i); __stcs(out0 + i, make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w)); }
        else {
            const float4 m = make_float4(clampf(p.x, 0.0f, 1.0f), clampf(p.y, 0.0f, 1.0f), clampf(p.z, 0.0f, 1.0f), clampf(p.w, 0.0f, 1.0f));
            __stcs(out0 + i, make_float4(x.x * (1.0f - m.x), x.y * (1.0f - m.y), x.z * (1.0f - m.z), x.w * (1.0f - m.w)));
            __stcs(out1 + i, make_float4(x.x * m.x, x.y * m.y, x.z * m.z, x.w * m.w));
        }
    }
}

/* graph output: voices.audio_out -> audio_out (instance sums in index order) + mix partials */
__global__ void __launch_bounds__(kCta) instance_mix_node_kernel(const __grid_constant__ KernelArgs a, const float *final_stream) {
    __shared__ float tile[kTileF * kTileW];
    __shared__ float isum[kTileF * kCta];
    const uint32_t n = a.n_voices;
    const uint32_t v = blockIdx.x * kCta + threadIdx.x;
    const uint32_t frames = a.ctl->frames;
    const uint32_t col = tile_col(threadIdx.x, a.vpi);
    for (uint32_t f0 = 0; f0 < frames; f0 += kTileF) {
        const uint32_t fe = min(kTileF, frames - f0);
        for (uint32_t ff = 0; ff < fe; ++ff)
            tile[ff * kTileW + col] = (v < n) ? __ldg(final_stream + (size_t)(f0 + ff) * n + v) : 0.0f;
        tile_flush_mono(a, tile, isum, f0, fe);
    }
}

/* value events become the inputs' new per-voice values once every node has run: one thread per
   event; the last event of a (voice, input) in list order wins, as sequential field writes would */
__global__ void commit_value_events_kernel(float *pv, const Event *ev, const BlockCtl *ctl, uint32_t n_voices, uint32_t n_inputs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t ne = ctl->n_events;
    if (i >= ne) return;
    const Event e = ev[i];
    if (e.input_id >= n_inputs) return;
    for (uint32_t j = i + 1; j < ne && ev[j].voice == e.voice; ++j)
        if (ev[j].input_id == e.input_id) return;
    pv[(size_t)e.input_id * n_voices + e.voice] = e.scalar;
}

}  // namespace oscen
