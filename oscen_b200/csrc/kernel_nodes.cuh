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

/* AdsrEnvelope node: pid = {attack, decay, sustain, release}; out0 = output */
__global__ void __launch_bounds__(kCta) adsr_node_kernel(const __grid_constant__ NodeArgs a) {
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
        a.out0[(size_t)f * n + v] = adsr_process(s, prm.P[2], d);
    }
    adsr_store(s, a.state, a.state_base, n, v);
}

/* FmOperator node: in0 = phase_mod (nullable), in1 = envelope; pid = {base_freq, ratio, feedback, level} */
__global__ void __launch_bounds__(kCta) fmop_node_kernel(const __grid_constant__ NodeArgs a) {
    OSC_NODE_PROLOGUE(4)
    FmOpState s;
    s.phase = a.state[(size_t)(a.state_base + 0) * n + v]; s.prev = a.state[(size_t)(a.state_base + 1) * n + v];
    s.inc = fmop_phase_inc(prm.P[0], prm.P[1], a.sr);
    for (uint32_t f0 = 0; f0 < frames; f0 += kPF) {
        float pm[kPF], en[kPF];
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) {
            const uint32_t f = f0 + k;
            pm[k] = (a.in0 != nullptr && f < frames) ? __ldg(a.in0 + (size_t)f * n + v) : 0.0f;
            en[k] = (f < frames) ? __ldg(a.in1 + (size_t)f * n + v) : 0.0f;
        }
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) {
            const uint32_t f = f0 + k;
            if (f < frames) {
                if (ramp_mask) { prm.ramps(a, ramp_mask, f); s.inc = fmop_phase_inc(prm.P[0], prm.P[1], a.sr); }
                if (ec.due(f)) { ec.advance(prm.events(a, ec, f)); s.inc = fmop_phase_inc(prm.P[0], prm.P[1], a.sr); }
                a.out0[(size_t)f * n + v] = fmop_process(s, pm[k], prm.P[2], en[k], prm.P[3]);
            }
        }
    }
    a.state[(size_t)(a.state_base + 0) * n + v] = s.phase; a.state[(size_t)(a.state_base + 1) * n + v] = s.prev;
}

/* TptFilter node: in0 = input, in1 = cutoff stream; pid = {q} */
__global__ void __launch_bounds__(kCta) tpt_node_kernel(const __grid_constant__ NodeArgs a) {
    OSC_NODE_PROLOGUE(1)
    TptState s; tpt_load(s, a.state, a.state_base, n, v);
    for (uint32_t f0 = 0; f0 < frames; f0 += kPF) {
        float x[kPF], c[kPF];
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) {
            const uint32_t f = f0 + k;
            x[k] = (f < frames) ? __ldg(a.in0 + (size_t)f * n + v) : 0.0f;
            c[k] = (f < frames) ? __ldg(a.in1 + (size_t)f * n + v) : 0.0f;
        }
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) {
            const uint32_t f = f0 + k;
            if (f < frames) {
                prm.ramps(a, ramp_mask, f);
                if (ec.due(f)) ec.advance(prm.events(a, ec, f));
                a.out0[(size_t)f * n + v] = tpt_process(s, x[k], c[k], prm.P[0], 0.0f, a.sr);
            }
        }
    }
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
