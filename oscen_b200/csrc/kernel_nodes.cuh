/*
 * kernel_nodes.cuh — the node-granular plan: ONE kernel per node type over
 * [voices x frames_per_block], inter-node streams materialised in HBM as [frame][voice] f32,
 * per-node state SoA in HBM, launched in the plan compiler's topological order (plan.hpp,
 * mirroring oscen-graph-compiler/src/ir/lower.rs:1015-1085) and replayed as a CUDA graph.
 * This is the HBM-bound formulation BASELINE.md's byte model N describes.
 *
 * Stateful nodes are "one thread = one voice, serial in time" (each is a nonlinear / time-varying
 * recurrence; a parallel scan would change the rounding order).  All ready nodes of one type
 * share a launch (blockIdx.y = node) so the serial recurrences of independent nodes overlap.
 * Stream inputs are software-pipelined 8 frames ahead (the loads of chunk k+1 are in flight while
 * chunk k is computed).  Stateless nodes are float4 grid-stride streams unless the block carries
 * value events.
 */
#pragma once
#include "kernels_common.cuh"
#include "kernel_fm.cuh"
#include "kernel_fm_seg.cuh"
#include "kernel_others.cuh"

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
    const float *in0, *in1, *in2;
    float *out0, *out1;
    uint32_t pid[8];                        /* value-input ids feeding this node's parameter slots (kNoInput: use cst): the
                                               node's value endpoints first, then the stand-ins of its stream endpoints, read
                                               when the stream pointer is null — an unconnected endpoint keeps its constructor
                                               value, a Value -> Stream edge broadcasts the value input per frame (plan.hpp) */
    float cst[8];
    float sr;
    uint32_t gate;                          /* AdsrEnvelope: gate events are connected to this node */
    float q;                                /* stateless nodes: second literal (x.clamp(p, q)) */
    /* TMA descriptors (CUtensorMap, 128 B each, in device memory) of the streams behind in0 / in1 / out0, viewed as 2-D
       tensors [kMaxBlock frames][n_voices] with a box of [kTmaF frames][kCta voices]; null when the engine could not
       build them (n_voices % 4 != 0: the row pitch must be a multiple of 16 bytes) */
    const void *tm_in0, *tm_in1, *tm_out0;
};
struct NodeArgs4 { NodeArgs n[4]; };        /* blockIdx.y selects the node */

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
    /* apply this frame's value events; returns the end of the frame's event range */
    OSC_D uint32_t events(const NodeArgs &a, EventCursor &ec, uint32_t f) {
        const uint32_t j = ec.end_of_frame(f);
        for (uint32_t q = ec.i; q < j; ++q) if (ec.ev[q].input_id != kEventGate) value_event(a, ec.ev[q]);
        return j;
    }
};

#define OSC_NODE_PROLOGUE(NP)                                                             \
    const NodeArgs &a = args.n[blockIdx.y];                                               \
    const uint32_t n = a.n_voices;                                                        \
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;                             \
    if (v >= n) return;                                                                   \
    const uint32_t frames = a.ctl->frames, ramp_mask = a.ctl->ramp_mask;                  \
    NodeParams<NP> prm; prm.load(a, v);                                                   \
    EventCursor ec; ec.init(a.events, a.ctl->n_events, v);

/* two-input software pipeline: A/B hold chunk k, A_N/B_N receive chunk k+1 */
#define OSC_LOAD2(A, B, PA, PB, F0)                                                                     \
    _Pragma("unroll") for (uint32_t k = 0; k < kPF; ++k) {                                              \
        const uint32_t f_ = (F0) + k;                                                                   \
        A[k] = ((PA) != nullptr && f_ < frames) ? __ldcs((PA) + (size_t)f_ * n + v) : 0.0f;             \
        B[k] = ((PB) != nullptr && f_ < frames) ? __ldcs((PB) + (size_t)f_ * n + v) : 0.0f;             \
    }
#define OSC_LOAD1(A, PA, F0)                                                                            \
    _Pragma("unroll") for (uint32_t k = 0; k < kPF; ++k) {                                              \
        const uint32_t f_ = (F0) + k;                                                                   \
        A[k] = ((PA) != nullptr && f_ < frames) ? __ldcs((PA) + (size_t)f_ * n + v) : 0.0f;             \
    }

/* ---- AdsrEnvelope: pid = {attack, decay, sustain, release}; out0 = output */
__global__ void __launch_bounds__(kCta) adsr_node_kernel(const __grid_constant__ NodeArgs4 args) {
    OSC_NODE_PROLOGUE(4)
    AdsrState s; adsr_load(s, a.state, a.state_base, n, v);
    AdsrDerived d; adsr_derive_r(d, prm.P[0], prm.P[1], prm.P[3], a.sr);
    adsr_begin_block(s, d, ec.due(0u));
    AdsrHot h; adsr_refresh<false>(s, h, prm.P[2], d);   /* re-run after every event / ramped frame */
    for (uint32_t f = 0; f < frames; ++f) {
        const bool due = ec.due(f);
        if (ramp_mask | (uint32_t)due) {
            prm.ramps(a, ramp_mask, f);
            uint32_t j = ec.i;
            if (due) j = prm.events(a, ec, f);
            if (adsr_derived_stale(d, prm.P[0], prm.P[1], prm.P[3])) adsr_derive_r(d, prm.P[0], prm.P[1], prm.P[3], a.sr);
            if (due) {
                for (uint32_t q = ec.i; q < j; ++q) if (ec.ev[q].input_id == kEventGate && a.gate)
                    adsr_on_gate(s, ec.ev[q].scalar, prm.P[0], prm.P[3], d);
                ec.advance(j);
            }
            adsr_refresh(s, h, prm.P[2], d);
        }
        __stcs(a.out0 + (size_t)f * n + v, adsr_process_t<false>(s, h, prm.P[2], d));
    }
    adsr_store(s, a.state, a.state_base, n, v);
}

/* ---- FmOperator: in0 = phase_mod, in1 = envelope (either nullable: then slot 4 / 5 stands in);
        pid = {base_freq, ratio, feedback, level, phase_mod, envelope} */
__global__ void __launch_bounds__(kCta) fmop_node_kernel(const __grid_constant__ NodeArgs4 args) {
    OSC_NODE_PROLOGUE(6)
    const bool pm_stream = a.in0 != nullptr, en_stream = a.in1 != nullptr;
    FmOpState s;
    s.phase = a.state[(size_t)(a.state_base + 0) * n + v]; s.prev = a.state[(size_t)(a.state_base + 1) * n + v];
    s.inc = fmop_phase_inc(prm.P[0], prm.P[1], a.sr);
    const OmSinRegs sk = om_sin_regs();
    float pm[kPF], en[kPF], pm_n[kPF], en_n[kPF];
    OSC_LOAD2(pm, en, a.in0, a.in1, 0u)
    for (uint32_t f0 = 0; f0 < frames; f0 += kPF) {
        OSC_LOAD2(pm_n, en_n, a.in0, a.in1, f0 + kPF)
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) {
            const uint32_t f = f0 + k;
            if (f < frames) {
                if (ramp_mask) { prm.ramps(a, ramp_mask, f); s.inc = fmop_phase_inc(prm.P[0], prm.P[1], a.sr); }
                if (ec.due(f)) { ec.advance(prm.events(a, ec, f)); s.inc = fmop_phase_inc(prm.P[0], prm.P[1], a.sr); }
                __stcs(a.out0 + (size_t)f * n + v, fmop_process_k(s, pm_stream ? pm[k] : prm.P[4], prm.P[2], en_stream ? en[k] : prm.P[5], prm.P[3], sk));
            }
        }
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) { pm[k] = pm_n[k]; en[k] = en_n[k]; }
    }
    a.state[(size_t)(a.state_base + 0) * n + v] = s.phase; a.state[(size_t)(a.state_base + 1) * n + v] = s.prev;
}

/* ---- TptFilter: in0 = input, in1 = cutoff, in2 = f_mod (each nullable: then the slot stands in, e.g. a
        value input feeding the `cutoff` stream endpoint is a broadcast per frame); pid = {q, cutoff, f_mod, input} */
__global__ void __launch_bounds__(kCta) tpt_node_kernel(const __grid_constant__ NodeArgs4 args) {
    OSC_NODE_PROLOGUE(4)
    TptState s; tpt_load(s, a.state, a.state_base, n, v);
    const bool in_stream = a.in0 != nullptr, cut_stream = a.in1 != nullptr, fm_stream = a.in2 != nullptr;
    float x[kPF], c[kPF], x_n[kPF], c_n[kPF];
    OSC_LOAD2(x, c, a.in0, a.in1, 0u)
    if (!fm_stream) {
        for (uint32_t f0 = 0; f0 < frames; f0 += kPF) {
            OSC_LOAD2(x_n, c_n, a.in0, a.in1, f0 + kPF)
#pragma unroll
            for (uint32_t k = 0; k < kPF; ++k) {
                const uint32_t f = f0 + k;
                if (f < frames) {
                    prm.ramps(a, ramp_mask, f);
                    if (ec.due(f)) ec.advance(prm.events(a, ec, f));
                    __stcs(a.out0 + (size_t)f * n + v, tpt_process(s, in_stream ? x[k] : prm.P[3], cut_stream ? c[k] : prm.P[1], prm.P[0], prm.P[2], a.sr));
                }
            }
#pragma unroll
            for (uint32_t k = 0; k < kPF; ++k) { x[k] = x_n[k]; c[k] = c_n[k]; }
        }
    } else {
        float m[kPF], m_n[kPF];
        OSC_LOAD1(m, a.in2, 0u)
        for (uint32_t f0 = 0; f0 < frames; f0 += kPF) {
            OSC_LOAD2(x_n, c_n, a.in0, a.in1, f0 + kPF)
            OSC_LOAD1(m_n, a.in2, f0 + kPF)
#pragma unroll
            for (uint32_t k = 0; k < kPF; ++k) {
                const uint32_t f = f0 + k;
                if (f < frames) {
                    prm.ramps(a, ramp_mask, f);
                    if (ec.due(f)) ec.advance(prm.events(a, ec, f));
                    __stcs(a.out0 + (size_t)f * n + v, tpt_process(s, in_stream ? x[k] : prm.P[3], cut_stream ? c[k] : prm.P[1], prm.P[0], m[k], a.sr));
                }
            }
#pragma unroll
            for (uint32_t k = 0; k < kPF; ++k) { x[k] = x_n[k]; c[k] = c_n[k]; m[k] = m_n[k]; }
        }
    }
    tpt_store(s, a.state, a.state_base, n, v);
}

/* ---- PolyBlepOscillator: in0 = frequency_mod, in1 = phase_mod (nullable: slots 3 / 4 stand in);
        pid = {frequency, amplitude, pulse_width, frequency_mod, phase_mod} */
template <int WAVE>
__global__ void __launch_bounds__(kCta) polyblep_node_kernel(const __grid_constant__ NodeArgs4 args) {
    OSC_NODE_PROLOGUE(5)
    float phase = a.state[(size_t)a.state_base * n + v];
    const bool fm_stream = a.in0 != nullptr, pm_stream = a.in1 != nullptr;
    float fm[kPF], fm_n[kPF], pm[kPF], pm_n[kPF];
    OSC_LOAD2(fm, pm, a.in0, a.in1, 0u)
    for (uint32_t f0 = 0; f0 < frames; f0 += kPF) {
        OSC_LOAD2(fm_n, pm_n, a.in0, a.in1, f0 + kPF)
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) {
            const uint32_t f = f0 + k;
            if (f < frames) {
                prm.ramps(a, ramp_mask, f);
                if (ec.due(f)) ec.advance(prm.events(a, ec, f));
                __stcs(a.out0 + (size_t)f * n + v, polyblep_process<WAVE>(phase, pm_stream ? pm[k] : prm.P[4], prm.P[0], fm_stream ? fm[k] : prm.P[3],
                                                                         prm.P[1], prm.P[2], a.sr));
            }
        }
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) { fm[k] = fm_n[k]; pm[k] = pm_n[k]; }
    }
    a.state[(size_t)a.state_base * n + v] = phase;
}

/* ---- LP18Filter: in0 = input (nullable: slot 3 stands in); pid = {cutoff, resonance, fmod, input} */
__global__ void __launch_bounds__(kCta) lp18_node_kernel(const __grid_constant__ NodeArgs4 args) {
    OSC_NODE_PROLOGUE(4)
    const bool in_stream = a.in0 != nullptr;
    Lp18State s; lp18_load(s, a.state, a.state_base, n, v);
    float x[kPF], x_n[kPF];
    [[maybe_unused]] float dummy[kPF], dummy_n[kPF];
    const float *no_stream = nullptr;
    OSC_LOAD2(x, dummy, a.in0, no_stream, 0u)
    for (uint32_t f0 = 0; f0 < frames; f0 += kPF) {
        OSC_LOAD2(x_n, dummy_n, a.in0, no_stream, f0 + kPF)
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) {
            const uint32_t f = f0 + k;
            if (f < frames) {
                prm.ramps(a, ramp_mask, f);
                if (ec.due(f)) ec.advance(prm.events(a, ec, f));
                __stcs(a.out0 + (size_t)f * n + v, lp18_process(s, in_stream ? x[k] : prm.P[3], prm.P[0], prm.P[2], prm.P[1], a.sr));
            }
        }
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) x[k] = x_n[k];
    }
    lp18_store(s, a.state, a.state_base, n, v);
}

/* ---- naive Oscillator: in0 = frequency_mod (nullable: slot 2 stands in); pid = {frequency, amplitude, frequency_mod} */
template <int WAVE>
__global__ void __launch_bounds__(kCta) osc_node_kernel(const __grid_constant__ NodeArgs4 args) {
    OSC_NODE_PROLOGUE(3)
    float phase = a.state[(size_t)a.state_base * n + v];
    const bool fm_stream = a.in0 != nullptr;
    float fm[kPF], fm_n[kPF];
    OSC_LOAD1(fm, a.in0, 0u)
    for (uint32_t f0 = 0; f0 < frames; f0 += kPF) {
        OSC_LOAD1(fm_n, a.in0, f0 + kPF)
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) {
            const uint32_t f = f0 + k;
            if (f < frames) {
                prm.ramps(a, ramp_mask, f);
                if (ec.due(f)) ec.advance(prm.events(a, ec, f));
                __stcs(a.out0 + (size_t)f * n + v, osc_process<WAVE>(phase, prm.P[0], fm_stream ? fm[k] : prm.P[2], prm.P[1], a.sr));
            }
        }
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) fm[k] = fm_n[k];
    }
    a.state[(size_t)a.state_base * n + v] = phase;
}

/* ---- IirLowpass: in0 = input (nullable: slot 2 stands in); pid = {cutoff, q, input} */
__global__ void __launch_bounds__(kCta) iirlp_node_kernel(const __grid_constant__ NodeArgs4 args) {
    OSC_NODE_PROLOGUE(3)
    IirLpState s;
    {
        const float *st = a.state + (size_t)a.state_base * n + v;
        s.b0 = st[0]; s.b1 = st[n]; s.b2 = st[2 * (size_t)n]; s.a1 = st[3 * (size_t)n]; s.a2 = st[4 * (size_t)n];
        s.v1 = st[5 * (size_t)n]; s.v2 = st[6 * (size_t)n]; s.counter = __float_as_uint(st[7 * (size_t)n]);
    }
    const bool in_stream = a.in0 != nullptr;
    float x[kPF], x_n[kPF];
    OSC_LOAD1(x, a.in0, 0u)
    for (uint32_t f0 = 0; f0 < frames; f0 += kPF) {
        OSC_LOAD1(x_n, a.in0, f0 + kPF)
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) {
            const uint32_t f = f0 + k;
            if (f < frames) {
                prm.ramps(a, ramp_mask, f);
                if (ec.due(f)) ec.advance(prm.events(a, ec, f));
                __stcs(a.out0 + (size_t)f * n + v, iirlp_process(s, in_stream ? x[k] : prm.P[2], prm.P[0], prm.P[1], a.sr));
            }
        }
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) x[k] = x_n[k];
    }
    {
        float *st = a.state + (size_t)a.state_base * n + v;
        st[0] = s.b0; st[n] = s.b1; st[2 * (size_t)n] = s.b2; st[3 * (size_t)n] = s.a1; st[4 * (size_t)n] = s.a2;
        st[5 * (size_t)n] = s.v1; st[6 * (size_t)n] = s.v2; st[7 * (size_t)n] = __uint_as_float(s.counter);
    }
}

/* ---- Delay (no feedback edge through it: a cycle needs the per-sample walk of a generated kernel):
        in0 = input (nullable: slot 2 stands in); pid = {delay_samples, feedback, input};
        ring capacity in NodeArgs.gate (a power of two), the reference capacity - 1 in NodeArgs.q */
__global__ void __launch_bounds__(kCta) delay_node_kernel(const __grid_constant__ NodeArgs4 args) {
    OSC_NODE_PROLOGUE(3)
    DelayState s;
    s.write_pos = __float_as_uint(a.state[(size_t)(a.state_base + 0u) * n + v]);
    s.counter = __float_as_uint(a.state[(size_t)(a.state_base + 1u) * n + v]);
    s.output = a.state[(size_t)(a.state_base + 2u) * n + v];
    float *ring = a.state + (size_t)(a.state_base + 3u) * n + v;
    const uint32_t cap = a.gate;
    const float max_delay = a.q;
    const bool in_stream = a.in0 != nullptr, cd = a.pid[0] != kNoInput, cf = a.pid[1] != kNoInput;
    for (uint32_t f = 0; f < frames; ++f) {
        prm.ramps(a, ramp_mask, f);
        if (ec.due(f)) ec.advance(prm.events(a, ec, f));
        const float x = in_stream ? __ldcs(a.in0 + (size_t)f * n + v) : prm.P[2];
        float y;
        if (cd) y = cf ? delay_process<true, true>(s, ring, n, cap, max_delay, x, prm.P[0], prm.P[1])
                       : delay_process<true, false>(s, ring, n, cap, max_delay, x, prm.P[0], prm.P[1]);
        else y = cf ? delay_process<false, true>(s, ring, n, cap, max_delay, x, prm.P[0], prm.P[1])
                    : delay_process<false, false>(s, ring, n, cap, max_delay, x, prm.P[0], prm.P[1]);
        __stcs(a.out0 + (size_t)f * n + v, y);
    }
    a.state[(size_t)(a.state_base + 0u) * n + v] = __uint_as_float(s.write_pos);
    a.state[(size_t)(a.state_base + 1u) * n + v] = __uint_as_float(s.counter);
    a.state[(size_t)(a.state_base + 2u) * n + v] = s.output;
}

/* ---- SamplePlayer (mono): no stream or value inputs; the slot table travels in NodeArgs.in1 (a `const AssetTable *`),
        the slot index in NodeArgs.gate (0xFFFFFFFF: `buf` is not connected — silence); state = the playhead */
__global__ void __launch_bounds__(kCta) sample_player_node_kernel(const __grid_constant__ NodeArgs4 args) {
    const NodeArgs &a = args.n[blockIdx.y];
    const uint32_t n = a.n_voices;
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint32_t frames = a.ctl->frames;
    const AssetTable *tab = reinterpret_cast<const AssetTable *>(a.in1);
    const bool bound = a.gate != 0xFFFFFFFFu && tab != nullptr;
    const float *buf = bound ? tab->data[a.gate] : nullptr;
    const uint32_t len = bound ? tab->frames[a.gate] : 0u;
    uint32_t playhead = (bound && tab->reset[a.gate]) ? 0u : __float_as_uint(a.state[(size_t)a.state_base * n + v]);
    if (playhead >= len) playhead = 0u;
    for (uint32_t f = 0; f < frames; ++f) __stcs(a.out0 + (size_t)f * n + v, sample_player_process(playhead, buf, len));
    a.state[(size_t)a.state_base * n + v] = __uint_as_float(playhead);
}

/* =====================================================================================================================
 * PLAIN variants: the block carries no event of any kind and no ramp (the host knows both before it launches, and a block
 * of a held chord is exactly that), so every value input is a constant of the block and the frame loop needs no event
 * cursor, no ramp test and no re-derivation.  The bodies are the lean forms of kernel_fm_seg.cuh (bounded-argument sine
 * behind one range test, x - trunc(x) phase wrap, warp-uniform filter refresh, per-warp envelope segments); stream
 * addressing is 32-bit (n_voices * 512 < 2^32 is checked at create).  Same operations per frame as the generic kernels
 * above, so the same bits (tests/test_dsl_gpu.py node-plan cases run both).
 * Stream inputs are register-prefetched two tiles deep (tile A is computed while tile B's loads are in flight), with the
 * two tiles in alternating roles instead of a copy.
 * ===================================================================================================================== */
#define OSC_PLAIN_PROLOGUE(NP)                                                            \
    const NodeArgs &a = args.n[blockIdx.y];                                               \
    const uint32_t n = a.n_voices;                                                        \
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;                             \
    if (v >= n) return;                                                                   \
    const uint32_t frames = a.ctl->frames;                                                \
    NodeParams<NP> prm; prm.load(a, v);
/* tile of kPF frames of one stream into registers (zeros past the block) */
#define OSC_PLAIN_LOAD(A, PA, F0)                                                                       \
    _Pragma("unroll") for (uint32_t k = 0; k < kPF; ++k) {                                              \
        const uint32_t f_ = (F0) + k;                                                                   \
        A[k] = ((PA) != nullptr && f_ < frames) ? __ldcs((PA) + (f_ * n + v)) : 0.0f;                   \
    }

/* ---- FmOperator, plain block */
__global__ void __launch_bounds__(kCta) fmop_node_plain_kernel(const __grid_constant__ NodeArgs4 args) {
    OSC_PLAIN_PROLOGUE(6)
    const bool pm_stream = a.in0 != nullptr, en_stream = a.in1 != nullptr;
    FmOpState s;
    s.phase = a.state[(size_t)(a.state_base + 0) * n + v]; s.prev = a.state[(size_t)(a.state_base + 1) * n + v];
    s.inc = fmop_phase_inc(prm.P[0], prm.P[1], a.sr);
    const OmSinRegs sk = om_sin_regs();
    const float fb = prm.P[2], lvl = prm.P[3], pm_c = prm.P[4], en_c = prm.P[5];
    float pmA[kPF], enA[kPF], pmB[kPF], enB[kPF];
#define OSC_FMOP_TILE(PM, EN, F0)                                                                       \
    _Pragma("unroll") for (uint32_t k = 0; k < kPF; ++k) {                                              \
        const uint32_t f = (F0) + k;                                                                    \
        if (f < frames) __stcs(a.out0 + (f * n + v), fmop_process_k(s, pm_stream ? PM[k] : pm_c, fb, en_stream ? EN[k] : en_c, lvl, sk)); \
    }
    OSC_PLAIN_LOAD(pmA, a.in0, 0u) OSC_PLAIN_LOAD(enA, a.in1, 0u)
    for (uint32_t f0 = 0; f0 < frames; f0 += 2u * kPF) {
        OSC_PLAIN_LOAD(pmB, a.in0, f0 + kPF) OSC_PLAIN_LOAD(enB, a.in1, f0 + kPF)
        OSC_FMOP_TILE(pmA, enA, f0)
        OSC_PLAIN_LOAD(pmA, a.in0, f0 + 2u * kPF) OSC_PLAIN_LOAD(enA, a.in1, f0 + 2u * kPF)
        OSC_FMOP_TILE(pmB, enB, f0 + kPF)
    }
#undef OSC_FMOP_TILE
    a.state[(size_t)(a.state_base + 0) * n + v] = s.phase; a.state[(size_t)(a.state_base + 1) * n + v] = s.prev;
}

/* ---- TptFilter, plain block.  ZERO_FMOD: no f_mod stream and the f_mod value is 0 — then the cutoff path is
   clamp -> refresh test -> core, with the refresh warp-uniform (tpt_process_mod); otherwise the general tpt_process */
template <bool ZERO_FMOD>
__global__ void __launch_bounds__(kCta) tpt_node_plain_kernel(const __grid_constant__ NodeArgs4 args) {
    const NodeArgs &a = args.n[blockIdx.y];
    const uint32_t n = a.n_voices;
    const uint32_t vt = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = vt < n;
    const uint32_t v = active ? vt : 0u;                     /* the whole warp stays for the votes of tpt_process_mod */
    const uint32_t frames = a.ctl->frames;
    NodeParams<4> prm; prm.load(a, v);
    TptState s; tpt_load(s, a.state, a.state_base, n, v);
    const bool in_stream = a.in0 != nullptr, cut_stream = a.in1 != nullptr, fm_stream = a.in2 != nullptr;
    const float sr = a.sr;
    const float max_cutoff = fminf(sr * 0.5f - FLT_EPSILON, 20000.0f), period = 0.5f / sr, two_sr = 2.0f * sr;
    const float qc = clampf(prm.P[0], 0.1f, 10.0f), inv_q = 1.0f / qc;
    const float cut_c = prm.P[1], fm_c = prm.P[2], in_c = prm.P[3];
    float xA[kPF], cA[kPF], xB[kPF], cB[kPF];
    [[maybe_unused]] float mA[kPF], mB[kPF];
#define OSC_TPT_TILE(X, C, M, F0)                                                                       \
    _Pragma("unroll") for (uint32_t k = 0; k < kPF; ++k) {                                              \
        const uint32_t f = (F0) + k;                                                                    \
        if (f < frames) {                                                                               \
            const float x_ = in_stream ? X[k] : in_c, c_ = cut_stream ? C[k] : cut_c;                   \
            float y_;                                                                                   \
            if (ZERO_FMOD) y_ = tpt_process_mod(s, x_, c_, qc, inv_q, max_cutoff, sr, two_sr, period);  \
            else y_ = tpt_process(s, x_, c_, prm.P[0], fm_stream ? M[k] : fm_c, sr);                    \
            if (active) __stcs(a.out0 + (f * n + v), y_);                                               \
        }                                                                                               \
    }
    const float *in2 = ZERO_FMOD ? nullptr : a.in2;
    OSC_PLAIN_LOAD(xA, a.in0, 0u) OSC_PLAIN_LOAD(cA, a.in1, 0u)
    if (!ZERO_FMOD) { OSC_PLAIN_LOAD(mA, in2, 0u) }
    for (uint32_t f0 = 0; f0 < frames; f0 += 2u * kPF) {
        OSC_PLAIN_LOAD(xB, a.in0, f0 + kPF) OSC_PLAIN_LOAD(cB, a.in1, f0 + kPF)
        if (!ZERO_FMOD) { OSC_PLAIN_LOAD(mB, in2, f0 + kPF) }
        OSC_TPT_TILE(xA, cA, mA, f0)
        OSC_PLAIN_LOAD(xA, a.in0, f0 + 2u * kPF) OSC_PLAIN_LOAD(cA, a.in1, f0 + 2u * kPF)
        if (!ZERO_FMOD) { OSC_PLAIN_LOAD(mA, in2, f0 + 2u * kPF) }
        OSC_TPT_TILE(xB, cB, mB, f0 + kPF)
    }
#undef OSC_TPT_TILE
    if (active) tpt_store(s, a.state, a.state_base, n, v);
}

/* =====================================================================================================================
 * TMA-TILED variants of the plain FmOperator / TptFilter kernels (north_star: "staged through shared memory ... TMA").
 * A CTA's 64 voices x kTmaF frames of every stream is one 2-D box: ONE elected thread issues
 * `cp.async.bulk.tensor.2d` loads (SASS UTMALDG) into a 3-stage shared-memory ring, completion on an mbarrier per stage;
 * the recurrence reads its inputs with LDS, writes its output tile with STS, and the same thread streams finished tiles out
 * with `cp.async.bulk.tensor.2d` stores (UTMASTG, bulk async-group, double-buffered).  Against the register-prefetch
 * kernels above this takes the stream address arithmetic, the bounds predicates and 32 prefetch registers per thread out of
 * the serial loop (fmop: 96 -> ~60 registers, so 16 instead of 10 CTAs per SM hide the sine's FP64 / conversion latency).
 * Voices past n_voices in the last CTA are zero-filled on load and clipped on store by the descriptor; frames past the
 * block inside the last box are rows of the stream buffer nobody reads.
 * ===================================================================================================================== */
constexpr uint32_t kTmaF = 8;                 /* frames per box */
constexpr uint32_t kTmaStages = 3;
constexpr uint32_t kTmaTile = kTmaF * kCta;   /* floats per box */
OSC_D void tma_load_2d(void *smem_dst, const void *tmap, uint32_t x, uint32_t y, unsigned long long *bar) {
    OSC_PTX("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
OSC_D void tma_store_2d(const void *tmap, uint32_t x, uint32_t y, const void *smem_src) {
    OSC_PTX("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(tmap), "r"(x), "r"(y), "r"(smem_u32(smem_src)) : "memory");
}
OSC_D void tma_store_commit() { OSC_PTX("cp.async.bulk.commit_group;" ::: "memory"); }
OSC_D void tma_store_wait_read0() { OSC_PTX("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
OSC_D void tma_store_wait_all() { OSC_PTX("cp.async.bulk.wait_group 0;" ::: "memory"); }

/* the tile pipeline shared by the two kernels: BODY(k, x0, x1) -> output sample of frame tile*kTmaF + k */
#define OSC_TMA_PIPELINE(HAS0, HAS1, BODY)                                                                          \
    __shared__ __align__(128) float in_ring[kTmaStages][2][kTmaTile];                                               \
    __shared__ __align__(128) float out_ring[2][kTmaTile];                                                          \
    __shared__ __align__(8) unsigned long long full_bar[kTmaStages];                                                \
    const uint32_t v0 = blockIdx.x * kCta, tid = threadIdx.x;                                                       \
    const uint32_t n_tiles = (frames + kTmaF - 1u) / kTmaF;                                                         \
    const uint32_t tx_bytes = ((HAS0) ? kTmaTile * 4u : 0u) + ((HAS1) ? kTmaTile * 4u : 0u);                        \
    if (tid == 0) { for (uint32_t st = 0; st < kTmaStages; ++st) mbar_init(&full_bar[st], 1u); }                    \
    __syncthreads();                                                                                                \
    auto issue = [&](uint32_t tile) {                                                                               \
        const uint32_t st = tile % kTmaStages;                                                                      \
        if (tx_bytes) {                                                                                             \
            mbar_expect_tx(&full_bar[st], tx_bytes);                                                                \
            if (HAS0) tma_load_2d(in_ring[st][0], a.tm_in0, v0, tile * kTmaF, &full_bar[st]);                       \
            if (HAS1) tma_load_2d(in_ring[st][1], a.tm_in1, v0, tile * kTmaF, &full_bar[st]);                       \
        }                                                                                                           \
    };                                                                                                              \
    if (tid == 0) for (uint32_t t0 = 0; t0 + 1u < kTmaStages && t0 < n_tiles; ++t0) issue(t0);                      \
    for (uint32_t tile = 0; tile < n_tiles; ++tile) {                                                               \
        const uint32_t st = tile % kTmaStages;                                                                      \
        if (tid == 0 && tile + kTmaStages - 1u < n_tiles) issue(tile + kTmaStages - 1u);   /* stage of tile - 1: consumed */ \
        if (tx_bytes) mbar_wait(&full_bar[st], (tile / kTmaStages) & 1u);                                           \
        const float *x0p = in_ring[st][0] + tid, *x1p = in_ring[st][1] + tid;                                       \
        float *op = out_ring[tile & 1u] + tid;                                                                      \
        const uint32_t fe = min(kTmaF, frames - tile * kTmaF);                                                      \
        _Pragma("unroll") for (uint32_t k = 0; k < kTmaF; ++k) {                                                    \
            if (k < fe) {                                                                                           \
                const float x0 = (HAS0) ? x0p[k * kCta] : 0.0f, x1 = (HAS1) ? x1p[k * kCta] : 0.0f;                 \
                op[k * kCta] = BODY(x0, x1);                                                                        \
            }                                                                                                       \
        }                                                                                                           \
        fence_async_smem();                               /* this thread's STS before the async proxy reads them */ \
        if (tid == 0) tma_store_wait_read0();             /* the store of tile - 1 has read out_ring[(tile + 1) & 1] */ \
        __syncthreads();                                                                                            \
        if (tid == 0) { tma_store_2d(a.tm_out0, v0, tile * kTmaF, out_ring[tile & 1u]); tma_store_commit(); }       \
    }                                                                                                               \
    if (tid == 0) tma_store_wait_all();

/* ---- FmOperator, plain block, TMA tiles */
__global__ void __launch_bounds__(kCta) fmop_node_tma_kernel(const __grid_constant__ NodeArgs4 args) {
    const NodeArgs &a = args.n[blockIdx.y];
    const uint32_t n = a.n_voices;
    const uint32_t vt = blockIdx.x * kCta + threadIdx.x;
    const bool active = vt < n;
    const uint32_t v = active ? vt : 0u;                     /* every thread of the CTA stays for the barriers */
    const uint32_t frames = a.ctl->frames;
    NodeParams<6> prm; prm.load(a, v);
    const bool pm_stream = a.in0 != nullptr, en_stream = a.in1 != nullptr;
    FmOpState s;
    s.phase = a.state[(size_t)(a.state_base + 0) * n + v]; s.prev = a.state[(size_t)(a.state_base + 1) * n + v];
    s.inc = fmop_phase_inc(prm.P[0], prm.P[1], a.sr);
    const OmSinRegs sk = om_sin_regs();
    const float fb = prm.P[2], lvl = prm.P[3], pm_c = prm.P[4], en_c = prm.P[5];
#define OSC_FMOP_BODY(X0, X1) fmop_process_k(s, pm_stream ? (X0) : pm_c, fb, en_stream ? (X1) : en_c, lvl, sk)
    OSC_TMA_PIPELINE(pm_stream, en_stream, OSC_FMOP_BODY)
#undef OSC_FMOP_BODY
    if (active) { a.state[(size_t)(a.state_base + 0) * n + v] = s.phase; a.state[(size_t)(a.state_base + 1) * n + v] = s.prev; }
}

/* ---- TptFilter (no f_mod), plain block, TMA tiles */
__global__ void __launch_bounds__(kCta) tpt_node_tma_kernel(const __grid_constant__ NodeArgs4 args) {
    const NodeArgs &a = args.n[blockIdx.y];
    const uint32_t n = a.n_voices;
    const uint32_t vt = blockIdx.x * kCta + threadIdx.x;
    const bool active = vt < n;
    const uint32_t v = active ? vt : 0u;
    const uint32_t frames = a.ctl->frames;
    NodeParams<4> prm; prm.load(a, v);
    TptState s; tpt_load(s, a.state, a.state_base, n, v);
    const bool in_stream = a.in0 != nullptr, cut_stream = a.in1 != nullptr;
    const float sr = a.sr;
    const float max_cutoff = fminf(sr * 0.5f - FLT_EPSILON, 20000.0f), period = 0.5f / sr, two_sr = 2.0f * sr;
    const float qc = clampf(prm.P[0], 0.1f, 10.0f), inv_q = 1.0f / qc;
    const float cut_c = prm.P[1], in_c = prm.P[3];
#define OSC_TPT_BODY(X0, X1) tpt_process_mod(s, in_stream ? (X0) : in_c, cut_stream ? (X1) : cut_c, qc, inv_q, max_cutoff, sr, two_sr, period)
    OSC_TMA_PIPELINE(in_stream, cut_stream, OSC_TPT_BODY)
#undef OSC_TPT_BODY
    if (active) tpt_store(s, a.state, a.state_base, n, v);
}

/* ---- AdsrEnvelope, plain block: no gate can arrive, so a stage only ends by running out.  Per warp: L = frames until the
   first lane's timed stage completes (one REDUX.MIN); L plain frames — Attack / Decay lanes step their one-pole, Sustain /
   Idle lanes hold a constant, Release lanes keep their per-frame division behind a warp-uniform test — then ONE general
   frame (adsr_process_t) takes the stage transition.  A warp whose voices all sit in Sustain / Idle only stores. */
__global__ void __launch_bounds__(kCta) adsr_node_plain_kernel(const __grid_constant__ NodeArgs4 args) {
    const NodeArgs &a = args.n[blockIdx.y];
    const uint32_t n = a.n_voices;
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = v < n;
    const uint32_t vs = active ? v : 0u;                     /* the whole warp stays for the reductions */
    const uint32_t frames = a.ctl->frames;
    NodeParams<4> prm; prm.load(a, vs);
    AdsrState s; adsr_load(s, a.state, a.state_base, n, vs);
    AdsrDerived d; adsr_derive_r(d, prm.P[0], prm.P[1], prm.P[3], a.sr);
    adsr_begin_block(s, d, false);
    AdsrHot h; adsr_refresh<false>(s, h, prm.P[2], d);
    const float sus = prm.P[2];
    const uint32_t kBig = 0x7FFFFFFFu;
    uint32_t idx = vs;                                       /* f * n + v */
    uint32_t f = 0;
    while (f < frames) {
        const bool timed = (s.stage == ST_ATTACK) | (s.stage == ST_DECAY) | (s.stage == ST_RELEASE);
        uint32_t np = timed ? (s.rem >= 2u ? s.rem - 1u : 0u) : kBig;
        if (f == 0u) np = 0u;                                /* the block's first frame commits the caches (adsr_refresh<true>) */
        const uint32_t L = umin(__reduce_min_sync(0xffffffffu, np), frames - f);
        if (L > 0u) {
            const bool ad = (s.stage == ST_ATTACK) | (s.stage == ST_DECAY), rl = s.stage == ST_RELEASE;
            const float tg = s.stage == ST_ATTACK ? 1.0f : h.sus, kk = s.stage == ST_ATTACK ? d.a_k : d.d_k;
            if (!__any_sync(0xffffffffu, rl)) {
                if (!__any_sync(0xffffffffu, ad)) {
                    const float y = s.level;
                    for (uint32_t i = 0; i < L; ++i) { if (active) __stcs(a.out0 + idx, y); idx += n; }
                } else {
                    for (uint32_t i = 0; i < L; ++i) { const float y = adsr_ad(s, ad, tg, kk); if (active) __stcs(a.out0 + idx, y); idx += n; }
                    if (ad) s.rem -= L;
                }
            } else {
                for (uint32_t i = 0; i < L; ++i) { const float y = adsr_plain_rel(s, ad, rl, tg, kk); if (active) __stcs(a.out0 + idx, y); idx += n; }
            }
            f += L;
            continue;
        }
        if (f == 0u) adsr_refresh<true>(s, h, sus, d);
        const float y = adsr_process_t<false>(s, h, sus, d);
        if (active) __stcs(a.out0 + idx, y);
        idx += n; f += 1u;
    }
    if (active) adsr_store(s, a.state, a.state_base, n, v);
}

/* ---- stateless nodes */
enum : int {
    EW_GAIN = 0,        /* Gain: in0 * (value | constant)        gain/mod.rs:30-35      */
    EW_ADD_VALUE = 1,   /* AddValue: in0 + value                 add_value.rs:33-35     */
    EW_MIXER = 2,       /* Mixer: in0 + in1                      mixer.rs:32-34         */
    EW_CROSSFADE = 3,   /* Crossfade: in0*(1-mix), in0*mix       crossfade.rs:39-44     */
    EW_MUL = 4,         /* Gain with a stream on `gain`: in0 * in1                       */
    EW_SUB = 5,         /* expression edge a - b                 codegen/mod.rs:433-472 */
    EW_HARDCLIP = 6,    /* HardClip                              oversampled-saturator main.rs:55-62 */
    /* the remaining operators of compound connection sources (plan::ExprOp; the reference emits them as
       Rust expressions evaluated per frame, codegen/mod.rs:433-472): one f32 rounding each, like the Rust */
    EW_DIV = 7,         /* a / b */
    EW_SUBV = 8,        /* a - v */
    EW_RSUBV = 9,       /* v - a */
    EW_DIVV = 10,       /* a / v */
    EW_RDIVV = 11,      /* v / a */
    EW_VALUE = 12,      /* a value input or literal read as a stream (Value -> Stream edge) */
    EW_TANH = 13,       /* f32::tanh */
    EW_ABS = 14,        /* f32::abs */
    EW_CLAMP = 15,      /* f32::clamp(p, q) */
    EW_MINV = 16,       /* f32::min(p): minNum */
    EW_MAXV = 17,       /* f32::max(p): maxNum */
    EW_SQRT = 18,       /* f32::sqrt (IEEE) */
    EW_SIN = 19, EW_COS = 20, EW_EXP = 21,   /* libm sinf / cosf / expf (om_math.h) */
};
template <int OP>
OSC_D void ew_apply(float x, float y, float p, float q, float &o0, float &o1) {
    o1 = 0.0f;
    if (OP == EW_GAIN) o0 = x * p;
    else if (OP == EW_ADD_VALUE) o0 = x + p;
    else if (OP == EW_MIXER) o0 = x + y;
    else if (OP == EW_MUL) o0 = x * y;
    else if (OP == EW_SUB) o0 = x - y;
    else if (OP == EW_HARDCLIP) o0 = hardclip_process(x);
    else if (OP == EW_DIV) o0 = x / y;
    else if (OP == EW_SUBV) o0 = x - p;
    else if (OP == EW_RSUBV) o0 = p - x;
    else if (OP == EW_DIVV) o0 = x / p;
    else if (OP == EW_RDIVV) o0 = p / x;
    else if (OP == EW_VALUE) o0 = p;
    else if (OP == EW_TANH) o0 = om_tanhf_u(x);
    else if (OP == EW_ABS) o0 = fabsf(x);
    else if (OP == EW_CLAMP) o0 = clampf(x, p, q);
    else if (OP == EW_MINV) o0 = fminf(x, p);
    else if (OP == EW_MAXV) o0 = fmaxf(x, p);
    else if (OP == EW_SQRT) o0 = sqrtf(x);
    else if (OP == EW_SIN) o0 = om_sinf(x);
    else if (OP == EW_COS) o0 = om_cosf(x);
    else if (OP == EW_EXP) o0 = om_expf(x);
    else { const float mix = clampf(p, 0.0f, 1.0f); o0 = x * (1.0f - mix); o1 = x * mix; }
}
template <int OP> OSC_HD constexpr bool ew_two_inputs() { return OP == EW_MIXER || OP == EW_MUL || OP == EW_SUB || OP == EW_DIV; }
template <int OP> OSC_HD constexpr bool ew_no_inputs() { return OP == EW_VALUE; }

/* serial per-voice form (used when the block carries value events) */
template <int OP>
__global__ void __launch_bounds__(kCta) ew_node_kernel(const __grid_constant__ NodeArgs4 args) {
    OSC_NODE_PROLOGUE(1)
    for (uint32_t f0 = 0; f0 < frames; f0 += kPF) {
        float x[kPF], y[kPF];
        const float *pb = ew_two_inputs<OP>() ? a.in1 : nullptr;
        OSC_LOAD2(x, y, a.in0, pb, f0)
#pragma unroll
        for (uint32_t k = 0; k < kPF; ++k) {
            const uint32_t f = f0 + k;
            if (f < frames) {
                prm.ramps(a, ramp_mask, f);
                if (ec.due(f)) ec.advance(prm.events(a, ec, f));
                float o0, o1;
                ew_apply<OP>(x[k], y[k], prm.P[0], a.q, o0, o1);
                a.out0[(size_t)f * n + v] = o0;
                if (OP == EW_CROSSFADE) a.out1[(size_t)f * n + v] = o1;
            }
        }
    }
}

/* vectorised form: without value events every element (frame, voice) is independent, so the
   node is a float4 grid-stride stream over [frames x voices] — the HBM-bound form.  Requires
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
        float4 x = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (!ew_no_inputs<OP>()) x = __ldcs(in0 + i);
        float4 y = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (ew_two_inputs<OP>()) y = __ldcs(in1 + i);
        float4 p;
        if (!has_p) p = make_float4(a.cst[0], a.cst[0], a.cst[0], a.cst[0]);
        else if (ramped) { const float r = a.ptab[(size_t)f * kMaxInputs + a.pid[0]]; p = make_float4(r, r, r, r); }
        else p = __ldg(pv + v4);
        float4 o0, o1;
        ew_apply<OP>(x.x, y.x, p.x, a.q, o0.x, o1.x); ew_apply<OP>(x.y, y.y, p.y, a.q, o0.y, o1.y);
        ew_apply<OP>(x.z, y.z, p.z, a.q, o0.z, o1.z); ew_apply<OP>(x.w, y.w, p.w, a.q, o0.w, o1.w);
        __stcs(out0 + i, o0);
        if (OP == EW_CROSSFADE) __stcs(out1 + i, o1);
    }
}

/* graph input generated on device (TWIN_PEAKS audio_in from the per-voice LCG): a source node */
__global__ void __launch_bounds__(kCta) lcg_source_kernel(float *state, float *out, const BlockCtl *ctl, uint32_t n, uint32_t field) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    unsigned long long lcg = (unsigned long long)__float_as_uint(state[(size_t)(field + 0) * n + v]) |
                             ((unsigned long long)__float_as_uint(state[(size_t)(field + 1) * n + v]) << 32);
    const uint32_t frames = ctl->frames;
    for (uint32_t f = 0; f < frames; ++f) out[(size_t)f * n + v] = lcg_next(lcg);
    state[(size_t)(field + 0) * n + v] = __uint_as_float((uint32_t)(lcg & 0xffffffffull));
    state[(size_t)(field + 1) * n + v] = __uint_as_float((uint32_t)(lcg >> 32));
}

/* graph output: voices.audio_out -> audio_out (instance sums in index order) + mix partials */
__global__ void __launch_bounds__(kCta) instance_mix_node_kernel(const __grid_constant__ KernelArgs a, const float *final_stream) {
    __shared__ float tile[kTileF * kTileW];
    __shared__ float isum[kTileF * kCta];
    const uint32_t n = a.n_voices;
    const uint32_t v = blockIdx.x * kCta + threadIdx.x;
    const uint32_t frames = a.ctl->frames;
    const uint32_t col = tile_col(threadIdx.x, a.vpi);
    /* software-pipelined: the 32 independent loads of tile k + 1 are in flight while tile k is folded (three barriers and a
       sequential per-instance fold): 0.41 -> see profiles/r02_ab_instance_mix_pipelined.log per 1 M-voice block */
    float x[kTileF];
    const bool on = v < n;
    {
        const uint32_t fe = min(kTileF, frames);
#pragma unroll
        for (uint32_t ff = 0; ff < kTileF; ++ff) x[ff] = (on && ff < fe) ? __ldcs(final_stream + (ff * n + v)) : 0.0f;
    }
    for (uint32_t f0 = 0; f0 < frames; f0 += kTileF) {
        const uint32_t fe = min(kTileF, frames - f0);
#pragma unroll
        for (uint32_t ff = 0; ff < kTileF; ++ff) tile[ff * kTileW + col] = x[ff];
        const uint32_t f1 = f0 + kTileF;
        if (f1 < frames) {
            const uint32_t fe1 = min(kTileF, frames - f1);
#pragma unroll
            for (uint32_t ff = 0; ff < kTileF; ++ff) x[ff] = (on && ff < fe1) ? __ldcs(final_stream + ((f1 + ff) * n + v)) : 0.0f;
        }
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
