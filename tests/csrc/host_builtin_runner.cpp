/*
 * tests/csrc/host_builtin_runner.cpp — TEST INFRASTRUCTURE.  Runs one block of a HAND-WRITTEN fused kernel of a built-in
 * graph kind on the host (kernel_fm.cuh, kernel_others.cuh, kernel_resample.cuh compiled by g++ against
 * tests/csrc/host_cuda/cuda_runtime.h): the dispatch restates engine.cu's launch_fused() / launch_clip().
 */
#include "kernel_fm.cuh"
#include "kernel_fm_seg.cuh"
#include "fm_select.hpp"
#include "kernel_others.cuh"
#include "kernel_resample.cuh"
#include <vector>

namespace {
using namespace oscen;
template <int P> void clip_policy_launch(const KernelArgs &a, uint32_t grid, int l) {
    switch (l) {
    case 1: host_launch(grid, 1, kCta, [&]() { clip_kernel<P, 1>(a); }); break;
    case 2: host_launch(grid, 1, kCta, [&]() { clip_kernel<P, 2>(a); }); break;
    default: host_launch(grid, 1, kCta, [&]() { clip_kernel<P, 3>(a); }); break;
    }
}
}  // namespace


/* stream_in_lcg: the graph's stream input is generated on the device from the per-voice LCG state (BlockCtl.lcg_on) */
static uint32_t g_lcg_on = 0, g_lcg_field = 0;
extern "C" void oscen_host_set_lcg(uint32_t on, uint32_t field) { g_lcg_on = on; g_lcg_field = field; }
/* oscen_b200_stream_in_lcg's lcg_seed_kernel (engine.cu), restated: the LCG state words of every voice */
extern "C" void oscen_host_lcg_seed(float *state, uint32_t n, unsigned long long seed, uint32_t voice_offset, uint32_t field) {
    for (uint32_t v = 0; v < n; ++v) {
        const unsigned long long s = oscen::lcg_seed_state(seed + (unsigned long long)(voice_offset + v));
        state[(size_t)(field + 0) * n + v] = __uint_as_float((uint32_t)(s & 0xffffffffull));
        state[(size_t)(field + 1) * n + v] = __uint_as_float((uint32_t)(s >> 32));
    }
}

/* FMVoice: which kernel variant the last block ran, and which engine's bound state the next block uses */
static int g_last_variant = -1, g_force_variant = -1;
static uint32_t g_bound_slot = 0;
extern "C" int oscen_host_last_fm_variant(void) { return g_last_variant; }
extern "C" void oscen_host_set_fm_slot(uint32_t slot) { g_bound_slot = slot; }
/* tests: -1 = the product's choice; FM_GENERAL = always the general kernel; FM_LEGACY = the round-1 fast kernel wherever it applies
   (what OSCEN_B200_FM_KERNEL=general|legacy does in the engine) */
extern "C" void oscen_host_force_fm_variant(int v) { g_force_variant = v; }

extern "C" int oscen_host_run_builtin_block(uint32_t kind, int general, uint32_t n_voices, uint32_t vpi, float sample_rate, float *state,
                                            float *pv, const oscen::Event *events, uint32_t n_events, const float *ptab, uint32_t ramp_mask,
                                            uint32_t frames, const float *stream_in, float *voice_out, float *inst_out, float *inst_state,
                                            float *mix_partial, const float *uniform_p) {
    using namespace oscen;
    if (n_voices == 0 || vpi == 0 || kCta % vpi != 0 || n_voices % vpi != 0 || frames > kMaxBlock) return -1;
    BlockCtl ctl{};
    ctl.frames = frames; ctl.n_events = n_events; ctl.ramp_mask = ramp_mask; ctl.lcg_on = g_lcg_on;
    KernelArgs a;
    memset(&a, 0, sizeof a);
    a.n_voices = n_voices; a.vpi = vpi; a.n_inst = n_voices / vpi; a.channels = kind == 4u ? 2u : 1u;
    a.state = state; a.pv = pv; a.events = events; a.ptab = ptab; a.ctl = &ctl; a.stream_in = stream_in;
    a.voice_out = voice_out; a.inst_out = inst_out; a.inst_state = inst_state; a.mix_partial = mix_partial;
    for (uint32_t i = 0; i < kMaxInputs; ++i) a.u.p[i] = uniform_p[i];
    a.u.sr = sample_rate;
    if (kind == 0u) {
        const float *p = a.u.p;
        adsr_derive(a.u.env[0], p[I_OP3_A], p[I_OP3_D], p[I_OP3_R], a.u.sr);
        adsr_derive(a.u.env[1], p[I_OP2_A], p[I_OP2_D], p[I_OP2_R], a.u.sr);
        adsr_derive(a.u.env[2], p[I_OP1_A], p[I_OP1_D], p[I_OP1_R], a.u.sr);
        adsr_derive(a.u.env[3], p[I_F_A], p[I_F_D], p[I_F_R], a.u.sr);
    }
    const uint32_t grid = (n_voices + kCta - 1) / kCta;
    std::vector<float> rows;
    switch (kind) {
    case 0: {
        /* the product's own choice (fm_select.hpp): `general` = per-voice inputs / value events off `frequency`; ramps are judged there */
        static FmBoundState bounds[64];                         /* per harness engine: fm_bound_slot */
        g_last_variant = fm_select_variant(general != 0, ramp_mask, a.u.p, ptab, frames, bounds[g_bound_slot % 64]);
        if (g_force_variant == FM_GENERAL) g_last_variant = FM_GENERAL;
        else if (g_force_variant == FM_LEGACY && g_last_variant != FM_GENERAL) g_last_variant = ramp_mask ? FM_GENERAL : FM_LEGACY;
        if (g_last_variant == FM_SEG_RAMP) {
            rows.resize((size_t)frames * kFmRowFloats);
            fm_fill_rows(ramp_mask, a.u.p, ptab, frames, rows.data());
            a.utab = rows.data();
        }
        switch (g_last_variant) {
        case FM_SEG: host_launch(grid, 1, kCta, [&]() { fm_voice_seg_kernel<false, false>(a); }); break;
        case FM_SEG_FMOD: host_launch(grid, 1, kCta, [&]() { fm_voice_seg_kernel<false, true>(a); }); break;
        case FM_SEG_RAMP: host_launch(grid, 1, kCta, [&]() { fm_voice_seg_kernel<true, true>(a); }); break;
        case FM_LEGACY: host_launch(grid, 1, kCta, [&]() { fm_voice_kernel<false>(a); }); break;
        default: host_launch(grid, 1, kCta, [&]() { fm_voice_kernel<true>(a); }); break;
        }
    } break;
    case 1: host_launch(grid, 1, kCta, [&]() { twin_peaks_kernel(a); }); break;
    case 2: host_launch(grid, 1, kCta, [&]() { sat_kernel<1>(a); }); break;
    case 3: host_launch(grid, 1, kCta, [&]() { sat_kernel<4>(a); }); break;
    case 4: host_launch(grid, 1, kEpThreads, [&]() { epiano_kernel(a); }); break;
    case 5: host_launch(grid, 1, kCta, [&]() { synth_kernel<false>(a); }); break;
    case 6: host_launch(grid, 1, kCta, [&]() { synth_kernel<true>(a); }); break;
    default: {
        if (kind < 7u || kind > 19u) return -1;
        if (kind == 7u) { host_launch(grid, 1, kCta, [&]() { clip_kernel<POL_LATCH, 0>(a); }); break; }
        const int pol = (int)(kind - 8u) / 3, l = 1 + (int)(kind - 8u) % 3;
        switch (pol) {
        case POL_SINC: clip_policy_launch<POL_SINC>(a, grid, l); break;
        case POL_SINC_IIR: clip_policy_launch<POL_SINC_IIR>(a, grid, l); break;
        case POL_LINEAR: clip_policy_launch<POL_LINEAR>(a, grid, l); break;
        default: clip_policy_launch<POL_LATCH>(a, grid, l); break;
        }
    } break;
    }
    return 0;
}

/* the final mix bus: mix_reduce_kernel (kernels_common.cuh), one 256-thread CTA per output element, as launch_mix_reduce() */
extern "C" int oscen_host_mix_reduce(const float *partial, float *mix, uint32_t n_cta, uint32_t channels, uint32_t frames) {
    using namespace oscen;
    BlockCtl ctl{};
    ctl.frames = frames;
    host_launch(frames * channels, 1, 256, [&]() { mix_reduce_kernel(partial, mix, n_cta, channels, &ctl); });
    return 0;
}
