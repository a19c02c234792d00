/*
 * fm_select.hpp — host-side choice of the FMVoice kernel for one block, and the per-frame uniform rows the segmented
 * kernel reads in ramp blocks.  Host-only C++ shared by engine.cu and the CPU suite's host-executed kernel runner
 * (tests/csrc/host_builtin_runner.cpp), so that the test harness takes exactly the product's decisions.
 *
 *   FM_GENERAL   fm_voice_kernel<true>       any input per-voice / value events on inputs other than `frequency` /
 *                                            ramps on envelope times, sustain levels, ratios or `frequency`
 *   FM_SEG       fm_voice_seg_kernel<0,0>    uniform inputs, no ramp, filter cutoff cannot move (filter_env_amount == 0)
 *   FM_SEG_FMOD  fm_voice_seg_kernel<0,1>    uniform inputs, no ramp, cutoff follows the filter envelope
 *   FM_SEG_RAMP  fm_voice_seg_kernel<1,1>    uniform inputs, ramps on {op3/op2 level, op3/op2 feedback, route, cutoff,
 *                                            resonance, filter_env_amount} only (kFmSegRampMask)
 *   FM_LEGACY    fm_voice_kernel<false>      round-1 per-frame kernel, kept for A/B (OSCEN_B200_FM_KERNEL=legacy)
 *
 * The segmented kernel's sine has no range test (om_sinf_rk): the host proves |phase argument| < 120 from the uniform
 * levels and feedbacks of this block AND of the previous one (an operator's `prev_output` state was produced with the
 * previous block's level), else it falls back to FM_GENERAL.
 */
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "kernel_fm_seg.cuh"

namespace oscen {

enum FmVariant : int { FM_GENERAL = 0, FM_LEGACY = 1, FM_SEG = 2, FM_SEG_FMOD = 3, FM_SEG_RAMP = 4 };

inline const char *fm_variant_name(int v) {
    switch (v) {
    case FM_GENERAL: return "fm_voice_kernel<true>";
    case FM_LEGACY: return "fm_voice_kernel<false>";
    case FM_SEG: return "fm_voice_seg_kernel<ramp=0,fmod=0>";
    case FM_SEG_FMOD: return "fm_voice_seg_kernel<ramp=0,fmod=1>";
    default: return "fm_voice_seg_kernel<ramp=1,fmod=1>";
    }
}

/* carried from block to block: the largest |op3_level| / |op2_level| the previous block rendered with */
struct FmBoundState { float lvl3 = 0.0f, lvl2 = 0.0f; };

/*
 * needs_general : some input other than `frequency` is per-voice, or the block carries value events on such inputs
 * ramp_mask     : inputs that follow ptab this block;  ptab: [frames][kMaxInputs] (only the masked columns are valid)
 * uni           : uniform values at block start (for ramped inputs: ignored, ptab rules)
 * Returns the variant and advances `bs`.
 */
inline int fm_select_variant(bool needs_general, uint32_t ramp_mask, const float *uni, const float *ptab, uint32_t frames,
                             FmBoundState &bs) {
    static const int force = []() {
        const char *e = getenv("OSCEN_B200_FM_KERNEL");
        if (!e) return -1;
        if (!strcmp(e, "legacy")) return (int)FM_LEGACY;
        if (!strcmp(e, "general")) return (int)FM_GENERAL;
        return -1;
    }();
    /* the range of |level| / |feedback| over the block */
    float l3 = std::fabs(uni[I_OP3_LEVEL]), l2 = std::fabs(uni[I_OP2_LEVEL]), f3 = std::fabs(uni[I_OP3_FB]), f2 = std::fabs(uni[I_OP2_FB]);
    bool finite = true;
    for (uint32_t i = 0; i < I_FM_N; ++i) if (!std::isfinite(uni[i])) finite = false;
    if (ramp_mask) {
        auto col = [&](uint32_t id, float &m) {
            if (!((ramp_mask >> id) & 1u)) return;
            m = 0.0f;
            for (uint32_t f = 0; f < frames; ++f) {
                const float x = ptab[(size_t)f * kMaxInputs + id];
                if (!std::isfinite(x)) finite = false;
                m = std::fmax(m, std::fabs(x));
            }
        };
        col(I_OP3_LEVEL, l3); col(I_OP2_LEVEL, l2); col(I_OP3_FB, f3); col(I_OP2_FB, f2);
        for (uint32_t id = 0; id < I_FM_N; ++id) if (((ramp_mask >> id) & 1u) && id != I_OP3_LEVEL && id != I_OP2_LEVEL && id != I_OP3_FB && id != I_OP2_FB)
            for (uint32_t f = 0; f < frames; ++f) if (!std::isfinite(ptab[(size_t)f * kMaxInputs + id])) finite = false;
    }
    const float L3 = std::fmax(l3, bs.lvl3), L2 = std::fmax(l2, bs.lvl2);
    bs.lvl3 = l3; bs.lvl2 = l2;
    if (force == FM_GENERAL) return FM_GENERAL;
    if (needs_general || (ramp_mask & ~kFmSegRampMask)) return FM_GENERAL;
    if (force == FM_LEGACY) return ramp_mask ? FM_GENERAL : FM_LEGACY;
    /* |(phase + phase_mod + prev * feedback) * TAU| with |phase| < 1, |prev| <= level, |sin| <= 1, envelopes in [0, 1] */
    const float a3 = 1.0f + L3 * f3, a2 = 1.0f + L3 + L2 * f2, a1 = 1.0f + L2 + L3;
    if (!finite || !(a3 < 18.0f && a2 < 18.0f && a1 < 18.0f)) return FM_GENERAL;       /* 18 * TAU * (1 + eps) < 120 */
    if (ramp_mask) return FM_SEG_RAMP;
    return uni[I_ENV_AMT] == 0.0f ? FM_SEG : FM_SEG_FMOD;
}

/* one FmRow per frame: the block's uniform values with the ramped columns taken from ptab (the exact tick() sequence) */
inline void fm_fill_rows(uint32_t ramp_mask, const float *uni, const float *ptab, uint32_t frames, float *rows) {
    float p[kMaxInputs];
    for (uint32_t i = 0; i < kMaxInputs; ++i) p[i] = uni[i];
    for (uint32_t f = 0; f < frames; ++f) {
        for (uint32_t i = 0; i < I_FM_N; ++i) if ((ramp_mask >> i) & 1u) p[i] = ptab[(size_t)f * kMaxInputs + i];
        const FmRow r = fm_row_from(p);
        memcpy(rows + (size_t)f * kFmRowFloats, &r, sizeof r);
    }
}

}  // namespace oscen
