/*
 * state_io.cuh — per-node state handoff between the SoA rows in HBM (`state[field][voice]`) and
 * registers, and the per-voice input LCG.  Shared by the hand-written fused kernels, the node
 * kernels and the kernels codegen.hpp generates at run time (so it holds device functions only —
 * no __global__ kernels: the generated translation unit includes it).
 */
#pragma once
#include "kernels_common.cuh"

namespace oscen {

OSC_D void adsr_load(AdsrState &s, const float *st, uint32_t base, uint32_t n, uint32_t v) {
    const uint32_t w = __float_as_uint(st[(size_t)(base + 0) * n + v]);      /* stage in bits 0-2, r_gate above */
    s.stage = w & 7u; s.r_gate = w >> 3;
    s.rem = __float_as_uint(st[(size_t)(base + 1) * n + v]);
    s.level = st[(size_t)(base + 2) * n + v];
    s.vel = st[(size_t)(base + 3) * n + v];
}
OSC_D void adsr_store(const AdsrState &s, float *st, uint32_t base, uint32_t n, uint32_t v) {
    st[(size_t)(base + 0) * n + v] = __uint_as_float(s.stage | (umin(s.r_gate, 0x1FFFFFFFu) << 3));
    st[(size_t)(base + 1) * n + v] = __uint_as_float(s.rem);
    st[(size_t)(base + 2) * n + v] = s.level;
    st[(size_t)(base + 3) * n + v] = s.vel;
}
OSC_D void tpt_load(TptState &s, const float *st, uint32_t base, uint32_t n, uint32_t v) {
    s.z0 = st[(size_t)(base + 0) * n + v]; s.z1 = st[(size_t)(base + 1) * n + v];
    s.cur_cutoff = st[(size_t)(base + 2) * n + v]; s.cur_q = st[(size_t)(base + 3) * n + v];
    s.h = st[(size_t)(base + 4) * n + v]; s.g = st[(size_t)(base + 5) * n + v];
    s.k = st[(size_t)(base + 6) * n + v];
}
OSC_D void tpt_store(const TptState &s, float *st, uint32_t base, uint32_t n, uint32_t v) {
    st[(size_t)(base + 0) * n + v] = s.z0; st[(size_t)(base + 1) * n + v] = s.z1;
    st[(size_t)(base + 2) * n + v] = s.cur_cutoff; st[(size_t)(base + 3) * n + v] = s.cur_q;
    st[(size_t)(base + 4) * n + v] = s.h; st[(size_t)(base + 5) * n + v] = s.g;
    st[(size_t)(base + 6) * n + v] = s.k;
}

OSC_D void lp18_load(Lp18State &s, const float *st, uint32_t base, uint32_t n, uint32_t v) {
    s.z0 = st[(size_t)(base + 0) * n + v]; s.z1 = st[(size_t)(base + 1) * n + v];
    s.z2 = st[(size_t)(base + 2) * n + v]; s.g = st[(size_t)(base + 3) * n + v];
    s.h = st[(size_t)(base + 4) * n + v]; s.last_cutoff = st[(size_t)(base + 5) * n + v];
    s.last_fmod = st[(size_t)(base + 6) * n + v]; s.last_resonance = st[(size_t)(base + 7) * n + v];
    s.resonance = st[(size_t)(base + 8) * n + v];
}
OSC_D void lp18_store(const Lp18State &s, float *st, uint32_t base, uint32_t n, uint32_t v) {
    st[(size_t)(base + 0) * n + v] = s.z0; st[(size_t)(base + 1) * n + v] = s.z1;
    st[(size_t)(base + 2) * n + v] = s.z2; st[(size_t)(base + 3) * n + v] = s.g;
    st[(size_t)(base + 4) * n + v] = s.h; st[(size_t)(base + 5) * n + v] = s.last_cutoff;
    st[(size_t)(base + 6) * n + v] = s.last_fmod; st[(size_t)(base + 7) * n + v] = s.last_resonance;
    st[(size_t)(base + 8) * n + v] = s.resonance;
}

/* 64-bit LCG noise of oscen-lib/tests/realtime_safety.rs:21-33 */
OSC_HD unsigned long long lcg_seed_state(unsigned long long seed) {
    return seed * 2862933555777941757ull + 3037000493ull;
}
OSC_HD float lcg_next(unsigned long long &state) {
    state = state * 6364136223846793005ull + 1442695040888963407ull;
    return (float)(uint32_t)(state >> 33) / 2147483648.0f - 1.0f;
}

}  // namespace oscen
