/*
 * kernel_others.cuh — fused one-thread-per-voice kernels for the other reference graphs on
 * the hot path: TwinPeaksGraph, SatGraph_{1x,4x}, README Synth (+ADSR), ElectricPiano.
 * Same structure as kernel_fm.cuh; value inputs are always read per voice (few inputs).
 */
#pragma once
#include "kernels_common.cuh"
#include "kernel_fm.cuh"

namespace oscen {

/* shared prologue for the small graphs: load NP per-voice inputs, apply ramps / value events */
template <uint32_t NP>
struct SmallParams {
    float P[NP];
    OSC_D void load(const KernelArgs &a, uint32_t vs) {
#pragma unroll
        for (uint32_t i = 0; i < NP; ++i) P[i] = a.pv[(size_t)i * a.n_voices + vs];
    }
    OSC_D void ramps(const KernelArgs &a, uint32_t ramp_mask, uint32_t f) {
        if (ramp_mask) {
            const float *row = a.ptab + (size_t)f * kMaxInputs;
#pragma unroll
            for (uint32_t i = 0; i < NP; ++i) if ((ramp_mask >> i) & 1u) P[i] = row[i];
        }
    }
    OSC_D void value_event(const KernelArgs &a, const Event &e, uint32_t v) {
        /* select form: written as `if (id == i) P[i] = x` the compiler turns the unrolled chain into ONE store with a run-time
           index, and P moves from registers to local memory for the whole kernel (LDL in every frame) */
#pragma unroll
        for (uint32_t i = 0; i < NP; ++i) P[i] = (e.input_id == i) ? e.scalar : P[i];
        if (e.input_id < NP) a.pv[(size_t)e.input_id * a.n_voices + v] = e.scalar;
    }
};

/* ============================================================== TwinPeaksGraph ==== */
/* examples/nih-twin-peaks/src/lib.rs:15-48: audio_in -> filter_a, filter_b; out = a - b */
__global__ void __launch_bounds__(kCta) twin_peaks_kernel(const __grid_constant__ KernelArgs a) {
    __shared__ float tile[kTileF * kTileW];
    __shared__ float isum[kTileF * kCta];
    const uint32_t n = a.n_voices;
    const uint32_t v = blockIdx.x * kCta + threadIdx.x;
    const bool active = v < n;
    const uint32_t vs = active ? v : 0u;
    const uint32_t frames = a.ctl->frames, ramp_mask = a.ctl->ramp_mask;
    const bool lcg_on = a.ctl->lcg_on != 0u;
    const float sr = a.u.sr;
    const uint32_t col = tile_col(threadIdx.x, a.vpi);

    SmallParams<3> prm; prm.load(a, vs);
    Lp18State fa, fb;
    lp18_load(fa, a.state, TW_A, n, vs); lp18_load(fb, a.state, TW_B, n, vs);
    unsigned long long lcg = (unsigned long long)__float_as_uint(a.state[(size_t)(TW_LCG + 0) * n + vs]) |
                             ((unsigned long long)__float_as_uint(a.state[(size_t)(TW_LCG + 1) * n + vs]) << 32);
    EventCursor ec; ec.init(a.events, active ? a.ctl->n_events : 0u, vs);

    for (uint32_t f0 = 0; f0 < frames; f0 += kTileF) {
        const uint32_t fe = min(kTileF, frames - f0);
        /* stage this tile's input column in shared memory: 32 independent coalesced loads in flight,
           then a compact (not unrolled) compute loop — the two LP18 bodies are ~400 instructions */
        if (active && !lcg_on && a.stream_in != nullptr) {
#pragma unroll 8
            for (uint32_t ff = 0; ff < fe; ++ff) isum[ff * kCta + threadIdx.x] = __ldcs(a.stream_in + (size_t)(f0 + ff) * n + v);
        }
        /* segmented like the generated kernels (codegen.hpp): per warp, L frames until the first lane's next value event run
           without ramp / event code (so the filters' parameter tests are loop-invariant), then one general frame */
        uint32_t ff = 0;
        while (ff < fe) {
            const uint32_t f = f0 + ff;
            /* (ramp blocks are segmented too: LP18Filter reads its parameters inside process(), nothing derived from them is
               kept outside it — the plain loop's twin loads the frame's ramp row first; two copies, so that without ramps
               nothing in the loop writes a parameter and the parameter tests stay loop-invariant) */
            const uint32_t np = active ? ec.next_frame - f : 0x7FFFFFFFu;
            const uint32_t L = umin(__reduce_min_sync(0xffffffffu, np), fe - ff);
            if (L > 0u) {
#define OSC_TWIN_PLAIN_LOOP(RAMPS)                                                                                               \
                _Pragma("unroll 1") for (uint32_t i = 0; i < L; ++i) {                                                         \
                    float y = 0.0f;                                                                                              \
                    if (active) {                                                                                                \
                        if (RAMPS) prm.ramps(a, ramp_mask, f + i);                                                               \
                        const float x = lcg_on ? lcg_next(lcg) : ((a.stream_in != nullptr) ? isum[(ff + i) * kCta + threadIdx.x] : 0.0f); \
                        const float ya = lp18_process(fa, x, prm.P[0], 0.0f, prm.P[2], sr);                                      \
                        const float yb = lp18_process(fb, x, prm.P[1], 0.0f, prm.P[2], sr);                                      \
                        y = ya - yb;                                                                                             \
                        if (a.voice_out != nullptr) a.voice_out[(size_t)(f + i) * n + v] = y;                                    \
                    }                                                                                                            \
                    tile[(ff + i) * kTileW + col] = y;                                                                           \
                }
                if (ramp_mask == 0u) { OSC_TWIN_PLAIN_LOOP(false) } else { OSC_TWIN_PLAIN_LOOP(true) }
#undef OSC_TWIN_PLAIN_LOOP
                ff += L;
                continue;
            }
            float y = 0.0f;
            if (active) {
                prm.ramps(a, ramp_mask, f);
                if (ec.due(f)) {
                    const uint32_t j = ec.end_of_frame(f);
                    for (uint32_t q = ec.i; q < j; ++q) prm.value_event(a, ec.ev[q], v);
                    ec.advance(j);
                }
                const float x = lcg_on ? lcg_next(lcg) : ((a.stream_in != nullptr) ? isum[ff * kCta + threadIdx.x] : 0.0f);
                const float ya = lp18_process(fa, x, prm.P[0], 0.0f, prm.P[2], sr);
                const float yb = lp18_process(fb, x, prm.P[1], 0.0f, prm.P[2], sr);
                y = ya - yb;
                if (a.voice_out != nullptr) a.voice_out[(size_t)f * n + v] = y;
            }
            tile[ff * kTileW + col] = y;
            ff += 1u;
        }
        tile_flush_mono(a, tile, isum, f0, fe);
    }
    if (active) {
        lp18_store(fa, a.state, TW_A, n, v); lp18_store(fb, a.state, TW_B, n, v);
        a.state[(size_t)(TW_LCG + 0) * n + v] = __uint_as_float((uint32_t)(lcg & 0xffffffffull));
        a.state[(size_t)(TW_LCG + 1) * n + v] = __uint_as_float((uint32_t)(lcg >> 32));
    }
}

/* ============================================================== SatGraph_{1x,4x} === */
/* examples/oversampled-saturator/src/main.rs:64-80; multirate body codegen/emit_frame.rs:114-176 */

/* hb_down_step_r / HbDownRegs (register-resident halfband histories) live in nodes.cuh, next to the shared-memory form */

#ifndef OSCEN_B200_SAT_UNROLL
#define OSCEN_B200_SAT_UNROLL 2      /* A/B on B200 (profiles/r02_ab_sat_unroll.log): 1: 0.817, 2: 0.813, 3: 0.848, 4: 0.835, 8: 1.10 ms per 262 144-voice block */
#endif
#define OSC_PRAGMA_STR(x) _Pragma(#x)
#define OSC_PRAGMA_UNROLL(n) OSC_PRAGMA_STR(unroll n)

template <int FACTOR>
__global__ void __launch_bounds__(kCta) sat_kernel(const __grid_constant__ KernelArgs a) {
    __shared__ float tile[kTileF * kTileW];
    __shared__ float isum[kTileF * kCta];
    const uint32_t n = a.n_voices;
    const uint32_t v = blockIdx.x * kCta + threadIdx.x;
    const bool active = v < n;
    const uint32_t vs = active ? v : 0u;
    const uint32_t frames = a.ctl->frames, ramp_mask = a.ctl->ramp_mask;
    const float sr_inner = a.u.sr * (float)FACTOR;         /* `* N` nodes: sample_rate * N */
    const uint32_t col = tile_col(threadIdx.x, a.vpi);

    SmallParams<2> prm; prm.load(a, vs);
    float phase = a.state[(size_t)SAT_PHASE * n + vs];
    HbDownRegs r1, r2;                                     /* SincDownFir<4>: stage 0 (4x -> 2x), stage 1 (2x -> 1x) */
    if (FACTOR == 4) { hb_down_load(r1, a.state, SAT_H1, n, vs); hb_down_load(r2, a.state, SAT_H2, n, vs); }
    EventCursor ec; ec.init(a.events, active ? a.ctl->n_events : 0u, vs);
    const bool keepv = active && a.voice_out != nullptr;

    for (uint32_t f0 = 0; f0 < frames; f0 += kTileF) {
        const uint32_t fe = min(kTileF, frames - f0);
        /* Inactive lanes of the last CTA render voice 0 again (no events, no stores): a frame body that updates the
           register histories under `if (active)` makes every history value a two-way merge at the end of the frame, which
           pins each shift to a real move; unconditional, the unrolled body renames them away.
           Segmented like the generated kernels: per warp, L frames until the first lane's next value event run without ramp /
           event code (frequency / sample-rate division and the other parameter work become loop-invariant), then one
           general frame. */
#define OSC_SAT_FRAME(FF)                                                                                              \
        {                                                                                                              \
            float y;                                                                                                   \
            if (FACTOR == 1) {                                                                                         \
                const float o = polyblep_process<WAVE_SAW>(phase, 0.0f, prm.P[0], 0.0f, prm.P[1], 0.5f, sr_inner);     \
                y = hardclip_process(o);                                                                               \
            } else {                                                                                                   \
                float buf[4];                                                                                          \
                _Pragma("unroll") for (int i = 0; i < 4; ++i) {                                                        \
                    const float o = polyblep_process<WAVE_SAW>(phase, 0.0f, prm.P[0], 0.0f, prm.P[1], 0.5f, sr_inner); \
                    buf[i] = hardclip_process(o);                                                                      \
                }                                                                                                      \
                /* SincDownFir<4>::downsample: stage 0 twice, stage 1 once (sinc_fir.rs:233-249) */                    \
                const float s0 = hb_down_step_r(r1, buf[0], buf[1]);                                                   \
                const float s1 = hb_down_step_r(r1, buf[2], buf[3]);                                                   \
                y = hb_down_step_r(r2, s0, s1);                                                                        \
            }                                                                                                          \
            if (keepv) a.voice_out[(f0 + (FF)) * n + v] = y;   /* n_voices * 512 < 2^32 (checked at create) */          \
            tile[(FF) * kTileW + col] = active ? y : 0.0f;                                                             \
        }
        uint32_t ff = 0;
        while (ff < fe) {
            const uint32_t f = f0 + ff;
            /* (ramp blocks stay per-frame here: a second copy of this kernel's unrolled plain loop for them costs the ramp-free
               blocks 0.9 % — 0.821 vs 0.814 ms at config 5, profiles/r02_ab_ramp_blocks_handwritten.log — and the saturator's
               benchmark workload never ramps; TwinPeaks and Synth have the second copy) */
            const uint32_t np = ramp_mask != 0u ? 0u : ec.next_frame - f;
            const uint32_t L = umin(__reduce_min_sync(0xffffffffu, np), fe - ff);
            if (L > 0u) {
OSC_PRAGMA_UNROLL(OSCEN_B200_SAT_UNROLL)
                for (uint32_t i = 0; i < L; ++i) OSC_SAT_FRAME(ff + i)
                ff += L;
                continue;
            }
            prm.ramps(a, ramp_mask, f);
            if (ec.due(f)) {
                const uint32_t j = ec.end_of_frame(f);
                for (uint32_t q = ec.i; q < j; ++q) prm.value_event(a, ec.ev[q], v);
                ec.advance(j);
            }
            OSC_SAT_FRAME(ff)
            ff += 1u;
        }
#undef OSC_SAT_FRAME
        tile_flush_mono(a, tile, isum, f0, fe);
    }
    if (active) {
        a.state[(size_t)SAT_PHASE * n + v] = phase;
        if (FACTOR == 4) { hb_down_store(r1, a.state, SAT_H1, n, v); hb_down_store(r2, a.state, SAT_H2, n, v); }
    }
}

/* ===================================================== README Synth (+ ADSR vca) ==== */
/* README.md:24-53; ADSR -> Gain wiring as oscen-lib/benches/static_vs_runtime.rs:57-59 */
template <bool WITH_ADSR>
__global__ void __launch_bounds__(kCta) synth_kernel(const __grid_constant__ KernelArgs a) {
    __shared__ float tile[kTileF * kTileW];
    __shared__ float isum[kTileF * kCta];
    const uint32_t n = a.n_voices;
    const uint32_t v = blockIdx.x * kCta + threadIdx.x;
    const bool active = v < n;
    const uint32_t vs = active ? v : 0u;
    const uint32_t frames = a.ctl->frames, ramp_mask = a.ctl->ramp_mask;
    const float sr = a.u.sr;
    const uint32_t col = tile_col(threadIdx.x, a.vpi);

    SmallParams<8> prm; prm.load(a, vs);
    float ph_mod = a.state[(size_t)SY_MOD * n + vs], ph_car = a.state[(size_t)SY_CAR * n + vs];
    TptState tp; tpt_load(tp, a.state, SY_TPT, n, vs);
    AdsrState env; adsr_load(env, a.state, SY_ENV, n, vs);
    AdsrDerived d;
    adsr_derive_r(d, prm.P[4], prm.P[5], prm.P[7], sr);
    AdsrHot hot; adsr_refresh<false>(env, hot, prm.P[6], d);
    EventCursor ec; ec.init(a.events, active ? a.ctl->n_events : 0u, vs);
    adsr_begin_block(env, d, ec.due(0u));

    for (uint32_t f0 = 0; f0 < frames; f0 += kTileF) {
        const uint32_t fe = min(kTileF, frames - f0);
        /* segmented like the generated kernels (codegen.hpp): per warp, L frames until the first lane's next event or
           envelope stage completion run without ramp / event / stage code, then one general frame */
        uint32_t ff = 0;
        while (ff < fe) {
            const uint32_t f = f0 + ff;
            /* (a block whose ramps stay off the envelope's inputs — attack, decay, sustain, release: the step data and the
               clamped sustain are derived outside process() — is segmented too; its plain loop loads the frame's ramp row first) */
            const bool ramp_seg = (ramp_mask & 0xF0u) == 0u;
            uint32_t np = ((ramp_mask != 0u && !ramp_seg) || f == 0u) ? 0u : (active ? ec.next_frame - f : 0x7FFFFFFFu);
            if (WITH_ADSR) np = umin(np, adsr_plain_frames(env));
            const uint32_t L = umin(__reduce_min_sync(0xffffffffu, np), fe - ff);
            if (L > 0u) {
                const bool ad = (env.stage == ST_ATTACK) | (env.stage == ST_DECAY), rl = env.stage == ST_RELEASE;
                const float tg = env.stage == ST_ATTACK ? 1.0f : hot.sus, kk = env.stage == ST_ATTACK ? d.a_k : d.d_k;
                const bool any_rl = WITH_ADSR && __any_sync(0xffffffffu, rl);
#define OSC_SYNTH_PLAIN_LOOP(RAMPS)                                                                                              \
                for (uint32_t i = 0; i < L; ++i) {                                                                               \
                    float y = 0.0f;                                                                                              \
                    if (active) {                                                                                                \
                        if (RAMPS) prm.ramps(a, ramp_mask, f + i);                                                               \
                        const float m = polyblep_process<WAVE_SINE>(ph_mod, 0.0f, prm.P[0], 0.0f, prm.P[1], 0.5f, sr);           \
                        float g = 1.0f;                                                                                          \
                        if (WITH_ADSR) g = any_rl ? adsr_plain_rel(env, ad, rl, tg, kk) : adsr_ad(env, ad, tg, kk);              \
                        const float c = polyblep_process<WAVE_SAW>(ph_car, 0.0f, prm.P[2], m, 0.5f, 0.5f, sr);                   \
                        const float lp = tpt_process(tp, c, prm.P[3], 0.707f, 0.0f, sr);                                         \
                        y = WITH_ADSR ? lp * g : lp;             /* vca: Gain */                                                 \
                        if (a.voice_out != nullptr) a.voice_out[(size_t)(f + i) * n + v] = y;                                    \
                    }                                                                                                            \
                    tile[(ff + i) * kTileW + col] = y;                                                                           \
                }
                if (ramp_mask == 0u) { OSC_SYNTH_PLAIN_LOOP(false) } else { OSC_SYNTH_PLAIN_LOOP(true) }
#undef OSC_SYNTH_PLAIN_LOOP
                if (WITH_ADSR && !any_rl && ad) env.rem -= L;    /* the Release form counts down per frame itself */
                ff += L;
                continue;
            }
            float y = 0.0f;
            if (active) {
                prm.ramps(a, ramp_mask, f);
                uint32_t j = ec.i;
                const bool due = ec.due(f);
                if (due) {
                    j = ec.end_of_frame(f);
                    for (uint32_t q = ec.i; q < j; ++q) if (ec.ev[q].input_id != kEventGate) prm.value_event(a, ec.ev[q], v);
                }
                const bool changed = (ramp_mask != 0u) || due;
                if (WITH_ADSR && changed && adsr_derived_stale(d, prm.P[4], prm.P[5], prm.P[7])) adsr_derive_r(d, prm.P[4], prm.P[5], prm.P[7], sr);
                if (due) {
                    if (WITH_ADSR)
                        for (uint32_t q = ec.i; q < j; ++q) if (ec.ev[q].input_id == kEventGate)
                            adsr_on_gate(env, ec.ev[q].scalar, prm.P[4], prm.P[7], d);
                    ec.advance(j);
                }
                if (WITH_ADSR && changed) adsr_refresh(env, hot, prm.P[6], d);
                const float m = polyblep_process<WAVE_SINE>(ph_mod, 0.0f, prm.P[0], 0.0f, prm.P[1], 0.5f, sr);
                float g = 1.0f;
                if (WITH_ADSR) g = adsr_process_t<false>(env, hot, prm.P[6], d);
                const float c = polyblep_process<WAVE_SAW>(ph_car, 0.0f, prm.P[2], m, 0.5f, 0.5f, sr);
                const float lp = tpt_process(tp, c, prm.P[3], 0.707f, 0.0f, sr);
                y = WITH_ADSR ? lp * g : lp;                     /* vca: Gain */
                if (a.voice_out != nullptr) a.voice_out[(size_t)f * n + v] = y;
            }
            tile[ff * kTileW + col] = y;
            ff += 1u;
        }
        tile_flush_mono(a, tile, isum, f0, fe);
    }
    if (active) {
        a.state[(size_t)SY_MOD * n + v] = ph_mod; a.state[(size_t)SY_CAR * n + v] = ph_car;
        tpt_store(tp, a.state, SY_TPT, n, v);
        adsr_store(env, a.state, SY_ENV, n, v);
    }
}

/* ================================================================ ElectricPiano ==== */
/* examples/electric-piano/src/electric_piano_voice.rs (AmplitudeSource, OscillatorBank,
   ElectricPianoVoiceNode), tremolo.rs, main.rs:33-97 (16 voices -> sum -> Tremolo -> Frame<2>).
   Per-voice state (259 words) lives in a [field][thread] shared-memory tile for the block. */
/* the two reference spectra (kEpVel0 / kEpVel127) live in nodes.cuh, next to the per-node forms of these nodes */

/* lanes per voice and harmonics per lane of epiano_kernel.  Measured on B200 at 131 072 instances x 16 voices
   (profiles/r02_ab_epiano_lanes.log): 4 lanes x 8 harmonics (128 registers, 2 CTAs of 256 threads per SM) 3.38e10
   voice-samples/s, 2 lanes x 16 harmonics (168 registers, 3 CTAs of 128 threads) 4.26e10 — the per-lane control flow
   (event / ramp / step / frequency tests, tile write) is replicated per lane, so fewer, fatter lanes issue fewer
   instructions per voice-sample, and 16 independent rotation chains per lane hide the FP32 latency at 12 warps per SM. */
#ifndef OSCEN_B200_EP_LANES
#define OSCEN_B200_EP_LANES 2
#endif
#ifndef OSCEN_B200_EP_SEGMENTS
#define OSCEN_B200_EP_SEGMENTS 0                   /* 1: per-warp plain segments like the other kernels — bit-identical and 4 % SLOWER here
                                                      (13.2 vs 12.7 ms per 131 072 x 16-voice block, profiles/r02_ab_epiano_segments.log):
                                                      the per-iteration tests it removes are a few instructions next to 16 harmonics x 12
                                                      FP32 operations per lane, and the second copy of the body costs more than it saves */
#endif
#ifndef OSCEN_B200_EP_MIN_BLOCKS
#define OSCEN_B200_EP_MIN_BLOCKS 3                 /* CTAs per SM the register allocation aims at */
#endif
constexpr uint32_t kEpLanes = OSCEN_B200_EP_LANES;
constexpr uint32_t kEpH = 32u / kEpLanes;
#define EPS(field, i) S[(size_t)((field) + (i)) * n_stride]
/* The two rare per-voice transitions run out of line ON THE STATE ROWS IN HBM (the caller reloads the registers they
   replace): inlined into the frame loop, their bodies (a 32-step product chain, eight double-precision cos / sin) cost the
   loop more than a hundred registers.
   AmplitudeSource::trigger_note + OscillatorBank::on_gate for the harmonics [h0, h0 + 8) of one voice, velocity > 0:
   writes decay, release, current_value and the unit oscillators (electric_piano_voice.rs:112-123, 230-297). */
static __device__ __noinline__ void ep_note_on_rows(float *S, size_t n_stride, uint32_t h0, float velocity, float brightness,
                                                    float velocity_scaling, float decay_rate, float harmonic_decay, float key_scaling,
                                                    float release_rate) {
    const float note = 60.0f;                                   /* note_pitch is never updated */
    const float base_decay_rate = (100.0f - decay_rate) / 40000.0f;
    const float harmonic_scaling = 1.0f - ((100.0f - harmonic_decay) / 200000.0f);
    const float scaling_multiplier = (48.0f - note) / 12.0f;
    const float key_scaling_factor = scaling_multiplier * (key_scaling * 0.02f);
    const float adjusted = (key_scaling_factor > 0.0f)
        ? 1.0f - (base_decay_rate / (1.0f + key_scaling_factor))
        : 1.0f - (base_decay_rate * (1.0f - key_scaling_factor));
    float scaling = 1.0f;                                       /* scaling *= harmonic_scaling, i times */
    for (uint32_t i = 0; i < h0; ++i) scaling *= harmonic_scaling;
    const float rv = 0.999f - ((100.0f - release_rate) / 1000.0f);
    float bs = -0.2f + (0.8f * (brightness * 0.01f));
    bs += velocity * velocity_scaling * 0.01f * 0.5f;
    for (uint32_t i = 0; i < kEpH; ++i) {
        EPS(EP_DEC, h0 + i) = adjusted * scaling; scaling *= harmonic_scaling;
        EPS(EP_REL, h0 + i) = rv;
        float amp = (kEpVel127[h0 + i] * velocity) + (kEpVel0[h0 + i] * (1.0f - velocity));
        amp *= 1.0f + bs * (float)(h0 + i);
        EPS(EP_CUR, h0 + i) = amp;
        EPS(EP_ORE, h0 + i) = 1.0f; EPS(EP_OIM, h0 + i) = 0.0f;
    }
}
/* address of a state row of this voice, computed where it is used: the empty asm keeps the compiler from hoisting dozens of
   loop-invariant 64-bit row addresses above the frame loop (where they would be spilled to local memory) */
OSC_D float *ep_row(float *S, size_t n_stride, uint32_t field) {
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+l"(n_stride));
#endif
    return S + (size_t)field * n_stride;
}
/* OscillatorBank::update_multipliers past its early return (electric_piano_voice.rs:132-149): new rotation multipliers
   and unit oscillators for the harmonics [h0, h0 + 8) */
static __device__ __noinline__ void ep_multiplier_rows(float *S, size_t n_stride, uint32_t h0, float freq, float sr) {
    const float nyquist = sr * 0.5f;
    for (uint32_t i = 0; i < kEpH; ++i) {
        const float hf = freq * (float)(h0 + i + 1u);
        float re = 1.0f, im = 0.0f;
        if (hf < nyquist) {
            const float angle = 2.0f * kPi * hf / sr;
            re = om_cosf_u(angle); im = om_sinf_u(angle);
        }
        EPS(EP_MRE, h0 + i) = re; EPS(EP_MIM, h0 + i) = im;
        EPS(EP_ORE, h0 + i) = 1.0f; EPS(EP_OIM, h0 + i) = 0.0f;
    }
}

/*
 * kEpLanes (2) threads = one voice: lane q of a voice owns the harmonics [16q, 16q + 16) and keeps their 96 state words (current /
 * target amplitude, oscillator and multiplier, each complex) in REGISTERS for the whole block — the voice state moves
 * HBM -> registers -> HBM once per block (2 * 1036 B per voice) instead of once per 32-frame tile per harmonic chunk as in
 * the first version (profiles/r02_ncu_epiano_v1_tiles.md: 8.9 GB of DRAM traffic per block, 835 SASS instructions per
 * warp-frame of 32 voices, 12 warps per SM at 168 registers).
 *
 * The reference adds the 32 weighted harmonics in index order (sum += osc[i].imag * amp[i], i = 0..31,
 * electric_piano_voice.rs:160-166); to keep that rounding order the partial sum travels through the voice's lanes as a
 * SYSTOLIC pipeline: in iteration j lane q renders frame j - q — its own rotations and amplitude steps need nothing
 * from the other lanes — and then continues the sum lane q - 1 left for that frame one iteration earlier (one SHFL.UP per
 * iteration).  All lanes are busy except for kEpLanes - 1 iterations at either end of the block.  Everything per-voice
 * (interpolation step, release flag, last frequency, ramps, gate / value events) is replayed by each lane at its own
 * frame, so the lanes never wait for each other.  64 voices per CTA as everywhere (64 x kEpLanes threads), so the instance fold, the
 * mix partials and the tile geometry are those of the other kernels.
 *
 * Flush per 32 frames: instance sums (voices in index order) by all threads, the tremolo's phase recurrence by one thread
 * per instance (serial in time, 4 flops per frame), then sine / pan / Frame<2> outputs by all threads again.
 */
constexpr uint32_t kEpThreads = kCta * kEpLanes;
constexpr uint32_t kEpDynFloats = kTileF * kTileW + kTileF * kCta * 2u + kTileF * kCta + kTileF * kCta;

__global__ void __launch_bounds__(kEpThreads, OSCEN_B200_EP_MIN_BLOCKS) epiano_kernel(const __grid_constant__ KernelArgs a) {
    extern __shared__ float dyn[];          /* tile [kTileF][kTileW] | isum [kTileF][kCta][2] | ssum [kTileF][kCta] | tph [kTileF][kCta] */
    float *tile = dyn;
    float *isum = tile + kTileF * kTileW;
    float *ssum = isum + kTileF * kCta * 2u;
    float *tph = ssum + kTileF * kCta;
    const uint32_t n = a.n_voices;
    const uint32_t vl = threadIdx.x / kEpLanes, q = threadIdx.x % kEpLanes, h0 = q * kEpH;
    const uint32_t v = blockIdx.x * kCta + vl;
    const bool active = v < n;
    const uint32_t vs = active ? v : 0u;
    const uint32_t frames = a.ctl->frames, ramp_mask = a.ctl->ramp_mask;
    const float sr = a.u.sr;
    const uint32_t vpi = a.vpi;
    const uint32_t col = tile_col(vl, vpi);

    SmallParams<9> prm; prm.load(a, vs);
    float *S = a.state + vs;
    const size_t n_stride = n;
    EventCursor ec; ec.init(a.events, active ? a.ctl->n_events : 0u, vs);
    const uint32_t ipc = kCta / vpi;
    const uint32_t my_inst = blockIdx.x * ipc + threadIdx.x;
    float trem_phase = (threadIdx.x < ipc && my_inst < a.n_inst) ? a.inst_state[my_inst] : 0.0f;

    float cur[kEpH], tgt[kEpH], ore[kEpH], oim[kEpH], mre[kEpH], mim[kEpH];
#pragma unroll
    for (int i = 0; i < (int)kEpH; ++i) {
        cur[i] = EPS(EP_CUR, h0 + i); tgt[i] = EPS(EP_TGT, h0 + i);
        ore[i] = EPS(EP_ORE, h0 + i); oim[i] = EPS(EP_OIM, h0 + i);
        mre[i] = EPS(EP_MRE, h0 + i); mim[i] = EPS(EP_MIM, h0 + i);
    }
    uint32_t step = (uint32_t)EPS(EP_STEP, 0);
    bool released = EPS(EP_RELEASED, 0) != 0.0f;
    float lastf = EPS(EP_LASTF, 0);
    float part = 0.0f;                                        /* this lane's running sum of the frame it rendered last */
    const bool keepv = active && a.voice_out != nullptr;

    uint32_t j = 0;                                           /* pipeline iteration: lane q renders frame j - q */
    for (uint32_t f0 = 0; f0 < frames; f0 += kTileF) {
        const uint32_t fe = min(kTileF, frames - f0);
        const uint32_t j_end = f0 + fe + (kEpLanes - 1u);     /* lane 3 has then written frames [f0, f0 + fe) */
        while (j < j_end) {
            /* (A/B form, off by default: see OSCEN_B200_EP_SEGMENTS) per warp, L pipeline iterations in which no lane has an event, a ramp, the
               first frame of the block, an interpolation step that computes a target (step 0) or lands on it (step 64), or a
               frame outside the block run the plain body — no event / step / frequency tests (the frequency only moves with
               an event, a ramp or between blocks) and the interpolation without its select; then ONE general iteration. */
            if (OSCEN_B200_EP_SEGMENTS) {
                const uint32_t fq = j - q;                    /* j < q wraps: not a frame yet */
                uint32_t np = 0u;
                if (fq < frames && fq != 0u && ramp_mask == 0u) {
                    np = frames - fq;
                    if (active) np = umin(umin(np, ec.next_frame - fq), (step - 1u) < 63u ? 64u - step : 0u);
                }
                const uint32_t L = umin(__reduce_min_sync(0xffffffffu, np), j_end - j);
                if (L > 0u) {
                    for (uint32_t i_ = 0; i_ < L; ++i_) {
                        float sum = __shfl_up_sync(0xffffffffu, part, 1);
                        if (q == 0u) sum = 0.0f;
                        const float t = (float)(step + 1u) / 64.0f;
                        const float omt = 1.0f - t;
#pragma unroll
                        for (int i = 0; i < (int)kEpH; ++i) cur[i] = cur[i] * omt + tgt[i] * t;
                        step += 1u;
#pragma unroll
                        for (int i = 0; i < (int)kEpH; ++i) {
                            const float nr = ore[i] * mre[i] - oim[i] * mim[i];
                            const float ni = ore[i] * mim[i] + oim[i] * mre[i];
                            ore[i] = nr; oim[i] = ni;
                            sum += ni * cur[i];
                        }
                        part = sum;
                        if (q == kEpLanes - 1u) {             /* inside a segment the last lane's frame is always in this tile */
                            const float y = active ? sum * 3.0f : 0.0f;
                            if (keepv) a.voice_out[(fq + i_) * n + v] = y;
                            tile[(fq + i_ - f0) * kTileW + col] = y;
                        }
                    }
                    j += L;
                    continue;
                }
            }
            float sum = __shfl_up_sync(0xffffffffu, part, 1);
            if (q == 0u) sum = 0.0f;
            const uint32_t f = j - q;                         /* j < q wraps: not a frame yet */
            if (active && f < frames) {
                prm.ramps(a, ramp_mask, f);
                if (ec.due(f)) {
                    const uint32_t je = ec.end_of_frame(f);
                    for (uint32_t k = ec.i; k < je; ++k) if (ec.ev[k].input_id != kEventGate) prm.value_event(a, ec.ev[k], v);
                    for (uint32_t k = ec.i; k < je; ++k) if (ec.ev[k].input_id == kEventGate) {
                        const float vel = ec.ev[k].scalar;          /* AmplitudeSource::on_gate, OscillatorBank::on_gate */
                        if (vel > 0.0f) {
                            ep_note_on_rows(S, n_stride, h0, vel, prm.P[1], prm.P[2], prm.P[3], prm.P[4], prm.P[5], prm.P[6]);
                            const float *row = ep_row(S, n_stride, EP_CUR + h0);
#pragma unroll
                            for (int i = 0; i < (int)kEpH; ++i) { cur[i] = row[(size_t)i * n_stride]; ore[i] = 1.0f; oim[i] = 0.0f; }
                            released = false;
                        } else {
                            released = true;
                        }
                        step = 0u;
                    }
                    ec.advance(je);
                }
                /* AmplitudeSource::process, electric_piano_voice.rs:322-351 */
                if (step == 0u) {
                    const float *row = ep_row(S, n_stride, (released ? EP_REL : EP_DEC) + h0);
#pragma unroll
                    for (int i = 0; i < (int)kEpH; ++i) tgt[i] = cur[i] * row[(size_t)i * n_stride];
                }
                {   /* interpolate towards the target for 64 frames, then land on it (one frame in 65): select form — as a
                       branch the two arms meet in 16 register moves per frame */
                    const bool lerp = step < 64u;
                    const float t = (float)(step + 1u) / 64.0f;
                    const float omt = 1.0f - t;
#pragma unroll
                    for (int i = 0; i < (int)kEpH; ++i) { const float nc = cur[i] * omt + tgt[i] * t; cur[i] = lerp ? nc : tgt[i]; }
                    step = lerp ? step + 1u : 0u;
                }
                /* OscillatorBank::process, electric_piano_voice.rs:127-170 */
                const float freq = prm.P[0];
                if (freq > 0.0f && !(fabsf(lastf - freq) < 0.01f)) {
                    lastf = freq;
                    ep_multiplier_rows(S, n_stride, h0, freq, sr);
                    const float *row = ep_row(S, n_stride, EP_MRE + h0);
#pragma unroll
                    for (int i = 0; i < (int)kEpH; ++i) { mre[i] = row[(size_t)i * n_stride]; mim[i] = row[(size_t)(i + 32) * n_stride]; ore[i] = 1.0f; oim[i] = 0.0f; }
                }
#pragma unroll
                for (int i = 0; i < (int)kEpH; ++i) {
                    const float nr = ore[i] * mre[i] - oim[i] * mim[i];
                    const float ni = ore[i] * mim[i] + oim[i] * mre[i];
                    ore[i] = nr; oim[i] = ni;
                    sum += ni * cur[i];
                }
                part = sum;
            }
            if (q == kEpLanes - 1u && f - f0 < fe) {          /* the voice's last lane owns the finished frame */
                const float y = active ? sum * 3.0f : 0.0f;   /* output scaling */
                if (keepv) a.voice_out[f * n + v] = y;        /* n_voices * 512 < 2^32 (checked at create) */
                tile[(f - f0) * kTileW + col] = y;
            }
            j += 1u;
        }
        /* ---- flush: instance sum (index order) -> Tremolo (sequential in time) -> Frame<2> */
        __syncthreads();
        for (uint32_t t = threadIdx.x; t < fe * ipc; t += kEpThreads) {
            const uint32_t ff = t / ipc, i = t - ff * ipc;
            const float *row = tile + ff * kTileW + tile_col(i * vpi, vpi);
            float s = 0.0f;
            for (uint32_t k = 0; k < vpi; ++k) s = s + row[k];
            ssum[ff * kCta + i] = s;
        }
        const uint32_t v_first = min((blockIdx.x * ipc + (threadIdx.x % ipc)) * vpi, n - 1u);   /* instance-level inputs: first voice */
        const float depth0 = a.pv[(size_t)7 * n + v_first];
        const float rate0 = a.pv[(size_t)8 * n + v_first];
        if (threadIdx.x < ipc) {                               /* Tremolo phase: phase = fract(phase + rate / sr) per frame */
            for (uint32_t ff = 0; ff < fe; ++ff) {
                const float rate = ((ramp_mask >> 8) & 1u) ? a.ptab[(size_t)(f0 + ff) * kMaxInputs + 8] : rate0;
                tph[ff * kCta + threadIdx.x] = trem_phase;
                const float inc = rate / sr;
                trem_phase = fractf(trem_phase + inc);
            }
        }
        __syncthreads();
        for (uint32_t t = threadIdx.x; t < fe * ipc; t += kEpThreads) {
            const uint32_t ff = t / ipc, i = t - ff * ipc;
            const float depth = ((ramp_mask >> 7) & 1u) ? a.ptab[(size_t)(f0 + ff) * kMaxInputs + 7] : depth0;
            float l, r;
            tremolo_output(tph[ff * kCta + i], ssum[ff * kCta + i], depth, l, r);
            isum[(ff * kCta + i) * 2 + 0] = l; isum[(ff * kCta + i) * 2 + 1] = r;
            const uint32_t inst = blockIdx.x * ipc + i;
            if (a.inst_out != nullptr && inst < a.n_inst) {
                a.inst_out[((size_t)(f0 + ff) * a.n_inst + inst) * 2 + 0] = l;
                a.inst_out[((size_t)(f0 + ff) * a.n_inst + inst) * 2 + 1] = r;
            }
        }
        __syncthreads();
        for (uint32_t t = threadIdx.x; t < fe * 2u; t += kEpThreads) {   /* mix partial: in-order over the CTA's instances */
            const uint32_t ff = t >> 1, c = t & 1u;
            float s = 0.0f;
            for (uint32_t i = 0; i < ipc; ++i) s = s + isum[(ff * kCta + i) * 2 + c];
            a.mix_partial[((size_t)blockIdx.x * kMaxBlock + f0 + ff) * 2 + c] = s;
        }
        /* no barrier here: the next tile's frames only write `tile` (last read before the second barrier above), and `isum`
           is not written again before the next flush's second barrier, which every thread reaches after its reads here */
    }
    if (active) {
        float *row = ep_row(S, n_stride, h0);
#pragma unroll
        for (int i = 0; i < (int)kEpH; ++i) {
            row[(size_t)(EP_CUR + i) * n_stride] = cur[i]; row[(size_t)(EP_TGT + i) * n_stride] = tgt[i];
            row[(size_t)(EP_ORE + i) * n_stride] = ore[i]; row[(size_t)(EP_OIM + i) * n_stride] = oim[i];
            row[(size_t)(EP_MRE + i) * n_stride] = mre[i]; row[(size_t)(EP_MIM + i) * n_stride] = mim[i];
        }
        if (q == 0u) { EPS(EP_STEP, 0) = (float)step; EPS(EP_RELEASED, 0) = released ? 1.0f : 0.0f; EPS(EP_LASTF, 0) = lastf; }
    }
    if (threadIdx.x < ipc && my_inst < a.n_inst) a.inst_state[my_inst] = trem_phase;
#undef EPS
}

}  // namespace oscen
