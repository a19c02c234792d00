/*
 * nodes.cuh — device-side node bodies of the oscen hot path, one inline function per
 * reference `SignalProcessor::process()`, written for
 * "one thread = one voice, state in registers for the whole block".
 *
 * Parity rules: compiled with -fmad=false (rustc never fuses a*b+c),
 * IEEE division/sqrt, no fast-math; transcendentals from om_math.h (glibc-faithful).
 * Where the reference recomputes a pure function of unchanged inputs every sample (ADSR step
 * counts / one-pole coefficients, FM phase increment, TPT clamp ratios) the value is cached
 * and recomputed only when its inputs change — result-identical, not an approximation.
 *
 * Citations: path:line under the reference repo (reedrosenbluth/oscen @ 8061f7d5).
 */
#pragma once
#if defined(__CUDACC_RTC__)
#include "rtc_prelude.h"
#else
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>
#endif

#include "om_math.h"

namespace oscen {

#define OSC_HD __host__ __device__ __forceinline__
#define OSC_D __device__ __forceinline__

constexpr float kPi = 3.14159265358979323846f;    /* std::f32::consts::PI  */
constexpr float kTau = 6.28318530717958647692f;   /* std::f32::consts::TAU */

/* f32::clamp: two compares, NaN propagates */
OSC_HD float clampf(float x, float lo, float hi) {
    if (x < lo) x = lo;
    if (x > hi) x = hi;
    return x;
}
/* `x as u32`: truncating, saturating, NaN -> 0 */
OSC_HD uint32_t f32_as_u32(float x) {
#if defined(__CUDA_ARCH__)
    return __float2uint_rz(x);
#else
    if (!(x == x) || x <= 0.0f) return 0u;
    if (x >= 4294967296.0f) return 4294967295u;
    return (uint32_t)x;
#endif
}
OSC_HD uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
OSC_HD uint32_t umax(uint32_t a, uint32_t b) { return a > b ? a : b; }
/* f32::rem_euclid(1.0): fmodf(x,1) is exact and equals x - trunc(x) (sign of zero kept).  For 0 <= x < 2 (every phase
   accumulator in normal operation) trunc(x) is 0 or 1 and the result is never negative, which avoids the slow-pipe FRND:
   x in [0, 1) is returned as is (that keeps -0, what copysign(0, x) gives), x in [1, 2) is x - 1 (+0 at x = 1). */
OSC_HD float rem_euclid1(float x) {
    if (x >= 0.0f && x < 2.0f) return (x >= 1.0f) ? x - 1.0f : x;
    float r = x - truncf(x);
    if (r == 0.0f) r = copysignf(0.0f, x);
    return (r < 0.0f) ? r + 1.0f : r;
}
/* f32::fract = x - trunc(x).  For 0 <= x < 2 (every phase accumulator in normal operation)
   trunc(x) is 0 or 1, which avoids the slow-pipe FRND; other inputs take the general form. */
OSC_HD float fractf(float x) {
    if (x >= 0.0f && x < 2.0f) return (x >= 1.0f) ? x - 1.0f : x - 0.0f;
    return x - truncf(x);
}
/* keep a loop-invariant value in a register (stops the compiler from re-loading it from the
   constant bank inside the frame loop, where LDC latency is exposed) */
#if defined(__CUDACC__)
OSC_D float pin(float x) { asm volatile("" : "+f"(x)); return x; }
OSC_D uint32_t pin(uint32_t x) { asm volatile("" : "+r"(x)); return x; }
OSC_D double pin(double x) { asm volatile("" : "+d"(x)); return x; }
#else   /* plain C++ (tests/host_kernel.py compiles these headers with g++): nothing to pin */
inline float pin(float x) { return x; }
inline uint32_t pin(uint32_t x) { return x; }
inline double pin(double x) { return x; }
#endif

/* tanf only runs when a filter coefficient has to be recomputed (parameter changes): on the device
   it is a real call, so the ~700-instruction fdlibm body is not replicated into every unrolled
   frame of every kernel (the fused TwinPeaks kernel had grown to 1 MB of SASS and spent most of
   its time in instruction-cache misses: stall_no_instruction 6.0 per issue) */
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ float tanf_cold(float x) { return om_tanf(x); }
#define OSC_TANF(x) tanf_cold(x)
#else
#define OSC_TANF(x) om_tanf(x)
#endif

/* ================================================================= AdsrEnvelope == */
/* oscen-lib/src/envelope/adsr.rs */
enum : uint32_t { ST_IDLE = 0, ST_ATTACK = 1, ST_DECAY = 2, ST_SUSTAIN = 3, ST_RELEASE = 4 };
constexpr float kAdsrMinTime = 1.0e-5f;      /* adsr.rs:5 */
constexpr float kAdsrCurveK = 4.6051702f;    /* adsr.rs:8 */

/* live state (what survives from one frame to the next; everything else in the reference
   struct is recomputed before use inside process(), see DESIGN.md "dead state") */
struct AdsrState {
    uint32_t stage, rem;
    float level, vel;
    uint32_t r_gate;          /* `release_samples` as cached by the last process() (adsr.rs:117-134): the countdown a gate-OFF
                                 starts from (set_stage(Release), adsr.rs:136-157) — it lags `release` by one frame, which is
                                 visible when the release time grows on the very frame of the note-off.  Equal to the derived
                                 r_n at all times except from block start to the first processed frame, so it costs no
                                 per-frame work; it travels in the upper bits of the stage word (state_io.cuh). */
};
/* recalculate_cached_steps(), adsr.rs:117-134 — a pure function of (attack, decay, release, sr) */
struct AdsrDerived {
    float attack, decay, release;     /* the raw inputs these were derived from */
    uint32_t a_n, d_n, r_n;
    float a_k, d_k;
};
/* cold (block start / parameter change): a real call on the device so the two expf bodies are not
   replicated at every call site of the frame loop */
#if defined(__CUDA_ARCH__)
static __device__ __noinline__
#else
static inline
#endif
void adsr_derive(AdsrDerived &d, float attack, float decay, float release, float sample_rate) {
    float sr = fmaxf(sample_rate, 1.0f);
    d.attack = attack; d.decay = decay; d.release = release;
    d.a_n = umax(f32_as_u32(fmaxf(attack, kAdsrMinTime) * sr), 1u);
    d.d_n = umax(f32_as_u32(fmaxf(decay, kAdsrMinTime) * sr), 1u);
    d.r_n = umax(f32_as_u32(fmaxf(release, kAdsrMinTime) * sr), 1u);
    d.a_k = 1.0f - om_expf(-kAdsrCurveK / (float)d.a_n);
    d.d_k = 1.0f - om_expf(-kAdsrCurveK / (float)d.d_n);
}
/* device call sites use this form: only the temporary's address escapes into the out-of-line call,
   so the caller's copy of the derived data stays in registers */
OSC_D void adsr_derive_r(AdsrDerived &d, float attack, float decay, float release, float sample_rate) {
    AdsrDerived t;
    adsr_derive(t, attack, decay, release, sample_rate);
    d = t;
}
OSC_D void adsr_pin(AdsrDerived &d) {
    d.a_n = pin(d.a_n); d.d_n = pin(d.d_n); d.r_n = pin(d.r_n); d.a_k = pin(d.a_k); d.d_k = pin(d.d_k);
}
OSC_HD bool adsr_derived_stale(const AdsrDerived &d, float attack, float decay, float release) {
    return !(d.attack == attack && d.decay == decay && d.release == release);
}
/* handle_gate_event(), adsr.rs:252-273.  `d` must be derived from THIS frame's inputs
   (edges are assigned before process_event_inputs(), codegen/emit_node.rs:365-379).
   set_stage(Attack|Decay|Release) always finds samples >= 1, so it reduces to the two stores
   below; the release increment it computes is recomputed in process() before its first use. */
OSC_D void adsr_on_gate(AdsrState &s, float velocity, float attack_raw, float release_raw,
                        const AdsrDerived &d) {
    if (velocity > 0.0f) {
        s.vel = clampf(velocity, 0.0f, 1.0f);
        s.r_gate = d.r_n;                                  /* gate-on runs update_sustain_level(): the cache is fresh again */
        if (attack_raw <= kAdsrMinTime) { s.level = 1.0f; s.stage = ST_DECAY; s.rem = d.d_n; }
        else { s.stage = ST_ATTACK; s.rem = d.a_n; }
    } else if (release_raw <= kAdsrMinTime) {
        s.stage = ST_IDLE; s.level = 0.0f; s.rem = 0u;
    } else {
        s.stage = ST_RELEASE; s.rem = s.r_gate;            /* the length cached by the LAST process(); this frame's process()
                                                              clamps it to the new one (adsr_refresh) */
    }
}
/* block start, after `d` was derived from the block's first inputs: a voice with no event at frame 0 is past its first
   process() before any gate can arrive, so its cache is the new length already; a fresh state (0) knows no older one */
OSC_D void adsr_begin_block(AdsrState &s, const AdsrDerived &d, bool event_at_frame0) {
    if (!event_at_frame0 || s.r_gate == 0u) s.r_gate = d.r_n;
}
/* process(), adsr.rs:84-115 (apply_parameters + update_sustain_level), :209-250 (process_stage),
   :175-207 (complete_stage).  Returns `output`.

   Two parts of update_sustain_level() are pure functions of values that only change at known
   points, so they are split out as adsr_refresh() and the per-frame body can skip them
   (REFRESH=false) when the caller guarantees it calls adsr_refresh() at block start and after
   every gate event / change of sustain or of the derived step counts:
     * sustain_level = clamp(clamp(sustain) * velocity)          (inputs: sustain, velocity)
     * samples_remaining = max(min(samples_remaining, stage_len), 1)  — idempotent while the stage
       lengths are unchanged: the countdown only decreases or is reset to exactly a stage length. */
struct AdsrHot { float sus; };
/* COMMIT: the call belongs to a frame's process() (after its gate handlers), which re-caches the step counts; the
   block-start call (before frame 0's events are seen) passes false */
template <bool COMMIT = true>
OSC_D void adsr_refresh(AdsrState &s, AdsrHot &h, float sustain_in, const AdsrDerived &d) {
    if (COMMIT) s.r_gate = d.r_n;
    const float sustain = clampf(sustain_in, 0.0f, 1.0f);
    h.sus = clampf(sustain * s.vel, 0.0f, 1.0f);
    if (s.rem > 0u) {
        if (s.stage == ST_ATTACK) s.rem = umax(umin(s.rem, d.a_n), 1u);
        else if (s.stage == ST_DECAY) s.rem = umax(umin(s.rem, d.d_n), 1u);
        else if (s.stage == ST_RELEASE) s.rem = umax(umin(s.rem, d.r_n), 1u);
    }
}
template <bool REFRESH>
OSC_D float adsr_process_t(AdsrState &s, AdsrHot &h, float sustain_in, const AdsrDerived &d) {
    if (REFRESH) adsr_refresh(s, h, sustain_in, d);
    const float sus = h.sus;
    const uint32_t st = s.stage;
    const bool isA = st == ST_ATTACK, isD = st == ST_DECAY, isR = st == ST_RELEASE;
    uint32_t rem = s.rem;
    float level = s.level;
    if (isA | isD | isR) {
        if (rem > 0u) {
            float delta;
            if (isR) {
                /* update_release_increment(), adsr.rs:159-173 -- recomputed every sample */
                const float cur = clampf(level, 0.0f, 1.0f);
                delta = (cur <= 0.0f) ? 0.0f : -cur / (float)rem;
            } else {
                /* one-pole approach: level += (target - level) * coeff (adsr.rs:213,223) */
                const float tgt = isA ? 1.0f : sus;
                const float k = isA ? d.a_k : d.d_k;
                delta = (tgt - level) * k;
            }
            level = clampf(level + delta, 0.0f, 1.0f);
            rem -= 1u;
        }
        if (rem == 0u) {                                 /* complete_stage(), adsr.rs:175-207 */
            level = isA ? 1.0f : (isD ? sus : 0.0f);
            s.stage = isA ? ST_DECAY : (isD ? ST_SUSTAIN : ST_IDLE);
            rem = isA ? d.d_n : 0u;
        }
    } else {
        level = (st == ST_SUSTAIN) ? sus : 0.0f;
    }
    s.rem = rem;
    s.level = level;
    return level;
}
OSC_D float adsr_process(AdsrState &s, float sustain_in, const AdsrDerived &d) {
    AdsrHot h;
    return adsr_process_t<true>(s, h, sustain_in, d);
}

/* ---- segment forms of the envelope (kernel_fm_seg.cuh, the plain node kernels, generated kernels): a warp runs L "plain"
   frames between event / stage boundaries with each lane's stage fixed */
/* frames this envelope can run before its stage completes (the completing frame itself is a general frame) */
OSC_D uint32_t adsr_plain_frames(const AdsrState &s) {
    const bool timed = (s.stage == ST_ATTACK) | (s.stage == ST_DECAY) | (s.stage == ST_RELEASE);
    return timed ? (s.rem >= 2u ? s.rem - 1u : 0u) : 0x7FFFFFFFu;
}
/* plain frame of one envelope when some lane of the warp is in Release: `ad` / `rl` = this lane is in Attack|Decay /
   Release with rem >= 2 for the whole segment; a static lane (Sustain, Idle) keeps `level`.  The countdown runs per frame
   here (the Release increment divides by it).  Segments without Release lanes use adsr_ad: no division — its slow path is
   a call, and a call inside the frame loop would force loop-invariant constants to be rebuilt after it. */
OSC_D float adsr_plain_rel(AdsrState &s, bool ad, bool rl, float tgt, float k) {
    if (ad | rl) {
        float delta;
        if (rl) {
            const float cur = clampf(s.level, 0.0f, 1.0f);
            delta = (cur <= 0.0f) ? 0.0f : -cur / (float)s.rem;
        } else {
            delta = (tgt - s.level) * k;
        }
        s.level = clampf(s.level + delta, 0.0f, 1.0f);
        s.rem -= 1u;
    }
    return s.level;
}

/* one-pole envelope frame for a lane in Attack|Decay, a no-op for a static lane (no lane of the warp is in Release) */
/* The reference clamps the result to [0, 1] (adsr.rs:213-225); here the clamp is the identity and is left out: level and
   tgt are in [0, 1] (or NaN, which both forms pass through), k = 1 - exp(-4.6051702 / n) <= 0.99 for every step count
   n >= 1, so |RN(RN(tgt - level) * k)| < |tgt - level| and the exact sum lies strictly between level and tgt; rounding to
   nearest is monotonic, so the rounded sum lies in [min(level, tgt), max(level, tgt)].  -0 is kept by both forms. */
OSC_D float adsr_ad(AdsrState &s, bool ad, float tgt, float k) {
    const float nl = s.level + (tgt - s.level) * k;
    s.level = ad ? nl : s.level;
    return s.level;
}
/* =================================================================== FmOperator == */
/* examples/fm-synth/src/nodes/fm_operator.rs:58-76 */
struct FmOpState {
    float phase, prev;
    float inc;                /* phase_inc = (base_freq * ratio) / sample_rate; recomputed by the
                                 caller whenever base_freq or ratio changes (not persisted) */
};
OSC_D float fmop_phase_inc(float base_freq, float ratio, float sample_rate) {
    return (base_freq * ratio) / sample_rate;
}
OSC_D float fmop_process(FmOpState &s, float phase_mod, float feedback, float envelope, float level) {
    const float feedback_mod = s.prev * feedback;
    const float total_phase_mod = phase_mod + feedback_mod;
    const float phase_rad = (s.phase + total_phase_mod) * kTau;
    const float out = om_sinf_u(phase_rad) * envelope * level;
    s.prev = out;
    s.phase += s.inc;
    s.phase = fractf(s.phase);
    return out;
}

/* sinf behind ONE range test: the bounded-argument form (both polynomials, no coefficient selects) for |y| < 120, the
   general routine (a call) beyond */
OSC_D float om_sinf_uk(float y, const OmSinRegs &k) {
    const uint32_t top = (om_asuint(y) >> 20) & 0x7ff;
    if (top >= 0x42f) return om_sinf_slow(y);
    return om_sinf_rk(y, k);
}
OSC_D float fmop_process_k(FmOpState &s, float phase_mod, float feedback, float envelope, float level, const OmSinRegs &sk) {
    const float feedback_mod = s.prev * feedback;
    const float total_phase_mod = phase_mod + feedback_mod;
    const float phase_rad = (s.phase + total_phase_mod) * kTau;
    const float out = om_sinf_uk(phase_rad, sk) * envelope * level;
    s.prev = out;
    const float ph = s.phase + s.inc;
    s.phase = ph - truncf(ph);
    return out;
}

/* ==================================================================== TptFilter == */
/* oscen-lib/src/filters/tpt/mod.rs:69-123 */
struct TptState {
    float z0, z1, cur_cutoff, cur_q, h, g, k;
};
OSC_HD void tpt_update_coefficients(TptState &s, float sample_rate, float cutoff, float q) {
    const float nyquist = sample_rate * 0.5f - FLT_EPSILON;
    const float freq = clampf(cutoff, 20.0f, nyquist);
    const float period = 0.5f / sample_rate;
    const float f = (2.0f * sample_rate) * OSC_TANF(2.0f * kPi * freq * period) * period;
    const float inv_q = 1.0f / q;
    s.h = 1.0f / (1.0f + inv_q * f + f * f);
    s.g = f;
    s.k = s.g + inv_q;
    s.cur_cutoff = cutoff;
    s.cur_q = q;
}
/* TptFilter::new(cutoff,q) then set_sample_rate(sr) + prepare(): mod.rs:46-67,129-131 */
OSC_HD void tpt_init(TptState &s, float sample_rate, float cutoff, float q) {
    s.z0 = 0.0f; s.z1 = 0.0f;
    tpt_update_coefficients(s, 44100.0f, cutoff, q);
    tpt_update_coefficients(s, sample_rate, cutoff, q);
}
OSC_D float tpt_process(TptState &s, float input, float cutoff_in, float q_in, float f_mod,
                        float sample_rate) {
    const float nyquist = sample_rate * 0.5f - FLT_EPSILON;
    const float max_cutoff = fminf(nyquist, 20000.0f);
    const float cutoff_base = clampf(cutoff_in, 20.0f, max_cutoff);
    const float q = clampf(q_in, 0.1f, 10.0f);
    const float modulation = clampf(f_mod, -1.0f, 1.0f);
    float cutoff;
    if (modulation == 0.0f) {
        /* factor = clamp(1, 20/cb <= 1, max/cb >= 1) = 1 and cb*1 = cb: skip the two divisions */
        cutoff = cutoff_base;
    } else {
        const float min_factor = 20.0f / cutoff_base;
        const float max_factor = max_cutoff / cutoff_base;
        const float factor = clampf(1.0f + modulation, min_factor, max_factor);
        cutoff = clampf(cutoff_base * factor, 20.0f, max_cutoff);
    }
    if (fabsf(cutoff - s.cur_cutoff) > FLT_EPSILON || fabsf(q - s.cur_q) > FLT_EPSILON)
        tpt_update_coefficients(s, sample_rate, cutoff, q);
    const float high = (input - s.z0 * s.k - s.z1) * s.h;
    const float band = high * s.g + s.z0;
    const float low = band * s.g + s.z1;
    s.z0 = high * s.g + band;
    s.z1 = band * s.g + low;
    return low;
}

/* =========================================================== PolyBlepOscillator == */
/* oscen-lib/src/oscillators/mod.rs:139-233 */
enum : int { WAVE_SINE = 0, WAVE_SAW = 1, WAVE_SQUARE = 2, WAVE_TRIANGLE = 3 };
OSC_D float poly_blep(float t, float dt) {
    if (dt <= FLT_EPSILON) return 0.0f;
    if (t < dt) { const float x = t / dt; return x + x - x * x - 1.0f; }
    if (t > 1.0f - dt) { const float x = (t - 1.0f) / dt; return x * x + x + x + 1.0f; }
    return 0.0f;
}
OSC_D float poly_blamp(float t, float dt) {
    if (dt <= FLT_EPSILON) return 0.0f;
    if (t < dt) { const float x = t / dt - 1.0f; return -(x * x * x) / 3.0f; }
    if (t > 1.0f - dt) { const float x = (t - 1.0f) / dt + 1.0f; return (x * x * x) / 3.0f; }
    return 0.0f;
}
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ float om_sinf_cold(float y) { return om_sinf_u(y); }
#else
static inline float om_sinf_cold(float y) { return om_sinf_u(y); }
#endif
template <int WAVE>
OSC_D float polyblep_process(float &phase_state, float phase_mod, float frequency_in, float frequency_mod,
                             float amplitude, float pulse_width_in, float sample_rate) {
    const float frequency = fmaxf(frequency_in * (1.0f + frequency_mod), 0.0f);
    float pulse_width = clampf(pulse_width_in, 0.0001f, 0.9999f);
    const float phase = rem_euclid1(phase_state + phase_mod);
    const float freq_per_sample = frequency / fmaxf(sample_rate, FLT_EPSILON);
    const float dt = fminf(freq_per_sample, 1.0f);
    if (pulse_width <= 0.0f) pulse_width = 0.0001f;
    float value;
    if (frequency >= sample_rate * 0.25f) {
        value = (WAVE == WAVE_SINE) ? om_sinf_u(phase * kTau) : om_sinf_cold(phase * kTau);   /* rare for the other waves: a call */
    } else if (WAVE == WAVE_SINE) {
        value = om_sinf_u(phase * kTau);
    } else if (WAVE == WAVE_SAW) {
        float y = 2.0f * phase - 1.0f;
        y -= poly_blep(phase, dt);
        value = y;
    } else if (WAVE == WAVE_SQUARE) {
        float y = (phase < pulse_width) ? 1.0f : -1.0f;
        y += poly_blep(phase, dt);
        const float t = rem_euclid1(phase + 1.0f - pulse_width);
        y -= poly_blep(t, dt);
        value = y;
    } else {
        float y = 4.0f * phase;
        if (y >= 3.0f) y -= 4.0f;
        else if (y > 1.0f) y = 2.0f - y;
        const float t1 = rem_euclid1(phase + 0.25f);
        const float t2 = rem_euclid1(phase + 0.75f);
        value = y + 4.0f * dt * (poly_blamp(t1, dt) - poly_blamp(t2, dt));
    }
    value = value * amplitude;
    phase_state = rem_euclid1(phase_state + freq_per_sample);
    return value;
}

/* =================================================================== LP18Filter == */
/* examples/nih-twin-peaks/src/lp18_filter.rs:62-107 */
struct Lp18State {
    float z0, z1, z2, g, h;
    float last_cutoff, last_fmod, last_resonance, resonance;
};
OSC_HD void lp18_update_cutoff(Lp18State &s, float cutoff, float fmod, float sample_rate) {
    const float modulated = cutoff + fmod;
    const float fc = clampf(modulated / sample_rate, 0.001f, 0.33f);
    s.g = OSC_TANF(kPi * fc);
}
/* LP18Filter::new(cutoff,res) + set_sample_rate + prepare(): lp18_filter.rs:44-60,74-77 */
OSC_HD void lp18_init(Lp18State &s, float cutoff, float resonance, float sample_rate) {
    s.z0 = s.z1 = s.z2 = 0.0f;
    s.resonance = clampf(resonance, 0.0f, 0.99f);
    s.last_cutoff = cutoff; s.last_fmod = 0.0f; s.last_resonance = resonance;
    lp18_update_cutoff(s, cutoff, 0.0f, sample_rate);
    s.h = 2.0f * s.resonance;
}
OSC_D float lp18_process(Lp18State &s, float input, float cutoff, float fmod, float resonance_in,
                         float sample_rate) {
    s.resonance = resonance_in;                       /* value edge assigned before process() */
    if (cutoff != s.last_cutoff || fmod != s.last_fmod) {
        s.last_cutoff = cutoff; s.last_fmod = fmod;
        lp18_update_cutoff(s, cutoff, fmod, sample_rate);
    }
    if (s.resonance != s.last_resonance) {
        s.last_resonance = s.resonance;
        s.resonance = clampf(s.resonance, 0.0f, 0.99f);
        s.h = 2.0f * s.resonance;
    }
    const float hp = (input - s.h * s.z0 - s.z1 - s.z2) / (1.0f + s.g);
    const float bp1 = s.g * hp + s.z0;
    s.z0 = om_tanhf_u(bp1);
    const float bp2 = s.g * bp1 + s.z1;
    s.z1 = bp2;
    const float lp = s.g * bp2 + s.z2;
    s.z2 = lp;
    return lp;
}

/* =================================================================== IirLowpass == */
/* oscen-lib/src/filters/iir_lowpass/mod.rs:15-166: JUCE-style biquad low-pass, Direct Form II Transposed;
   coefficients are refreshed from (cutoff, q) every 32 frames (frame_counter == 0), not per sample */
struct IirLpState {
    float b0, b1, b2, a1, a2, v1, v2;
    uint32_t counter;
};
OSC_HD void iirlp_update_coefficients(IirLpState &s, float sample_rate, float cutoff, float q_in) {
    const float nyquist = sample_rate * 0.5f - FLT_EPSILON;
    const float freq = clampf(cutoff, 20.0f, nyquist);
    const float q = fmaxf(q_in, 0.01f);
    const float n = 1.0f / OSC_TANF(kPi * freq / sample_rate);
    const float n_squared = n * n;
    const float c1 = 1.0f / (1.0f + 1.0f / q * n + n_squared);
    s.b0 = c1;
    s.b1 = c1 * 2.0f;
    s.b2 = c1;
    s.a1 = c1 * 2.0f * (1.0f - n_squared);
    s.a2 = c1 * (1.0f - 1.0f / q * n + n_squared);
}
/* IirLowpass::new(cutoff, q) + set_sample_rate + prepare(): mod.rs:64-75,150-153 */
OSC_HD void iirlp_init(IirLpState &s, float sample_rate, float cutoff, float q) {
    s.v1 = 0.0f; s.v2 = 0.0f; s.counter = 0u;
    iirlp_update_coefficients(s, sample_rate, cutoff, q);
}
OSC_D float iirlp_process(IirLpState &s, float input_in, float cutoff, float q, float sample_rate) {
    if (s.counter == 0u) iirlp_update_coefficients(s, sample_rate, cutoff, q);
    s.counter = (s.counter + 1u) % 32u;
    const float input = fabsf(input_in) < 1e-15f ? 0.0f : input_in;
    const float output = s.b0 * input + s.v1;
    s.v1 = s.b1 * input - s.a1 * output + s.v2;
    s.v2 = s.b2 * input - s.a2 * output;
    if (fabsf(s.v1) < 1e-15f) s.v1 = 0.0f;
    if (fabsf(s.v2) < 1e-15f) s.v2 = 0.0f;
    return output;
}

/* =================================================================== Oscillator == */
/* oscen-lib/src/oscillators/mod.rs:7-76: the naive (not band-limited) oscillator; `%` is fmodf */
enum : int { OSC_SINE = 0, OSC_SQUARE = 1, OSC_SAW = 2 };
OSC_HD float fmod1(float x) {                          /* x % 1.0: exact, sign of the dividend */
    float r = x - truncf(x);
    if (r == 0.0f) r = copysignf(0.0f, x);
    return r;
}
template <int WAVE>
OSC_D float osc_process(float &phase, float frequency_in, float frequency_mod, float amplitude, float sample_rate) {
    const float frequency = frequency_in * (1.0f + frequency_mod);
    const float p = fmod1(phase);
    float w;
    if (WAVE == OSC_SINE) {
        w = om_sinf_u(p * 2.0f * kPi);
    } else if (WAVE == OSC_SQUARE) {
        w = p < 0.5f ? 1.0f : -1.0f;
    } else {
        const float transition_width = 0.1f;
        const float raw_saw = 2.0f * p - 1.0f;
        if (p > (1.0f - transition_width / 2.0f)) {
            const float t = (p - (1.0f - transition_width / 2.0f)) / (transition_width / 2.0f);
            w = -1.0f + (1.0f - t * t) * (raw_saw + 1.0f);
        } else {
            w = raw_saw;
        }
    }
    const float out = w * amplitude;
    phase += frequency / sample_rate;
    phase = fmod1(phase);
    return out;
}

/* ======================================================================== Delay == */
/* oscen-lib/src/delay/mod.rs:5-86 over ring_buffer/mod.rs:29-202 (PowerOfTwo mode).  The ring lives in the
   voice's SoA state rows (`ring[slot * stride]`, stride = n_voices: slot-major, so voices that share a delay
   length touch consecutive addresses).  `cap` is a power of two: the reference's capacity
   next_power_of_two(min((2 * sr) as usize, 88200)) in general, or a smaller ring when the delay is a constant
   whole number of samples (the exact-sample read path never looks further back than that). */
struct DelayState {
    uint32_t write_pos, counter;
    float output;                                      /* persists: a feedback edge reads last frame's value */
};
OSC_HD uint32_t delay_reference_capacity(float sample_rate) {
    uint32_t size = f32_as_u32(2.0f * sample_rate);
    if (size > 88200u) size = 88200u;
    if (size < 1u) size = 1u;
    uint32_t cap = 1u;
    while (cap < size) cap <<= 1;
    return cap;
}
/* RingBuffer::get (ring_buffer/mod.rs:167-202) */
OSC_D float ring_get(const float *ring, size_t stride, uint32_t cap, uint32_t write_pos, float offset) {
    const uint32_t mask = cap - 1u;
    const float nn = fmaxf(offset, 0.0f);
    const float fr = nn - truncf(nn);
    if (fr < 1e-6f || (1.0f - fr) < 1e-6f) {
        const float r = roundf(nn);
        /* `as usize` saturates; (offset_samples % capacity) with a power-of-two capacity */
        const unsigned long long os = r >= 18446744073709551616.0f ? 0xFFFFFFFFFFFFFFFFull : (unsigned long long)r;
        const uint32_t idx = ((write_pos + cap) - (uint32_t)(os & (unsigned long long)mask) - 1u) & mask;
        return ring[(size_t)idx * stride];
    }
    const float n = (float)cap;
    const float rp0 = (float)write_pos - nn - 1.0f;
    const float rp = fmodf(fmodf(rp0, n) + n, n);
    const uint32_t i = (uint32_t)rp;
    const float f = rp - truncf(rp);
    const float v0 = ring[(size_t)((i - 1u) & mask) * stride], v1 = ring[(size_t)(i & mask) * stride];
    const float v2 = ring[(size_t)((i + 1u) & mask) * stride], v3 = ring[(size_t)((i + 2u) & mask) * stride];
    if (cap < 4u) return fmaf(v1, 1.0f - f, v2 * f);     /* get_linear: a.mul_add(1.0 - f, b * f) */
    const float c0 = v1;
    const float c1 = 0.5f * (v2 - v0);
    const float c2 = v0 - 2.5f * v1 + 2.0f * v2 - 0.5f * v3;
    const float c3 = 0.5f * (v3 - v0) + 1.5f * (v1 - v2);
    return c0 + f * (c1 + f * (c2 + f * c3));
}
/* Delay::process.  CONNECTED_D / CONNECTED_F: the value endpoint is assigned from a graph input every frame, so the
   clamp of apply_parameter_updates (every 32 frames) only holds for that frame; an unconnected endpoint keeps its
   clamped constructor value (passed in already clamped). */
template <bool CONNECTED_D, bool CONNECTED_F>
OSC_D float delay_process(DelayState &s, float *ring, size_t stride, uint32_t cap, float max_delay,
                          float input, float delay_samples, float feedback) {
    if (s.counter == 0u) {
        if (CONNECTED_D) delay_samples = clampf(delay_samples, 0.0f, max_delay);
        if (CONNECTED_F) feedback = clampf(feedback, 0.0f, 0.99f);
    }
    s.counter = (s.counter + 1u) % 32u;
    const float delayed = ring_get(ring, stride, cap, s.write_pos, delay_samples);
    ring[(size_t)s.write_pos * stride] = input + delayed * feedback;
    s.write_pos = (s.write_pos + 1u) & (cap - 1u);
    s.output = delayed;
    return delayed;
}

/* ===================================================================== HardClip == */
/* examples/oversampled-saturator/src/main.rs:55-62 */
OSC_D float hardclip_process(float input) {
    const float driven = input * 1.5f;
    return clampf(driven, -0.7f, 0.7f);
}

/* ======================================================= Halfband2xDownStage::step == */
/* oscen-lib/src/resample/sinc_fir.rs:113-138, taps resample/coeffs.rs:17-27.
   `hist` is this thread's ring stored twice, 48 slots (stride `stride` floats: one column of a
   [48][threads] shared-memory tile, conflict-free); `head` in [0,24) as in the reference. */
constexpr float kHb0 = -3.8558514e-5f, kHb1 = 1.2218465e-3f, kHb2 = -7.2854808e-3f,
                kHb3 = 2.6409210e-2f, kHb4 = -7.8128843e-2f, kHb5 = 3.0782697e-1f;
constexpr float kHbCenter = 0.4999897f;
OSC_D float hb_down_step(float *hist, int stride, uint32_t &head, float x0, float x1) {
    /* the ring is stored twice (slot i and i + 24), so x[2m - d] = hist[head + 23 - d] needs no
       modulo: 13 shared-memory loads with immediate offsets instead of 13 mod-24 index chains */
    head = (head + 1u == 24u) ? 0u : head + 1u; hist[head * stride] = x0; hist[(head + 24u) * stride] = x0;
    head = (head + 1u == 24u) ? 0u : head + 1u; hist[head * stride] = x1; hist[(head + 24u) * stride] = x1;
    const float *h = hist + (head + 23u) * stride;          /* h[-d*stride] = x[2m - d], d = 0..22 */
#define OSC_HB_AT(d) h[-(int)(d) * stride]
    float acc = OSC_HB_AT(11) * kHbCenter;
    acc = acc + (OSC_HB_AT(0) + OSC_HB_AT(22)) * kHb0;
    acc = acc + (OSC_HB_AT(2) + OSC_HB_AT(20)) * kHb1;
    acc = acc + (OSC_HB_AT(4) + OSC_HB_AT(18)) * kHb2;
    acc = acc + (OSC_HB_AT(6) + OSC_HB_AT(16)) * kHb3;
    acc = acc + (OSC_HB_AT(8) + OSC_HB_AT(14)) * kHb4;
    acc = acc + (OSC_HB_AT(10) + OSC_HB_AT(12)) * kHb5;
#undef OSC_HB_AT
    return acc;
}

/* ========================================================= Halfband2xUpStage::step == */
/* oscen-lib/src/resample/sinc_fir.rs:54-79.  `hist`: this thread's 12-slot ring stored twice
   (24 slots, stride `stride`), `head` in [0,12) as in the reference: x[n - d] = hist[head + 12 - d]. */
OSC_D void hb_up_step(float *hist, int stride, uint32_t &head, float x, float &out0, float &out1) {
    head = (head + 1u == 12u) ? 0u : head + 1u; hist[head * stride] = x; hist[(head + 12u) * stride] = x;
    const float *h = hist + (head + 12u) * stride;          /* h[-d*stride] = x[n - d], d = 0..11 */
#define OSC_UP_AT(d) h[-(int)(d) * stride]
    out1 = OSC_UP_AT(5) * (2.0f * kHbCenter);
    float acc = 0.0f;
    acc = acc + (OSC_UP_AT(0) + OSC_UP_AT(11)) * kHb0;
    acc = acc + (OSC_UP_AT(1) + OSC_UP_AT(10)) * kHb1;
    acc = acc + (OSC_UP_AT(2) + OSC_UP_AT(9)) * kHb2;
    acc = acc + (OSC_UP_AT(3) + OSC_UP_AT(8)) * kHb3;
    acc = acc + (OSC_UP_AT(4) + OSC_UP_AT(7)) * kHb4;
    acc = acc + (OSC_UP_AT(5) + OSC_UP_AT(6)) * kHb5;
#undef OSC_UP_AT
    out0 = acc * 2.0f;
}

/* Halfband2xDownStage::step (sinc_fir.rs:113-138) with the 24-sample history in REGISTERS.  A step pushes xs[0] = x[2m]
   and xs[1] = x[2m + 1] and reads x[2m - d] for the even delays d = 0, 2, .., 22 and for d = 11 only (every other odd tap
   of a halfband filter is zero), so the history splits into two polyphase shift registers: a[k] = x[2m - 2k] (the FIRST
   sample of the step k steps ago, k = 0..11) and b[k] = x[2m + 1 - 2k] (its second sample, k = 0..6; b[6] is the centre
   tap d = 11, b[0] the sample this step only stores).  Shifting by one entry per step is free once the frame loop is
   unrolled (the compiler renames registers inside the unrolled body and moves 19 values per ring at the back edge); the
   shared-memory ring this replaces cost 13 LDS + 4 STS per step and capped the kernel at 5 CTAs per SM
   (profiles/r02_ncu_sat4x_smem_rings.md; every [sinc] cross-rate edge of the hand-written and the generated kernels
   uses this form since round 2).  Same taps, same order of additions as hb_down_step: the same bits. */
struct HbDownRegs { float a[12]; float b[7]; };
OSC_D float hb_down_step_r(HbDownRegs &r, float x0, float x1) {
#pragma unroll
    for (int i = 11; i > 0; --i) r.a[i] = r.a[i - 1];
    r.a[0] = x0;
#pragma unroll
    for (int i = 6; i > 0; --i) r.b[i] = r.b[i - 1];
    r.b[0] = x1;
    float acc = r.b[6] * kHbCenter;                          /* d = 11 */
    acc = acc + (r.a[0] + r.a[11]) * kHb0;                   /* d = 0, 22 */
    acc = acc + (r.a[1] + r.a[10]) * kHb1;
    acc = acc + (r.a[2] + r.a[9]) * kHb2;
    acc = acc + (r.a[3] + r.a[8]) * kHb3;
    acc = acc + (r.a[4] + r.a[7]) * kHb4;
    acc = acc + (r.a[5] + r.a[6]) * kHb5;                    /* d = 10, 12 */
    return acc;
}
/* the stage's 24 state rows are chronological (row 23 = newest = x[2m + 1], row 22 - d = x[2m - d]).  The rows of the odd
   delays 13 .. 21 are never read by a step (now or later: samples only get older) and are not maintained. */
OSC_D void hb_down_load(HbDownRegs &r, const float *state, uint32_t first, uint32_t n, uint32_t v) {
#pragma unroll
    for (int k = 0; k < 12; ++k) r.a[k] = state[(size_t)(first + 22u - 2u * k) * n + v];
#pragma unroll
    for (int k = 0; k < 7; ++k) r.b[k] = state[(size_t)(first + 23u - 2u * k) * n + v];
}
OSC_D void hb_down_store(const HbDownRegs &r, float *state, uint32_t first, uint32_t n, uint32_t v) {
#pragma unroll
    for (int k = 0; k < 12; ++k) state[(size_t)(first + 22u - 2u * k) * n + v] = r.a[k];
#pragma unroll
    for (int k = 0; k < 7; ++k) state[(size_t)(first + 23u - 2u * k) * n + v] = r.b[k];
}

/* Halfband2xUpStage::step with the 12-sample history in registers: h[d] = x[n - d].  The state rows are chronological
   (row 11 = newest), as hb_up_step's ring is stored. */
struct HbUpRegs { float h[12]; };
OSC_D void hb_up_step_r(HbUpRegs &r, float x, float &out0, float &out1) {
#pragma unroll
    for (int i = 11; i > 0; --i) r.h[i] = r.h[i - 1];
    r.h[0] = x;
    out1 = r.h[5] * (2.0f * kHbCenter);
    float acc = 0.0f;
    acc = acc + (r.h[0] + r.h[11]) * kHb0;
    acc = acc + (r.h[1] + r.h[10]) * kHb1;
    acc = acc + (r.h[2] + r.h[9]) * kHb2;
    acc = acc + (r.h[3] + r.h[8]) * kHb3;
    acc = acc + (r.h[4] + r.h[7]) * kHb4;
    acc = acc + (r.h[5] + r.h[6]) * kHb5;
    out0 = acc * 2.0f;
}
OSC_D void hb_up_load(HbUpRegs &r, const float *state, uint32_t first, uint32_t n, uint32_t v) {
#pragma unroll
    for (int k = 0; k < 12; ++k) r.h[k] = state[(size_t)(first + 11u - (uint32_t)k) * n + v];
}
OSC_D void hb_up_store(const HbUpRegs &r, float *state, uint32_t first, uint32_t n, uint32_t v) {
#pragma unroll
    for (int k = 0; k < 12; ++k) state[(size_t)(first + 11u - (uint32_t)k) * n + v] = r.h[k];
}

/* ===================================================================== IirHalfband2x == */
/* oscen-lib/src/resample/halfband_iir.rs:24-143, coeffs.rs:47-54 */
constexpr float kIirA0 = 0.1355741f, kIirA1 = 0.6975849f, kIirB0 = 0.4253804f, kIirB1 = 0.9055601f;
constexpr float kDenormalThreshold = 1e-15f;               /* halfband_iir.rs:22 */
OSC_D float flush_denormal(float x) { return fabsf(x) < kDenormalThreshold ? 0.0f : x; }   /* frame.rs:87-93 */
struct Allpass1 { float x_prev, y_prev; };
/* Allpass1::step (halfband_iir.rs:45-57): the returned y is the pre-snap value */
OSC_D float allpass1_step(Allpass1 &s, float a, float x) {
    const float y = (x - s.y_prev) * a + s.x_prev;
    s.x_prev = flush_denormal(x);
    s.y_prev = flush_denormal(y);
    return y;
}
struct IirHb { Allpass1 a0, a1, b0, b1; float prev_odd_in; };
OSC_D void iir_hb_step_up(IirHb &h, float x, float &out0, float &out1) {         /* halfband_iir.rs:104-116 */
    out0 = allpass1_step(h.a1, kIirA1, allpass1_step(h.a0, kIirA0, x));
    out1 = allpass1_step(h.b1, kIirB1, allpass1_step(h.b0, kIirB0, x));
}
OSC_D float iir_hb_step_down(IirHb &h, float x0, float x1) {                      /* halfband_iir.rs:124-136 */
    const float a = allpass1_step(h.a1, kIirA1, allpass1_step(h.a0, kIirA0, x0));
    const float b = allpass1_step(h.b1, kIirB1, allpass1_step(h.b0, kIirB0, h.prev_odd_in));
    h.prev_odd_in = x1;
    return (a + b) * 0.5f;
}

/* ====================================================================== Tremolo == */
/* examples/electric-piano/src/tremolo.rs:41-61 */
OSC_D void tremolo_process(float &phase, float input, float rate, float depth, float sample_rate,
                           float &out_l, float &out_r) {
    const float lfo = om_sinf_u(phase * 2.0f * kPi);
    const float scaled_depth = depth / 3.0f;
    const float pan = 0.5f + lfo * scaled_depth;
    out_l = input * pan;
    out_r = input * (1.0f - pan);
    const float inc = rate / sample_rate;
    phase = fractf(phase + inc);
}

/* the same node split for a tile of frames: the phase recurrence is serial in time (the caller runs `phase =
   fractf(phase + rate / sample_rate)` per frame), the outputs of a frame depend only on that frame's phase */
OSC_D void tremolo_output(float phase, float input, float depth, float &out_l, float &out_r) {
    const float lfo = om_sinf_u(phase * 2.0f * kPi);
    const float scaled_depth = depth / 3.0f;
    const float pan = 0.5f + lfo * scaled_depth;
    out_l = input * pan;
    out_r = input * (1.0f - pan);
}

/* ============================================ AmplitudeSource / OscillatorBank as graph nodes == */
/* examples/electric-piano/src/electric_piano_voice.rs:6-352.  These are the per-node forms a DSL-compiled graph uses
   (codegen.hpp): each node's state lives in its SoA rows in HBM and is walked IN PLACE, frame by frame (`S` = the row of
   field 0 for this voice, rows `n` floats apart), and `amplitudes: [f32; 32]` is the one array-typed stream of the node
   library — a 32-float local array between the two nodes.  The hand-written epiano_kernel (kernel_others.cuh) renders the
   same arithmetic with the state in registers across four lanes per voice and is ~20x faster; this form exists so that
   ElectricPianoGraph compiles from its own `graph!` text.  Same operations in the same order as the oracle. */
__constant__ float kEpVel0[32] = { 0.02f, 0.05f };
__constant__ float kEpVel127[32] = {
    0.150869f, 0.385766f, 0.215543f, 0.117811f, 0.100411f, 0.0128637f, 0.0288844f, 0.00243388f,
    0.00963092f, 0.0035634f, 0.00256945f, 0.00184799f, 0.000399878f, 0.000660576f, 3.00995e-05f,
    0.00021866f, 9.33705e-05f, 0.000177973f, 0.0002545f, 0.000323602f, 0.000779045f, 0.000116569f,
    0.000772873f, 0.000364486f, 0.000248027f, 0.00018236f, 3.27292e-05f, 6.64988e-05f,
    0.0f, 0.0f, 0.0f, 0.0f };

constexpr uint32_t kEpHarmonics = 32u;
/* State words of the two array nodes, HOT words first: the words every frame reads AND writes (amplitudes, targets, the step
   counter; the unit oscillators).  The COLD words behind them are read-mostly — the per-note decay / release factors (read once
   per 64-frame interpolation), the rotation multipliers (read every frame, written when the frequency moves) — and are reached
   through a second (base, stride) pair, so that a generated kernel may keep only the hot words in shared memory for the block
   and leave the cold ones in their HBM rows, cached by L1 (codegen.hpp).  With one home for all words both pairs are equal. */
enum : uint32_t { AS_CUR = 0, AS_TGT = 32, AS_RELEASED = 64, AS_STEP = 65, AS_NHOT = 66, AS_DEC = 66, AS_REL = 98, AS_NSTATE = 130 };
enum : uint32_t { OB_ORE = 0, OB_OIM = 32, OB_LASTF = 64, OB_NHOT = 65, OB_MRE = 65, OB_MIM = 97, OB_NSTATE = 129 };
/* AmplitudeSource::on_gate -> trigger_note / release_note (electric_piano_voice.rs:230-318) */
OSC_D void ampsrc_on_gate(float *S, size_t n, float *C, size_t cn, float scalar, float brightness, float velocity_scaling, float decay_rate,
                          float harmonic_decay, float key_scaling, float release_rate) {
    if (scalar > 0.0f) {
        const float velocity = scalar;
        const float note = 60.0f;                                   /* note_pitch is never updated */
        const float base_decay_rate = (100.0f - decay_rate) / 40000.0f;
        const float harmonic_scaling = 1.0f - ((100.0f - harmonic_decay) / 200000.0f);
        const float scaling_multiplier = (48.0f - note) / 12.0f;
        const float key_scaling_factor = scaling_multiplier * (key_scaling * 0.02f);
        const float adjusted = (key_scaling_factor > 0.0f)
            ? 1.0f - (base_decay_rate / (1.0f + key_scaling_factor))
            : 1.0f - (base_decay_rate * (1.0f - key_scaling_factor));
        const float rv = 0.999f - ((100.0f - release_rate) / 1000.0f);
        float bs = -0.2f + (0.8f * (brightness * 0.01f));
        bs += velocity * velocity_scaling * 0.01f * 0.5f;
        float scaling = 1.0f;
        for (uint32_t i = 0; i < kEpHarmonics; ++i) {
            C[(size_t)(AS_DEC + i) * cn] = adjusted * scaling; scaling *= harmonic_scaling;
            C[(size_t)(AS_REL + i) * cn] = rv;
            float amp = (kEpVel127[i] * velocity) + (kEpVel0[i] * (1.0f - velocity));
            amp *= 1.0f + bs * (float)i;
            S[(size_t)(AS_CUR + i) * n] = amp;
        }
        S[(size_t)AS_RELEASED * n] = 0.0f;
    } else {
        S[(size_t)AS_RELEASED * n] = 1.0f;
    }
    S[(size_t)AS_STEP * n] = 0.0f;
}
/* AmplitudeSource::process (electric_piano_voice.rs:322-351): amps = this frame's `amplitudes` */
OSC_D void ampsrc_process(float *S, size_t n, const float *C, size_t cn, float *amps) {
    uint32_t step = (uint32_t)S[(size_t)AS_STEP * n];
    const bool released = S[(size_t)AS_RELEASED * n] != 0.0f;
    if (step == 0u)
        for (uint32_t i = 0; i < kEpHarmonics; ++i)
            S[(size_t)(AS_TGT + i) * n] = S[(size_t)(AS_CUR + i) * n] * C[(size_t)((released ? AS_REL : AS_DEC) + i) * cn];
    if (step < 64u) {
        const float t = (float)(step + 1u) / 64.0f;
        for (uint32_t i = 0; i < kEpHarmonics; ++i) {
            const float c = S[(size_t)(AS_CUR + i) * n] * (1.0f - t) + S[(size_t)(AS_TGT + i) * n] * t;
            S[(size_t)(AS_CUR + i) * n] = c; amps[i] = c;
        }
        step += 1u;
    } else {
        for (uint32_t i = 0; i < kEpHarmonics; ++i) { const float c = S[(size_t)(AS_TGT + i) * n]; S[(size_t)(AS_CUR + i) * n] = c; amps[i] = c; }
        step = 0u;
    }
    S[(size_t)AS_STEP * n] = (float)step;
}
/* OscillatorBank::on_gate (electric_piano_voice.rs:112-123): a note-on restarts the oscillators at zero phase */
OSC_D void oscbank_on_gate(float *S, size_t n, float scalar) {
    if (scalar > 0.0f)
        for (uint32_t i = 0; i < kEpHarmonics; ++i) { S[(size_t)(OB_ORE + i) * n] = 1.0f; S[(size_t)(OB_OIM + i) * n] = 0.0f; }
}
/* OscillatorBank::process (electric_piano_voice.rs:127-170) */
OSC_D float oscbank_process(float *S, size_t n, float *C, size_t cn, float frequency, float sample_rate, const float *amps) {
    if (frequency > 0.0f && !(fabsf(S[(size_t)OB_LASTF * n] - frequency) < 0.01f)) {      /* update_multipliers */
        S[(size_t)OB_LASTF * n] = frequency;
        const float nyquist = sample_rate * 0.5f;
        for (uint32_t i = 0; i < kEpHarmonics; ++i) {
            const float hf = frequency * (float)(i + 1u);
            float re = 1.0f, im = 0.0f;
            if (hf < nyquist) {
                const float angle = 2.0f * kPi * hf / sample_rate;
                re = om_cosf_u(angle); im = om_sinf_u(angle);
            }
            C[(size_t)(OB_MRE + i) * cn] = re; C[(size_t)(OB_MIM + i) * cn] = im;
            S[(size_t)(OB_ORE + i) * n] = 1.0f; S[(size_t)(OB_OIM + i) * n] = 0.0f;
        }
    }
    float sum = 0.0f;
    for (uint32_t i = 0; i < kEpHarmonics; ++i) {
        const float ore = S[(size_t)(OB_ORE + i) * n], oim = S[(size_t)(OB_OIM + i) * n];
        const float mre = C[(size_t)(OB_MRE + i) * cn], mim = C[(size_t)(OB_MIM + i) * cn];
        const float nr = ore * mre - oim * mim;
        const float ni = ore * mim + oim * mre;
        S[(size_t)(OB_ORE + i) * n] = nr; S[(size_t)(OB_OIM + i) * n] = ni;
        sum += ni * amps[i];
    }
    return sum * 3.0f;                                              /* output scaling */
}

/* ---- SamplePlayer, mono (oscen-lib/src/sample_player/mod.rs:104-127): loops the slot's buffer; an empty / never
   loaded slot is silence.  The swap-and-reset of `buf.take()` is the caller's `playhead = 0` on the first block after a
   load (engine_types.h AssetTable::reset). */
OSC_HD float sample_player_process(uint32_t &playhead, const float *buf, uint32_t len) {
    if (len == 0u) return 0.0f;
    const float y = buf[playhead];
    playhead += 1u;
    if (playhead >= len) playhead = 0u;
    return y;
}

}  // namespace oscen
