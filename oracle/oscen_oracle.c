/*
 * oscen_oracle.c — CPU ORACLE (test infrastructure; see oscen_oracle.h).
 *
 * Plain-C restatement of the reference's node process() bodies and of the code
 * the graph! macro generates around them.  Every function cites the reference
 * file:line it follows (paths relative to /root/reference).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp (see Makefile).
 * Rust semantics that matter:
 *   a*b+c is never fused; f32::clamp = two compares (NaN propagates);
 *   f32::max/min = fmaxf/fminf; `as u32` saturates; rem_euclid / fract spelled out.
 */
#include "oscen_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

static const float ORC_PI = 3.14159265358979323846f;   /* std::f32::consts::PI  */
static const float ORC_TAU = 6.28318530717958647692f;  /* std::f32::consts::TAU */

/* f32::clamp(self, lo, hi) */
static inline float clampf(float x, float lo, float hi) {
    if (x < lo) x = lo;
    if (x > hi) x = hi;
    return x;
}
/* `x as u32`: truncate toward zero, saturating, NaN -> 0 */
static inline uint32_t f32_as_u32(float x) {
    if (!(x == x)) return 0u;
    if (x <= 0.0f) return 0u;
    if (x >= 4294967296.0f) return 4294967295u;
    return (uint32_t)x;
}
/* f32::rem_euclid(1.0) */
static inline float rem_euclid1(float x) {
    float r = fmodf(x, 1.0f);
    return (r < 0.0f) ? r + 1.0f : r;
}
/* f32::fract */
static inline float fractf(float x) { return x - truncf(x); }
static inline uint32_t min_u32(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t max_u32(uint32_t a, uint32_t b) { return a > b ? a : b; }

/* ======================================================== ValueRampState == */
/* graph/types.rs:320-328 */
void orc_ramp_new(orc_ramp *r, float initial) {
    r->current = initial; r->target = initial; r->increment = 0.0f; r->frames_remaining = 0;
}
/* graph/types.rs:331-337 */
void orc_ramp_set_immediate(orc_ramp *r, float v) {
    r->current = v; r->target = v; r->increment = 0.0f; r->frames_remaining = 0;
}
/* graph/types.rs:340-349 */
void orc_ramp_set_with_ramp(orc_ramp *r, float target, uint32_t frames) {
    if (frames == 0) { orc_ramp_set_immediate(r, target); return; }
    r->target = target;
    r->increment = (target - r->current) / (float)frames;
    r->frames_remaining = frames;
}
/* graph/types.rs:354-366 */
int orc_ramp_tick(orc_ramp *r) {
    if (r->frames_remaining > 0) {
        r->frames_remaining -= 1;
        if (r->frames_remaining == 0) {
            r->current = r->target; r->increment = 0.0f;
            return 1;
        }
        r->current += r->increment;
    }
    return 0;
}
int orc_ramp_is_ramping(const orc_ramp *r) { return r->frames_remaining > 0; }

/* ============================================================ AdsrEnvelope == */
#define ADSR_MIN_TIME 1.0e-5f            /* adsr.rs:5 */
#define ADSR_CURVE_K 4.6051702f          /* adsr.rs:8  (4.605_170_2) */

/* adsr.rs:159-173 */
static void adsr_update_release_increment(orc_adsr *e) {
    if (e->samples_remaining == 0 || e->stage != ORC_RELEASE) { e->release_increment = 0.0f; return; }
    float current = clampf(e->level, 0.0f, 1.0f);
    e->release_increment = (current <= 0.0f) ? 0.0f : -current / (float)e->samples_remaining;
}
/* adsr.rs:117-134 */
static void adsr_recalculate_cached_steps(orc_adsr *e) {
    float sr = fmaxf(e->sample_rate, 1.0f);
    e->attack_samples = max_u32(f32_as_u32(fmaxf(e->attack, ADSR_MIN_TIME) * sr), 1u);
    e->decay_samples = max_u32(f32_as_u32(fmaxf(e->decay, ADSR_MIN_TIME) * sr), 1u);
    e->release_samples = max_u32(f32_as_u32(fmaxf(e->release, ADSR_MIN_TIME) * sr), 1u);
    e->attack_coeff = 1.0f - expf(-ADSR_CURVE_K / (float)e->attack_samples);
    e->decay_coeff = 1.0f - expf(-ADSR_CURVE_K / (float)e->decay_samples);
}
/* adsr.rs:92-115 */
static void adsr_update_sustain_level(orc_adsr *e) {
    e->sustain_level = clampf(e->sustain * e->velocity, 0.0f, 1.0f);
    adsr_recalculate_cached_steps(e);
    if (e->samples_remaining > 0) {
        if (e->stage == ORC_ATTACK)
            e->samples_remaining = max_u32(min_u32(e->samples_remaining, e->attack_samples), 1u);
        else if (e->stage == ORC_DECAY)
            e->samples_remaining = max_u32(min_u32(e->samples_remaining, e->decay_samples), 1u);
        else if (e->stage == ORC_RELEASE)
            e->samples_remaining = max_u32(min_u32(e->samples_remaining, e->release_samples), 1u);
    }
    if (e->stage == ORC_DECAY || e->stage == ORC_SUSTAIN) e->target_level = e->sustain_level;
    else if (e->stage == ORC_RELEASE) e->target_level = 0.0f;
    if (e->stage == ORC_RELEASE) adsr_update_release_increment(e);
}
/* adsr.rs:58-82 */
void orc_adsr_new(orc_adsr *e, float a, float d, float s, float r) {
    memset(e, 0, sizeof *e);
    e->attack = a; e->decay = d; e->sustain = s; e->release = r;
    e->stage = ORC_IDLE;
    e->sustain_level = clampf(s, 0.0f, 1.0f);
    e->velocity = 1.0f;
    e->sample_rate = 44100.0f;               /* SampleRate::default, graph/types.rs:259-264 */
    adsr_update_sustain_level(e);
}
void orc_adsr_set_sample_rate(orc_adsr *e, float sr) { e->sample_rate = sr; }
/* adsr.rs:277-279 */
void orc_adsr_prepare(orc_adsr *e) { adsr_update_sustain_level(e); }

static void adsr_complete_stage(orc_adsr *e);
/* adsr.rs:136-157 */
static void adsr_set_stage(orc_adsr *e, int stage, float target_level) {
    e->stage = stage;
    e->target_level = clampf(target_level, 0.0f, 1.0f);
    uint32_t samples = 0;
    if (stage == ORC_ATTACK) samples = e->attack_samples;
    else if (stage == ORC_DECAY) samples = e->decay_samples;
    else if (stage == ORC_RELEASE) samples = e->release_samples;
    if (samples == 0) {
        e->samples_remaining = 0;
        e->release_increment = 0.0f;
        e->level = e->target_level;
        if (!(stage == ORC_SUSTAIN || stage == ORC_IDLE)) adsr_complete_stage(e);
    } else {
        e->samples_remaining = samples;
        adsr_update_release_increment(e);
    }
}
/* adsr.rs:175-207 */
static void adsr_complete_stage(orc_adsr *e) {
    switch (e->stage) {
    case ORC_ATTACK:
        e->level = 1.0f;
        adsr_set_stage(e, ORC_DECAY, e->sustain_level);
        break;
    case ORC_DECAY:
        e->level = e->sustain_level; e->stage = ORC_SUSTAIN;
        e->samples_remaining = 0; e->release_increment = 0.0f;
        break;
    case ORC_RELEASE:
        e->level = 0.0f; e->stage = ORC_IDLE;
        e->samples_remaining = 0; e->release_increment = 0.0f;
        break;
    case ORC_SUSTAIN:
        e->level = e->sustain_level; e->samples_remaining = 0; e->release_increment = 0.0f;
        break;
    default: /* Idle */
        e->level = 0.0f; e->samples_remaining = 0; e->release_increment = 0.0f;
        break;
    }
}
/* adsr.rs:209-250 */
static void adsr_process_stage(orc_adsr *e) {
    switch (e->stage) {
    case ORC_ATTACK:
        if (e->samples_remaining > 0) {
            e->level += (1.0f - e->level) * e->attack_coeff;
            e->samples_remaining -= 1;
            e->level = clampf(e->level, 0.0f, 1.0f);
        }
        if (e->samples_remaining == 0) { e->level = 1.0f; adsr_complete_stage(e); }
        break;
    case ORC_DECAY:
        if (e->samples_remaining > 0) {
            e->level += (e->sustain_level - e->level) * e->decay_coeff;
            e->samples_remaining -= 1;
            e->level = clampf(e->level, 0.0f, 1.0f);
        }
        if (e->samples_remaining == 0) { e->level = e->sustain_level; adsr_complete_stage(e); }
        break;
    case ORC_RELEASE:
        if (e->samples_remaining > 0) {
            e->level += e->release_increment;
            e->samples_remaining -= 1;
            e->level = clampf(e->level, 0.0f, 1.0f);
        }
        if (e->samples_remaining == 0) { e->level = 0.0f; adsr_complete_stage(e); }
        break;
    case ORC_SUSTAIN: e->level = e->sustain_level; break;
    default: e->level = 0.0f; break;
    }
}
/* adsr.rs:252-273 (handle_gate_event) + :301-305 (on_gate) */
void orc_adsr_on_gate(orc_adsr *e, float velocity) {
    if (velocity > 0.0f) {
        e->velocity = clampf(velocity, 0.0f, 1.0f);
        adsr_update_sustain_level(e);
        if (e->attack <= ADSR_MIN_TIME) {
            e->level = 1.0f;
            adsr_set_stage(e, ORC_DECAY, e->sustain_level);
        } else {
            adsr_set_stage(e, ORC_ATTACK, 1.0f);
        }
    } else if (e->release <= ADSR_MIN_TIME) {
        e->stage = ORC_IDLE; e->level = 0.0f; e->samples_remaining = 0; e->release_increment = 0.0f;
    } else {
        adsr_set_stage(e, ORC_RELEASE, 0.0f);
    }
}
/* adsr.rs:84-90 (apply_parameters) + :281-291 (process) */
void orc_adsr_process(orc_adsr *e) {
    e->attack = fmaxf(e->attack, 0.0f);
    e->decay = fmaxf(e->decay, 0.0f);
    e->sustain = clampf(e->sustain, 0.0f, 1.0f);
    e->release = fmaxf(e->release, 0.0f);
    adsr_update_sustain_level(e);
    adsr_process_stage(e);
    e->output = e->level;
}

/* =============================================================== TptFilter == */
/* tpt/mod.rs:69-82 */
static void tpt_update_coefficients(orc_tpt *f, float sample_rate, float cutoff, float q) {
    float nyquist = sample_rate * 0.5f - FLT_EPSILON;
    float freq = clampf(cutoff, 20.0f, nyquist);
    float period = 0.5f / sample_rate;
    float fw = (2.0f * sample_rate) * tanf(2.0f * ORC_PI * freq * period) * period;
    float inv_q = 1.0f / q;
    f->h = 1.0f / (1.0f + inv_q * fw + fw * fw);
    f->g = fw;
    f->r = inv_q;
    f->k = f->g + f->r;
    f->current_cutoff = cutoff;
    f->current_q = q;
}
/* tpt/mod.rs:46-67 */
void orc_tpt_new(orc_tpt *f, float cutoff, float q) {
    memset(f, 0, sizeof *f);
    f->cutoff = cutoff; f->q = q; f->f_mod = 0.0f;
    f->current_cutoff = cutoff; f->current_q = q;
    f->sample_rate = 44100.0f;
    tpt_update_coefficients(f, 44100.0f, cutoff, q);
}
void orc_tpt_set_sample_rate(orc_tpt *f, float sr) { f->sample_rate = sr; }
/* tpt/mod.rs:129-131 */
void orc_tpt_prepare(orc_tpt *f) { tpt_update_coefficients(f, f->sample_rate, f->cutoff, f->q); }
/* tpt/mod.rs:84-123 */
void orc_tpt_process(orc_tpt *f) {
    float sample_rate = f->sample_rate;
    float nyquist = sample_rate * 0.5f - FLT_EPSILON;
    float max_cutoff = fminf(nyquist, 20000.0f);
    float cutoff_base = clampf(f->cutoff, 20.0f, max_cutoff);
    float q = clampf(f->q, 0.1f, 10.0f);
    float modulation = clampf(f->f_mod, -1.0f, 1.0f);
    float min_factor = 20.0f / cutoff_base;
    float max_factor = max_cutoff / cutoff_base;
    float factor = clampf(1.0f + modulation, min_factor, max_factor);
    float cutoff = clampf(cutoff_base * factor, 20.0f, max_cutoff);
    if (fabsf(cutoff - f->current_cutoff) > FLT_EPSILON || fabsf(q - f->current_q) > FLT_EPSILON)
        tpt_update_coefficients(f, sample_rate, cutoff, q);

    float high = (f->input - f->z[0] * f->k - f->z[1]) * f->h;
    float band = high * f->g + f->z[0];
    float low = band * f->g + f->z[1];
    f->z[0] = high * f->g + band;
    f->z[1] = band * f->g + low;
    f->output = low;
}

/* ====================================================== PolyBlepOscillator == */
/* oscillators/mod.rs:139-153 */
float orc_poly_blep(float t, float dt) {
    if (dt <= FLT_EPSILON) return 0.0f;
    if (t < dt) { float x = t / dt; return x + x - x * x - 1.0f; }
    if (t > 1.0f - dt) { float x = (t - 1.0f) / dt; return x * x + x + x + 1.0f; }
    return 0.0f;
}
/* oscillators/mod.rs:155-169 */
float orc_poly_blamp(float t, float dt) {
    if (dt <= FLT_EPSILON) return 0.0f;
    if (t < dt) { float x = t / dt - 1.0f; return -(x * x * x) / 3.0f; }
    if (t > 1.0f - dt) { float x = (t - 1.0f) / dt + 1.0f; return (x * x * x) / 3.0f; }
    return 0.0f;
}
/* oscillators/mod.rs:108-120 */
void orc_polyblep_new(orc_polyblep *o, float freq, float amp, int waveform) {
    memset(o, 0, sizeof *o);
    o->frequency = freq; o->amplitude = amp; o->pulse_width = 0.5f;
    o->waveform = waveform; o->sample_rate = 44100.0f;
}
void orc_polyblep_set_sample_rate(orc_polyblep *o, float sr) { o->sample_rate = sr; }
/* oscillators/mod.rs:176-232 */
void orc_polyblep_process(orc_polyblep *o) {
    float frequency = fmaxf(o->frequency * (1.0f + o->frequency_mod), 0.0f);
    float amplitude = o->amplitude;
    float pulse_width = clampf(o->pulse_width, 0.0001f, 0.9999f);
    float phase = rem_euclid1(o->phase + o->phase_mod);
    float freq_per_sample = frequency / fmaxf(o->sample_rate, FLT_EPSILON);
    float dt = fminf(freq_per_sample, 1.0f);
    if (pulse_width <= 0.0f) pulse_width = 0.0001f;

    float value;
    if (frequency >= o->sample_rate * 0.25f) {
        value = sinf(phase * ORC_TAU);
    } else {
        switch (o->waveform) {
        case ORC_WAVE_SINE: value = sinf(phase * ORC_TAU); break;
        case ORC_WAVE_SAW: {
            float y = 2.0f * phase - 1.0f;
            y -= orc_poly_blep(phase, dt);
            value = y;
        } break;
        case ORC_WAVE_SQUARE: {
            float y = (phase < pulse_width) ? 1.0f : -1.0f;
            y += orc_poly_blep(phase, dt);
            float t = rem_euclid1(phase + 1.0f - pulse_width);
            y -= orc_poly_blep(t, dt);
            value = y;
        } break;
        default: { /* Triangle */
            float y = 4.0f * phase;
            if (y >= 3.0f) y -= 4.0f;
            else if (y > 1.0f) y = 2.0f - y;
            float t1 = rem_euclid1(phase + 0.25f);
            float t2 = rem_euclid1(phase + 0.75f);
            value = y + 4.0f * dt * (orc_poly_blamp(t1, dt) - orc_poly_blamp(t2, dt));
        } break;
        }
    }
    value = value * amplitude;
    o->output = value;
    o->phase = rem_euclid1(o->phase + freq_per_sample);
}

/* naive Oscillator: oscillators/mod.rs:25-76 */
void orc_osc_new(orc_osc *o, float freq, float amp, int waveform) {
    memset(o, 0, sizeof *o);
    o->frequency = freq; o->amplitude = amp; o->waveform = waveform; o->sample_rate = 44100.0f;
}
void orc_osc_process(orc_osc *o) {
    float frequency = o->frequency * (1.0f + o->frequency_mod);
    float p = fmodf(o->phase, 1.0f);
    float w;
    if (o->waveform == 0) w = sinf(p * 2.0f * ORC_PI);
    else if (o->waveform == 1) w = (p < 0.5f) ? 1.0f : -1.0f;
    else {
        float transition_width = 0.1f;
        float raw_saw = 2.0f * p - 1.0f;
        if (p > (1.0f - transition_width / 2.0f)) {
            float t = (p - (1.0f - transition_width / 2.0f)) / (transition_width / 2.0f);
            w = -1.0f + (1.0f - t * t) * (raw_saw + 1.0f);
        } else w = raw_saw;
    }
    o->output = w * o->amplitude;
    o->phase += frequency / o->sample_rate;
    o->phase = fmodf(o->phase, 1.0f);
}

/* ============================================================ SamplePlayer == */
/* sample_player/mod.rs:77-84 */
void orc_sample_player_new(orc_sample_player *p) { memset(p, 0, sizeof *p); }
/* sample_player/mod.rs:36-49 with F = f32 (F::CHANNELS == 1): sc = 0 for a mono file, else min(0, src_ch - 1) = 0 */
void orc_sample_player_build_mono(float *dst, const float *interleaved, size_t frames, size_t channels) {
    for (size_t t = 0; t < frames; ++t) dst[t] = interleaved[t * channels];
}
void orc_sample_player_publish(orc_sample_player *p, const float *samples, size_t frames) {
    p->pending = samples; p->pending_len = frames; p->has_pending = 1;
}
/* sample_player/mod.rs:104-127 */
void orc_sample_player_process(orc_sample_player *p) {
    if (p->has_pending) {                       /* buf.take(): swap, reset the playhead */
        p->current = p->pending; p->current_len = p->pending_len; p->has_pending = 0;
        p->playhead = 0;
    }
    if (p->current_len == 0) { p->output = 0.0f; return; }
    p->output = p->current[p->playhead];
    p->playhead += 1;
    if (p->playhead >= p->current_len) p->playhead = 0;
}

/* ============================================================== FmOperator == */
/* fm_operator.rs:35-49 */
void orc_fmop_new(orc_fmop *o) {
    memset(o, 0, sizeof *o);
    o->sample_rate = 44100.0f; o->base_freq = 440.0f; o->ratio = 1.0f;
    o->envelope = 1.0f; o->level = 1.0f;
}
/* fm_operator.rs:58-76 */
void orc_fmop_process(orc_fmop *o) {
    float frequency = o->base_freq * o->ratio;
    float feedback_mod = o->prev_output * o->feedback;
    float total_phase_mod = o->phase_mod + feedback_mod;
    float phase_rad = (o->phase + total_phase_mod) * ORC_TAU;
    float output = sinf(phase_rad) * o->envelope * o->level;
    o->output = output;
    o->prev_output = output;
    float phase_inc = frequency / o->sample_rate;
    o->phase += phase_inc;
    o->phase = fractf(o->phase);
}

/* ============================================================== LP18Filter == */
/* lp18_filter.rs:44-60 */
void orc_lp18_new(orc_lp18 *f, float cutoff, float resonance) {
    memset(f, 0, sizeof *f);
    f->cutoff = cutoff; f->fmod = 0.0f;
    f->resonance = clampf(resonance, 0.0f, 0.99f);
    f->last_cutoff = cutoff; f->last_fmod = 0.0f; f->last_resonance = resonance;
    f->sample_rate = 44100.0f;
}
/* lp18_filter.rs:62-66 */
static void lp18_update_cutoff(orc_lp18 *f) {
    float modulated = f->cutoff + f->fmod;
    float fc = clampf(modulated / f->sample_rate, 0.001f, 0.33f);
    f->g = tanf(ORC_PI * fc);
}
/* lp18_filter.rs:68-70 */
static void lp18_update_resonance(orc_lp18 *f) { f->h = 2.0f * f->resonance; }
/* lp18_filter.rs:74-77 */
void orc_lp18_prepare(orc_lp18 *f) { lp18_update_cutoff(f); lp18_update_resonance(f); }
/* lp18_filter.rs:79-107 */
void orc_lp18_process(orc_lp18 *f) {
    if (f->cutoff != f->last_cutoff || f->fmod != f->last_fmod) {
        f->last_cutoff = f->cutoff; f->last_fmod = f->fmod;
        lp18_update_cutoff(f);
    }
    if (f->resonance != f->last_resonance) {
        f->last_resonance = f->resonance;
        f->resonance = clampf(f->resonance, 0.0f, 0.99f);
        lp18_update_resonance(f);
    }
    float hp = (f->input - f->h * f->z[0] - f->z[1] - f->z[2]) / (1.0f + f->g);
    float bp1 = f->g * hp + f->z[0];
    f->z[0] = tanhf(bp1);
    float bp2 = f->g * bp1 + f->z[1];
    f->z[1] = bp2;
    float lp = f->g * bp2 + f->z[2];
    f->z[2] = lp;
    f->output = lp;
}

/* ============================================================== IirLowpass == */
/* oscen-lib/src/filters/iir_lowpass/mod.rs:15-166 */
void orc_iirlp_new(orc_iirlp *f, float cutoff, float q) {
    memset(f, 0, sizeof *f);
    f->cutoff = cutoff; f->q = q; f->b0 = 1.0f;
    f->frames_per_update = 32; f->sample_rate = 44100.0f;
}
static void iirlp_update_coefficients(orc_iirlp *f, float sample_rate) {     /* mod.rs:84-101 */
    float nyquist = sample_rate * 0.5f - FLT_EPSILON;
    float freq = clampf(f->cutoff, 20.0f, nyquist);
    float q = fmaxf(f->q, 0.01f);
    float n = 1.0f / tanf(ORC_PI * freq / sample_rate);
    float n_squared = n * n;
    float c1 = 1.0f / (1.0f + 1.0f / q * n + n_squared);
    f->b0 = c1;
    f->b1 = c1 * 2.0f;
    f->b2 = c1;
    f->a1 = c1 * 2.0f * (1.0f - n_squared);
    f->a2 = c1 * (1.0f - 1.0f / q * n + n_squared);
}
void orc_iirlp_prepare(orc_iirlp *f) { iirlp_update_coefficients(f, f->sample_rate); }
void orc_iirlp_process(orc_iirlp *f) {                                        /* mod.rs:109-166 */
    if (f->frame_counter == 0) iirlp_update_coefficients(f, f->sample_rate);
    f->frame_counter = (f->frame_counter + 1) % f->frames_per_update;
    float input = f->input;
    if (fabsf(input) < 1e-15f) input = 0.0f;
    float output = f->b0 * input + f->v1;
    f->v1 = f->b1 * input - f->a1 * output + f->v2;
    f->v2 = f->b2 * input - f->a2 * output;
    if (fabsf(f->v1) < 1e-15f) f->v1 = 0.0f;
    if (fabsf(f->v2) < 1e-15f) f->v2 = 0.0f;
    f->output = output;
}

/* ============================================================ Delay / RingBuffer == */
/* oscen-lib/src/ring_buffer/mod.rs:29-202 (BufferMode::PowerOfTwo, the default) */
static size_t next_pow2(size_t x) { size_t p = 1; while (p < x) p <<= 1; return p; }
static void ring_alloc(orc_ring *r, size_t size) {
    free(r->buffer);
    r->capacity = next_pow2(size < 1 ? 1 : size);
    r->mask = r->capacity - 1;
    r->write_pos = 0;
    r->buffer = (float *)calloc(r->capacity, sizeof(float));
}
static void ring_push(orc_ring *r, float v) {
    r->buffer[r->write_pos] = v;
    r->write_pos = (r->write_pos + 1) & r->mask;
}
static float ring_read_pos(const orc_ring *r, float offset) {                 /* mod.rs:80-91 */
    float n = (float)r->capacity;
    float rp = (float)r->write_pos - offset - 1.0f;
    return fmodf(fmodf(rp, n) + n, n);
}
static float ring_get(const orc_ring *r, float offset) {                      /* mod.rs:167-202 */
    float nn = fmaxf(offset, 0.0f);
    float fr = nn - truncf(nn);
    if (fr < 1e-6f || (1.0f - fr) < 1e-6f) {
        float rd = roundf(nn);
        size_t os = rd >= 18446744073709551616.0f ? (size_t)-1 : (size_t)rd;  /* `as usize` saturates */
        size_t idx = ((r->write_pos + r->capacity) - (os % r->capacity) - 1) % r->capacity;
        return r->buffer[idx];
    }
    float rp = ring_read_pos(r, nn);
    size_t i = (size_t)rp;
    float f = rp - truncf(rp);
    if (r->capacity < 4) {                                                    /* get_linear, mod.rs:94-118 */
        float a = r->buffer[i], b = r->buffer[(i + 1) & r->mask];
        return fmaf(a, 1.0f - f, b * f);
    }
    float v0 = r->buffer[(i - 1) & r->mask], v1 = r->buffer[i], v2 = r->buffer[(i + 1) & r->mask], v3 = r->buffer[(i + 2) & r->mask];
    float c0 = v1;
    float c1 = 0.5f * (v2 - v0);
    float c2 = v0 - 2.5f * v1 + 2.0f * v2 - 0.5f * v3;
    float c3 = 0.5f * (v3 - v0) + 1.5f * (v1 - v2);
    return c0 + f * (c1 + f * (c2 + f * c3));
}
/* test hooks on the ring itself (ring_buffer/tests.rs) */
void orc_ring_init(orc_ring *r, size_t size) { memset(r, 0, sizeof *r); ring_alloc(r, size); }
void orc_ring_push(orc_ring *r, float v) { ring_push(r, v); }
float orc_ring_get(const orc_ring *r, float offset) { return ring_get(r, offset); }
void orc_ring_free(orc_ring *r) { free(r->buffer); r->buffer = NULL; }

/* oscen-lib/src/delay/mod.rs:5-86 */
void orc_delay_new(orc_delay *d, float delay_samples, float feedback) {
    memset(d, 0, sizeof *d);
    d->delay_samples = delay_samples; d->feedback = feedback;
    d->frames_per_update = 32; d->sample_rate = 44100.0f;
    ring_alloc(&d->ring, 1024);
}
void orc_delay_prepare(orc_delay *d) {                                         /* mod.rs:59-69 */
    float target = 2.0f * d->sample_rate;
    size_t size = target >= 18446744073709551616.0f ? (size_t)-1 : (target > 0.0f ? (size_t)target : 0);
    if (size > 88200) size = 88200;
    ring_alloc(&d->ring, size);
}
void orc_delay_process(orc_delay *d) {                                         /* mod.rs:47-85 */
    if (d->frame_counter == 0) {
        float max_delay = (float)d->ring.capacity - 1.0f;
        d->delay_samples = clampf(d->delay_samples, 0.0f, max_delay);
        d->feedback = clampf(d->feedback, 0.0f, 0.99f);
    }
    d->frame_counter = (d->frame_counter + 1) % d->frames_per_update;
    float delayed = ring_get(&d->ring, d->delay_samples);
    ring_push(&d->ring, d->input + delayed * d->feedback);
    d->output = delayed;
}
void orc_delay_free(orc_delay *d) { free(d->ring.buffer); d->ring.buffer = NULL; }

/* ===================================================== halfband FIR cascade == */
/* resample/coeffs.rs:17-27 */
static const float HB23_HALF[6] = {
    -3.8558514e-5f, 1.2218465e-3f, -7.2854808e-3f, 2.6409210e-2f, -7.8128843e-2f, 3.0782697e-1f,
};
static const float HB23_CENTER = 0.4999897f;
#define HB_GROUP_DELAY 11      /* coeffs.rs:34 */

/* sinc_fir.rs:113-138 (Halfband2xDownStage::step) */
static float hb_down_step(orc_hb_down *s, float x0, float x1) {
    const size_t cap = 24;
    s->head = (s->head + 1) % cap; s->history[s->head] = x0;
    s->head = (s->head + 1) % cap; s->history[s->head] = x1;
#define HB_AT(d) (s->history[(s->head + cap - 1 - (size_t)(d)) % cap])
    float acc = HB_AT(11) * HB23_CENTER;
    for (int k = 0; k < 6; ++k) {
        float left = HB_AT(2 * k);
        float right = HB_AT(22 - 2 * k);
        acc = acc + (left + right) * HB23_HALF[k];
    }
#undef HB_AT
    return acc;
}
/* sinc_fir.rs:54-79 (Halfband2xUpStage::step) */
static void hb_up_step(orc_hb_up *s, float x, float out[2]) {
    const size_t cap = 12;
    s->head = (s->head + 1) % cap; s->history[s->head] = x;
#define UP_AT(d) (s->history[(s->head + cap - (size_t)(d)) % cap])
    out[1] = UP_AT(5) * (2.0f * HB23_CENTER);
    float acc = 0.0f;
    for (int k = 0; k < 6; ++k) {
        float left = UP_AT(k);
        float right = UP_AT(11 - k);
        acc = acc + (left + right) * HB23_HALF[k];
    }
#undef UP_AT
    out[0] = acc * 2.0f;
}
static int log2_pow2(int n) { int s = 0; while ((1 << s) < n) ++s; return s; }
/* sinc_fir.rs:215-224 */
void orc_sinc_down_new(orc_sinc_down *d, int n) {
    memset(d, 0, sizeof *d); d->n = n; d->n_stages = log2_pow2(n);
}
void orc_sinc_down_reset(orc_sinc_down *d) { int n = d->n; orc_sinc_down_new(d, n); }
/* sinc_fir.rs:233-249 */
float orc_sinc_down_process(orc_sinc_down *d, const float *xs) {
    float buf[8] = {0};
    for (int i = 0; i < d->n; ++i) buf[i] = xs[i];
    int len = d->n;
    for (int s = 0; s < d->n_stages; ++s) {
        float next[8] = {0};
        int half = len / 2;
        for (int i = 0; i < half; ++i) next[i] = hb_down_step(&d->stages[s], buf[2 * i], buf[2 * i + 1]);
        len = half;
        for (int i = 0; i < len; ++i) buf[i] = next[i];
    }
    return buf[0];
}
/* sinc_fir.rs:250-260 */
int orc_sinc_down_latency(const orc_sinc_down *d) {
    return d->n_stages == 0 ? 0 : HB_GROUP_DELAY * ((1 << d->n_stages) - 1);
}
/* sinc_fir.rs:153-162 */
void orc_sinc_up_new(orc_sinc_up *u, int n) {
    memset(u, 0, sizeof *u); u->n = n; u->n_stages = log2_pow2(n);
}
/* sinc_fir.rs:171-191 */
void orc_sinc_up_process(orc_sinc_up *u, float x, float *out) {
    float buf[8] = {0}, next[8] = {0};
    int len = 1; buf[0] = x;
    for (int s = 0; s < u->n_stages; ++s) {
        for (int i = 0; i < len; ++i) {
            float pair[2]; hb_up_step(&u->stages[s], buf[i], pair);
            next[2 * i] = pair[0]; next[2 * i + 1] = pair[1];
        }
        len *= 2;
        for (int i = 0; i < len; ++i) buf[i] = next[i];
    }
    for (int i = 0; i < u->n; ++i) out[i] = buf[i];
}
int orc_sinc_up_latency(const orc_sinc_up *u) {
    return u->n_stages == 0 ? 0 : HB_GROUP_DELAY * ((1 << u->n_stages) - 1);
}

/* ================================================= IIR halfband cascade == */
/* resample/coeffs.rs:47-54 */
static const float IIR_A_BETAS[2] = { 0.1355741f, 0.6975849f };
static const float IIR_B_BETAS[2] = { 0.4253804f, 0.9055601f };
#define IIR_GROUP_DELAY 2
#define DENORMAL_THRESHOLD 1e-15f              /* halfband_iir.rs:22 */
/* frame.rs:87-93 */
static inline float flush_denormal(float x, float thr) { return fabsf(x) < thr ? 0.0f : x; }
/* halfband_iir.rs:45-57 (Allpass1::step): returns the pre-snap y */
static inline float allpass1_step(orc_allpass1 *s, float x) {
    float y = (x - s->y_prev) * s->a + s->x_prev;
    s->x_prev = x; s->y_prev = y;
    s->x_prev = flush_denormal(s->x_prev, DENORMAL_THRESHOLD);
    s->y_prev = flush_denormal(s->y_prev, DENORMAL_THRESHOLD);
    return y;
}
static void iir_hb_new(orc_iir_hb *h) {            /* halfband_iir.rs:84-96 */
    memset(h, 0, sizeof *h);
    h->branch_a[0].a = IIR_A_BETAS[0]; h->branch_a[1].a = IIR_A_BETAS[1];
    h->branch_b[0].a = IIR_B_BETAS[0]; h->branch_b[1].a = IIR_B_BETAS[1];
}
/* halfband_iir.rs:104-116 */
static void iir_hb_step_up(orc_iir_hb *h, float x, float out[2]) {
    float a = x; for (int k = 0; k < 2; ++k) a = allpass1_step(&h->branch_a[k], a);
    float b = x; for (int k = 0; k < 2; ++k) b = allpass1_step(&h->branch_b[k], b);
    out[0] = a; out[1] = b;
}
/* halfband_iir.rs:124-136 */
static float iir_hb_step_down(orc_iir_hb *h, float x0, float x1) {
    float a = x0; for (int k = 0; k < 2; ++k) a = allpass1_step(&h->branch_a[k], a);
    float b = h->prev_odd_in; for (int k = 0; k < 2; ++k) b = allpass1_step(&h->branch_b[k], b);
    h->prev_odd_in = x1;
    return (a + b) * 0.5f;
}
void orc_iir_up_new(orc_iir_up *u, int n) {
    memset(u, 0, sizeof *u); u->n = n; u->n_stages = log2_pow2(n);
    for (int s = 0; s < u->n_stages; ++s) iir_hb_new(&u->stages[s]);
}
/* halfband_iir.rs:173-190 */
void orc_iir_up_process(orc_iir_up *u, float x, float *out) {
    float buf[8] = {0}, next[8] = {0};
    int len = 1; buf[0] = x;
    for (int s = 0; s < u->n_stages; ++s) {
        for (int i = 0; i < len; ++i) {
            float pair[2]; iir_hb_step_up(&u->stages[s], buf[i], pair);
            next[2 * i] = pair[0]; next[2 * i + 1] = pair[1];
        }
        len *= 2;
        for (int i = 0; i < len; ++i) buf[i] = next[i];
    }
    for (int i = 0; i < u->n; ++i) out[i] = buf[i];
}
int orc_iir_up_latency(const orc_iir_up *u) {      /* halfband_iir.rs:192-203 */
    return u->n_stages == 0 ? 0 : IIR_GROUP_DELAY * ((1 << u->n_stages) - 1);
}
void orc_iir_down_new(orc_iir_down *d, int n) { orc_iir_up_new(d, n); }
/* halfband_iir.rs:235-251 */
float orc_iir_down_process(orc_iir_down *d, const float *xs) {
    float buf[8] = {0};
    for (int i = 0; i < d->n; ++i) buf[i] = xs[i];
    int len = d->n;
    for (int s = 0; s < d->n_stages; ++s) {
        float next[8] = {0};
        int half = len / 2;
        for (int i = 0; i < half; ++i) next[i] = iir_hb_step_down(&d->stages[s], buf[2 * i], buf[2 * i + 1]);
        len = half;
        for (int i = 0; i < len; ++i) buf[i] = next[i];
    }
    return buf[0];
}
int orc_iir_down_latency(const orc_iir_down *d) { return orc_iir_up_latency(d); }

/* ======================================================== linear / latch == */
void orc_linear_up_new(orc_linear_up *u, int n) { u->prev = 0.0f; u->n = n; }
/* linear.rs:23-32 */
void orc_linear_up_process(orc_linear_up *u, float x, float *out) {
    float n_inv = 1.0f / (float)u->n;
    float delta = x - u->prev;
    for (int i = 0; i < u->n; ++i) out[i] = u->prev + delta * ((float)i * n_inv);
    u->prev = x;
}
/* linear.rs:59-66 */
float orc_linear_down_process(int n, const float *xs) {
    float acc = 0.0f;
    for (int i = 0; i < n; ++i) acc = acc + xs[i];
    return acc * (1.0f / (float)n);
}
int orc_linear_up_latency(int n) { return n; }                 /* linear.rs:33-36 */
int orc_linear_down_latency(int n) { return (n - 1) / 2; }     /* linear.rs:67-70 */
/* latch.rs:17-23, 45-49 */
void orc_latch_up_process(int n, float x, float *out) { for (int i = 0; i < n; ++i) out[i] = x; }
float orc_latch_down_process(int n, const float *xs) { (void)n; return xs[0]; }

/* ==================================================================== MIDI == */
/* midi.rs:69-72 */
float orc_midi_note_to_freq(uint8_t note) {
    float semitone_offset = (float)note - 69.0f;
    return 440.0f * powf(2.0f, semitone_offset / 12.0f);
}
/* midi.rs:147-171 */
int orc_midi_parse(const uint8_t *data, size_t len, uint8_t *note, float *velocity) {
    if (len < 3) return 0;
    uint8_t status = data[0] & 0xF0;
    *note = data[1];
    uint8_t vel = data[2];
    if (status == 0x80) return 2;
    if (status == 0x90) {
        if (vel == 0) return 2;
        *velocity = clampf((float)vel / 127.0f, 0.0f, 1.0f);
        return 1;
    }
    return 0;
}
/* voice_allocator.rs:44-54 */
void orc_voice_alloc_new(orc_voice_alloc *a, int num_voices) {
    memset(a, 0, sizeof *a);
    a->num_voices = num_voices;
    for (int i = 0; i < ORC_MAX_VOICES; ++i) a->note[i] = -1;
}
/* voice_allocator.rs:57-91 */
int orc_voice_alloc_allocate(orc_voice_alloc *a, uint8_t note) {
    for (int i = 0; i < a->num_voices; ++i) {
        if (!a->active[i]) {
            a->active[i] = 1; a->released[i] = 0; a->note[i] = note;
            a->age[i] = a->current_age; a->current_age += 1;
            return i;
        }
    }
    /* min_by_key((release_priority, age)): first minimum wins */
    int best = 0; int have = 0; uint32_t best_p = 0, best_age = 0;
    for (int i = 0; i < a->num_voices; ++i) {
        uint32_t p = a->released[i] ? 0u : 1u;
        if (!have || p < best_p || (p == best_p && a->age[i] < best_age)) {
            have = 1; best = i; best_p = p; best_age = a->age[i];
        }
    }
    a->active[best] = 1; a->released[best] = 0; a->note[best] = note;
    a->age[best] = a->current_age; a->current_age += 1;
    return best;
}
/* voice_allocator.rs:94-100 */
int orc_voice_alloc_find(const orc_voice_alloc *a, uint8_t note) {
    for (int i = 0; i < a->num_voices; ++i)
        if (a->active[i] && !a->released[i] && a->note[i] == (int)note) return i;
    return -1;
}
/* voice_allocator.rs:103-110 */
void orc_voice_alloc_release(orc_voice_alloc *a, int idx) {
    if (idx >= 0 && idx < a->num_voices) { a->released[idx] = 1; a->note[idx] = -1; }
}

/* ======================================================== electric piano ==== */
#define EP_H 32                       /* electric_piano_voice.rs:5 */
#define EP_INTERP 64                  /* electric_piano_voice.rs:6 */
/* electric_piano_voice.rs:10-48 */
static const float EP_VEL0[EP_H] = { 0.02f, 0.05f };
static const float EP_VEL127[EP_H] = {
    0.150869f, 0.385766f, 0.215543f, 0.117811f, 0.100411f, 0.0128637f, 0.0288844f, 0.00243388f,
    0.00963092f, 0.0035634f, 0.00256945f, 0.00184799f, 0.000399878f, 0.000660576f, 3.00995e-05f,
    0.00021866f, 9.33705e-05f, 0.000177973f, 0.0002545f, 0.000323602f, 0.000779045f, 0.000116569f,
    0.000772873f, 0.000364486f, 0.000248027f, 0.00018236f, 3.27292e-05f, 6.64988e-05f,
    0.0f, 0.0f, 0.0f, 0.0f,
};
typedef struct {            /* AmplitudeSource, electric_piano_voice.rs:172-215 */
    float brightness, velocity_scaling, decay_rate, harmonic_decay, key_scaling, release_rate;
    float amplitudes[EP_H], current_value[EP_H], target_value[EP_H], decay[EP_H], release[EP_H];
    int released; float note_pitch, velocity; uint32_t interpolation_step;
} ep_amp;
typedef struct {            /* OscillatorBank, electric_piano_voice.rs:78-101 */
    float frequency, output;
    float amplitudes[EP_H];
    float osc_re[EP_H], osc_im[EP_H], mul_re[EP_H], mul_im[EP_H];
    float last_frequency, sample_rate;
} ep_bank;
/* electric_piano_voice.rs:217-239 */
static void ep_amp_new(ep_amp *a) {
    memset(a, 0, sizeof *a);
    a->brightness = 30.0f; a->velocity_scaling = 50.0f; a->decay_rate = 90.0f;
    a->harmonic_decay = 70.0f; a->key_scaling = 50.0f; a->release_rate = 40.0f;
    a->note_pitch = 60.0f; a->interpolation_step = EP_INTERP;
}
/* electric_piano_voice.rs:241-266 (get_decay) */
static void ep_get_decay(const ep_amp *a, float note, float *decay) {
    float base_decay_rate = (100.0f - a->decay_rate) / 40000.0f;
    float harmonic_scaling = 1.0f - ((100.0f - a->harmonic_decay) / 200000.0f);
    float scaling_multiplier = (48.0f - note) / 12.0f;
    float key_scaling_factor = scaling_multiplier * (a->key_scaling * 0.02f);
    float adjusted = (key_scaling_factor > 0.0f)
        ? 1.0f - (base_decay_rate / (1.0f + key_scaling_factor))
        : 1.0f - (base_decay_rate * (1.0f - key_scaling_factor));
    float scaling = 1.0f;
    for (int i = 0; i < EP_H; ++i) { decay[i] = adjusted * scaling; scaling *= harmonic_scaling; }
}
/* electric_piano_voice.rs:276-292 (get_initial_amplitudes) */
static void ep_get_initial(const ep_amp *a, float velocity, float *amps) {
    for (int i = 0; i < EP_H; ++i)
        amps[i] = (EP_VEL127[i] * velocity) + (EP_VEL0[i] * (1.0f - velocity));
    float bs = -0.2f + (0.8f * (a->brightness * 0.01f));
    bs += velocity * a->velocity_scaling * 0.01f * 0.5f;
    for (int i = 0; i < EP_H; ++i) amps[i] *= 1.0f + bs * (float)i;
}
/* electric_piano_voice.rs:294-306, 309-320 (on_gate) */
static void ep_amp_on_gate(ep_amp *a, float scalar) {
    if (scalar > 0.0f) {
        a->velocity = scalar;
        ep_get_decay(a, a->note_pitch, a->decay);
        float rv = 0.999f - ((100.0f - a->release_rate) / 1000.0f);     /* get_release :268-272 */
        for (int i = 0; i < EP_H; ++i) a->release[i] = rv;
        ep_get_initial(a, scalar, a->current_value);
        a->released = 0; a->interpolation_step = 0;
    } else {
        a->released = 1; a->interpolation_step = 0;
    }
}
/* electric_piano_voice.rs:322-351 */
static void ep_amp_process(ep_amp *a) {
    if (a->interpolation_step == 0) {
        const float *m = a->released ? a->release : a->decay;
        for (int i = 0; i < EP_H; ++i) a->target_value[i] = a->current_value[i] * m[i];
    }
    if (a->interpolation_step < EP_INTERP) {
        float t = (float)(a->interpolation_step + 1) / (float)EP_INTERP;
        for (int i = 0; i < EP_H; ++i)
            a->current_value[i] = a->current_value[i] * (1.0f - t) + a->target_value[i] * t;
        a->interpolation_step += 1;
    } else {
        memcpy(a->current_value, a->target_value, sizeof a->current_value);
        a->interpolation_step = 0;
    }
    memcpy(a->amplitudes, a->current_value, sizeof a->amplitudes);
}
/* electric_piano_voice.rs:103-114 */
static void ep_bank_new(ep_bank *b) {
    memset(b, 0, sizeof *b);
    b->frequency = 440.0f; b->sample_rate = 44100.0f; b->last_frequency = 0.0f;
    for (int i = 0; i < EP_H; ++i) { b->osc_re[i] = 1.0f; b->mul_re[i] = 1.0f; }
}
/* electric_piano_voice.rs:116-123 */
static void ep_bank_on_gate(ep_bank *b, float scalar) {
    if (scalar > 0.0f) for (int i = 0; i < EP_H; ++i) { b->osc_re[i] = 1.0f; b->osc_im[i] = 0.0f; }
}
/* electric_piano_voice.rs:127-151 */
static void ep_bank_update_multipliers(ep_bank *b, float f) {
    if (fabsf(b->last_frequency - f) < 0.01f) return;
    b->last_frequency = f;
    float nyquist = b->sample_rate * 0.5f;
    for (int i = 0; i < EP_H; ++i) {
        float hn = (float)(i + 1);
        float hf = f * hn;
        if (hf < nyquist) {
            float angle = 2.0f * ORC_PI * hf / b->sample_rate;
            b->mul_re[i] = cosf(angle); b->mul_im[i] = sinf(angle);
        } else { b->mul_re[i] = 1.0f; b->mul_im[i] = 0.0f; }
    }
    for (int i = 0; i < EP_H; ++i) { b->osc_re[i] = 1.0f; b->osc_im[i] = 0.0f; }
}
/* electric_piano_voice.rs:154-170 (+ Complex::mul :65-71) */
static void ep_bank_process(ep_bank *b) {
    if (b->frequency > 0.0f) ep_bank_update_multipliers(b, b->frequency);
    float sum = 0.0f;
    for (int i = 0; i < EP_H; ++i) {
        float nr = b->osc_re[i] * b->mul_re[i] - b->osc_im[i] * b->mul_im[i];
        float ni = b->osc_re[i] * b->mul_im[i] + b->osc_im[i] * b->mul_re[i];
        b->osc_re[i] = nr; b->osc_im[i] = ni;
        sum += b->osc_im[i] * b->amplitudes[i];
    }
    b->output = sum * 3.0f;
}
/* tremolo.rs:8-61 */
typedef struct { float input, rate, depth, out[2], phase, sample_rate; } ep_tremolo;
static void ep_tremolo_new(ep_tremolo *t) {
    memset(t, 0, sizeof *t); t->rate = 5.0f; t->depth = 0.5f; t->sample_rate = 44100.0f;
}
static void ep_tremolo_process(ep_tremolo *t) {
    float input = t->input, rate = t->rate, depth = t->depth;
    float lfo = sinf(t->phase * 2.0f * ORC_PI);
    float scaled_depth = depth / 3.0f;
    float pan = 0.5f + lfo * scaled_depth;
    t->out[0] = input * pan; t->out[1] = input * (1.0f - pan);
    float inc = rate / t->sample_rate;
    t->phase = fractf(t->phase + inc);
}

/* ================================================================== graphs == */
enum { K_FM = 0, K_TWIN = 1, K_SAT1 = 2, K_SAT4 = 3, K_EPIANO = 4, K_SYNTH = 5, K_SYNTH_ADSR = 6,
       K_CLIP_REF = 7, K_CLIP_OS = 8 /* .. 19: 8 + policy*3 + (log2 factor - 1) */, K_COUNT = 20 };
static int kind_is_clip(uint32_t k) { return k >= K_CLIP_REF && k < K_COUNT; }
static int clip_policy(uint32_t k) { return (int)(k - K_CLIP_OS) / 3; }
static int clip_factor(uint32_t k) { return k == K_CLIP_REF ? 1 : 2 << ((k - K_CLIP_OS) % 3); }
#define MAX_IN 28
#define MAXQ 32            /* MAX_STATIC_EVENTS_PER_ENDPOINT, graph/types.rs:18 */
static const int N_INPUTS[K_COUNT] = { 28, 3, 2, 2, 9, 4, 8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
/* default values of the value inputs (fm_voice.rs:10-48; nih-twin-peaks lib.rs:21-23;
   oversampled-saturator main.rs:71 saw(2000,.6); electric-piano main.rs:39-46 + 440;
   README.md:28-31) */
static const float DEFAULTS[K_COUNT][MAX_IN] = {
    { 440.0f, 3.0f, 0.5f, 0.0f, 0.01f, 0.1f, 0.7f, 0.3f,   2.0f, 0.5f, 0.0f, 0.01f, 0.1f, 0.7f, 0.3f,
      1.0f, 0.01f, 0.2f, 0.8f, 0.5f,   0.0f,   2000.0f, 0.707f, 0.01f, 0.2f, 0.5f, 0.3f, 0.0f },
    { 1000.0f, 1900.0f, 0.54f },
    { 2000.0f, 0.6f },
    { 2000.0f, 0.6f },
    { 440.0f, 30.0f, 50.0f, 90.0f, 70.0f, 50.0f, 40.0f, 0.3f, 5.0f },
    { 5.0f, 0.2f, 440.0f, 1200.0f },
    { 5.0f, 0.2f, 440.0f, 1200.0f, 0.01f, 0.2f, 0.7f, 0.3f },
};
enum { FM_FREQ = 0, FM_OP3_RATIO, FM_OP3_LEVEL, FM_OP3_FB, FM_OP3_A, FM_OP3_D, FM_OP3_S, FM_OP3_R,
       FM_OP2_RATIO, FM_OP2_LEVEL, FM_OP2_FB, FM_OP2_A, FM_OP2_D, FM_OP2_S, FM_OP2_R,
       FM_OP1_RATIO, FM_OP1_A, FM_OP1_D, FM_OP1_S, FM_OP1_R, FM_ROUTE,
       FM_CUTOFF, FM_RES, FM_F_A, FM_F_D, FM_F_S, FM_F_R, FM_ENV_AMT };

typedef struct {          /* FMVoice: examples/fm-synth/src/fm_voice.rs:6-156 */
    orc_adsr env3, env2, env1, env_filter;
    float feg_in, feg_gain, feg_out;           /* filter_env_gain = Gain::new(0.0) */
    float cm_in, cm_value, cm_out;             /* cutoff_mod = AddValue::new(2000.0) */
    orc_fmop op3, op2, op1;
    float xf_in, xf_mix, xf_a, xf_b;           /* op3_route = Crossfade */
    float mx_a, mx_b, mx_out;                  /* op1_mod_mixer = Mixer */
    orc_tpt filter;
    float og_in, og_gain, og_out;              /* output_gain = Gain::new(0.3) */
    float audio_out;
} fm_voice;

static void fm_voice_new(fm_voice *v) {
    memset(v, 0, sizeof *v);
    orc_adsr_new(&v->env3, 0.01f, 0.1f, 0.7f, 0.3f);
    orc_adsr_new(&v->env2, 0.01f, 0.1f, 0.7f, 0.3f);
    orc_adsr_new(&v->env1, 0.01f, 0.2f, 0.8f, 0.5f);
    orc_adsr_new(&v->env_filter, 0.01f, 0.2f, 0.5f, 0.3f);
    v->feg_gain = 0.0f; v->cm_value = 2000.0f;
    orc_fmop_new(&v->op3); orc_fmop_new(&v->op2); orc_fmop_new(&v->op1);
    orc_tpt_new(&v->filter, 2000.0f, 0.707f);
    v->og_gain = 0.3f;
}
/* generated init(): set_sample_rate on every node, then prepare() (codegen/mod.rs:1335-1382) */
static void fm_voice_init(fm_voice *v, float sr) {
    orc_adsr_set_sample_rate(&v->env3, sr); orc_adsr_set_sample_rate(&v->env2, sr);
    orc_adsr_set_sample_rate(&v->env1, sr); orc_adsr_set_sample_rate(&v->env_filter, sr);
    v->op3.sample_rate = sr; v->op2.sample_rate = sr; v->op1.sample_rate = sr;
    orc_tpt_set_sample_rate(&v->filter, sr);
    orc_adsr_prepare(&v->env3); orc_adsr_prepare(&v->env2);
    orc_adsr_prepare(&v->env1); orc_adsr_prepare(&v->env_filter);
    orc_tpt_prepare(&v->filter);
}
static inline void fm_env(orc_adsr *e, const float *gate, int ng, float a, float d, float s, float r) {
    e->attack = a; e->decay = d; e->sustain = s; e->release = r;
    for (int i = 0; i < ng; ++i) orc_adsr_on_gate(e, gate[i]);     /* process_event_inputs */
    orc_adsr_process(e);
}
/* one frame of the generated FMVoice::process() — SURVEY.md Appendix A;
   a valid Kahn order of ir/lower.rs:1015-1085 over fm_voice.rs:83-155 */
static void fm_voice_process(fm_voice *v, const float *in, const float *gate, int ng) {
    fm_env(&v->env3, gate, ng, in[FM_OP3_A], in[FM_OP3_D], in[FM_OP3_S], in[FM_OP3_R]);
    fm_env(&v->env2, gate, ng, in[FM_OP2_A], in[FM_OP2_D], in[FM_OP2_S], in[FM_OP2_R]);
    fm_env(&v->env1, gate, ng, in[FM_OP1_A], in[FM_OP1_D], in[FM_OP1_S], in[FM_OP1_R]);
    fm_env(&v->env_filter, gate, ng, in[FM_F_A], in[FM_F_D], in[FM_F_S], in[FM_F_R]);
    /* filter_env_gain (gain/mod.rs:30-35) */
    v->feg_in = v->env_filter.output; v->feg_gain = in[FM_ENV_AMT];
    v->feg_out = v->feg_in * v->feg_gain;
    /* cutoff_mod (add_value.rs:33-35) */
    v->cm_in = v->feg_out; v->cm_value = in[FM_CUTOFF];
    v->cm_out = v->cm_in + v->cm_value;
    /* op3 */
    v->op3.base_freq = in[FM_FREQ]; v->op3.ratio = in[FM_OP3_RATIO]; v->op3.feedback = in[FM_OP3_FB];
    v->op3.envelope = v->env3.output; v->op3.level = in[FM_OP3_LEVEL];
    orc_fmop_process(&v->op3);
    /* op3_route (crossfade.rs:39-44) */
    v->xf_in = v->op3.output; v->xf_mix = in[FM_ROUTE];
    { float mix = clampf(v->xf_mix, 0.0f, 1.0f); float x = v->xf_in;
      v->xf_a = x * (1.0f - mix); v->xf_b = x * mix; }
    /* op2 */
    v->op2.base_freq = in[FM_FREQ]; v->op2.ratio = in[FM_OP2_RATIO]; v->op2.feedback = in[FM_OP2_FB];
    v->op2.envelope = v->env2.output; v->op2.level = in[FM_OP2_LEVEL]; v->op2.phase_mod = v->xf_a;
    orc_fmop_process(&v->op2);
    /* op1_mod_mixer (mixer.rs:32-34) */
    v->mx_a = v->op2.output; v->mx_b = v->xf_b; v->mx_out = v->mx_a + v->mx_b;
    /* op1 (feedback, level never connected: stay 0.0 / 1.0) */
    v->op1.base_freq = in[FM_FREQ]; v->op1.ratio = in[FM_OP1_RATIO];
    v->op1.envelope = v->env1.output; v->op1.phase_mod = v->mx_out;
    orc_fmop_process(&v->op1);
    /* filter */
    v->filter.cutoff = v->cm_out; v->filter.input = v->op1.output; v->filter.q = in[FM_RES];
    orc_tpt_process(&v->filter);
    /* output_gain */
    v->og_in = v->filter.output; v->og_out = v->og_in * v->og_gain;
    v->audio_out = v->og_out;
}

typedef struct {          /* TwinPeaksGraph: examples/nih-twin-peaks/src/lib.rs:15-48 */
    orc_lp18 a, b; float audio_out;
} twin_voice;
static void twin_new(twin_voice *v) {
    orc_lp18_new(&v->a, 1000.0f, 0.54f); orc_lp18_new(&v->b, 1900.0f, 0.54f); v->audio_out = 0.0f;
}
static void twin_init(twin_voice *v, float sr) {
    v->a.sample_rate = sr; v->b.sample_rate = sr; orc_lp18_prepare(&v->a); orc_lp18_prepare(&v->b);
}
static void twin_process(twin_voice *v, const float *in, float audio_in) {
    v->a.input = audio_in; v->a.cutoff = in[0]; v->a.resonance = in[2]; orc_lp18_process(&v->a);
    v->b.input = audio_in; v->b.cutoff = in[1]; v->b.resonance = in[2]; orc_lp18_process(&v->b);
    v->audio_out = v->a.output - v->b.output;      /* compound source expr, codegen/mod.rs:433-472 */
}

typedef struct {          /* SatGraph_{1x,4x}: examples/oversampled-saturator/src/main.rs:64-80 */
    orc_polyblep osc; float clip_in, clip_out; orc_sinc_down down; int factor; float audio_out;
} sat_voice;
static void sat_new(sat_voice *v, int factor) {
    memset(v, 0, sizeof *v);
    orc_polyblep_new(&v->osc, 2000.0f, 0.6f, ORC_WAVE_SAW);
    orc_sinc_down_new(&v->down, factor); v->factor = factor;
}
/* set_sample_rate: inner (`* N`) nodes get sr*N (emit_struct.rs:578-581); prepare resets resamplers */
static void sat_init(sat_voice *v, float sr) {
    orc_polyblep_set_sample_rate(&v->osc, sr * (float)v->factor);
    orc_sinc_down_reset(&v->down);
}
/* main.rs:55-62 */
static inline float hardclip(float x) { float driven = x * 1.5f; return clampf(driven, -0.7f, 0.7f); }
/* multirate body, codegen/emit_frame.rs:114-176 */
static void sat_process(sat_voice *v, const float *in) {
    v->osc.frequency = in[0]; v->osc.amplitude = in[1];
    if (v->factor == 1) {
        orc_polyblep_process(&v->osc);
        v->clip_in = v->osc.output; v->clip_out = hardclip(v->clip_in);
        v->audio_out = v->clip_out;
        return;
    }
    float buf[8];
    for (int i = 0; i < v->factor; ++i) {
        orc_polyblep_process(&v->osc);
        v->clip_in = v->osc.output; v->clip_out = hardclip(v->clip_in);
        buf[i] = v->clip_out;
    }
    v->audio_out = orc_sinc_down_process(&v->down, buf);
}

typedef struct {          /* ClipRef / ClipOversampled{,Iir}: oscen-lib/tests/multirate_graph.rs:282-325;
                             HardClip here is clamp(+-0.7) with no drive (multirate_graph.rs:35-40) */
    int policy, factor;
    orc_sinc_up sup; orc_sinc_down sdown; orc_iir_up iup; orc_iir_down idown; orc_linear_up lup;
    float clip_in, clip_out, out;
} clip_voice;
static void clip_new(clip_voice *v, int policy, int factor) {
    memset(v, 0, sizeof *v); v->policy = policy; v->factor = factor;
    orc_sinc_up_new(&v->sup, factor); orc_sinc_down_new(&v->sdown, factor);
    orc_iir_up_new(&v->iup, factor); orc_iir_down_new(&v->idown, factor);
    orc_linear_up_new(&v->lup, factor);
}
/* multirate body, codegen/emit_frame.rs:114-176: up-edge warm-up (emit_frame.rs:254-307), N inner
   ticks {write, process, capture} (:388-461), down finalize (:474-514) */
static void clip_process(clip_voice *v, float input) {
    if (v->factor == 1) { v->clip_in = input; v->clip_out = clampf(v->clip_in, -0.7f, 0.7f); v->out = v->clip_out; return; }
    float up[8], down[8];
    switch (v->policy) {
    case ORC_POLICY_SINC: orc_sinc_up_process(&v->sup, input, up); break;
    case ORC_POLICY_SINC_IIR: orc_iir_up_process(&v->iup, input, up); break;
    case ORC_POLICY_LINEAR: orc_linear_up_process(&v->lup, input, up); break;
    default: orc_latch_up_process(v->factor, input, up); break;
    }
    for (int i = 0; i < v->factor; ++i) {
        v->clip_in = up[i]; v->clip_out = clampf(v->clip_in, -0.7f, 0.7f); down[i] = v->clip_out;
    }
    switch (v->policy) {
    case ORC_POLICY_SINC: v->out = orc_sinc_down_process(&v->sdown, down); break;
    case ORC_POLICY_SINC_IIR: v->out = orc_iir_down_process(&v->idown, down); break;
    case ORC_POLICY_LINEAR: v->out = orc_linear_down_process(v->factor, down); break;
    default: v->out = orc_latch_down_process(v->factor, down); break;
    }
}

typedef struct {          /* ElectricPianoVoiceNode: electric_piano_voice.rs:362-402 */
    ep_amp amp; ep_bank bank; float output;
} ep_voice;
static void ep_voice_new(ep_voice *v) { ep_amp_new(&v->amp); ep_bank_new(&v->bank); v->output = 0.0f; }
static void ep_voice_init(ep_voice *v, float sr) { v->bank.sample_rate = sr; }
static void ep_voice_process(ep_voice *v, const float *in, const float *gate, int ng) {
    ep_amp *a = &v->amp;
    a->brightness = in[1]; a->velocity_scaling = in[2]; a->decay_rate = in[3];
    a->harmonic_decay = in[4]; a->key_scaling = in[5]; a->release_rate = in[6];
    for (int i = 0; i < ng; ++i) ep_amp_on_gate(a, gate[i]);
    ep_amp_process(a);
    v->bank.frequency = in[0];
    memcpy(v->bank.amplitudes, a->amplitudes, sizeof a->amplitudes);
    for (int i = 0; i < ng; ++i) ep_bank_on_gate(&v->bank, gate[i]);
    ep_bank_process(&v->bank);
    v->output = v->bank.output;
}

typedef struct {          /* README.md:24-53 `Synth` (+ amp ADSR -> Gain vca for kind 6,
                             wiring as benches/static_vs_runtime.rs:57-59) */
    orc_polyblep modulator, carrier; orc_tpt filter; orc_adsr amp_env;
    float vca_gain; int with_adsr; float audio_out;
} synth_voice;
static void synth_new(synth_voice *v, int with_adsr) {
    memset(v, 0, sizeof *v);
    orc_polyblep_new(&v->modulator, 5.0f, 0.2f, ORC_WAVE_SINE);
    orc_polyblep_new(&v->carrier, 440.0f, 0.5f, ORC_WAVE_SAW);
    orc_tpt_new(&v->filter, 1200.0f, 0.707f);
    orc_adsr_new(&v->amp_env, 0.01f, 0.2f, 0.7f, 0.3f);
    v->vca_gain = 1.0f; v->with_adsr = with_adsr;
}
static void synth_init(synth_voice *v, float sr) {
    orc_polyblep_set_sample_rate(&v->modulator, sr); orc_polyblep_set_sample_rate(&v->carrier, sr);
    orc_tpt_set_sample_rate(&v->filter, sr); orc_adsr_set_sample_rate(&v->amp_env, sr);
    orc_tpt_prepare(&v->filter); orc_adsr_prepare(&v->amp_env);
}
static void synth_process(synth_voice *v, const float *in, const float *gate, int ng) {
    v->modulator.frequency = in[0]; v->modulator.amplitude = in[1];
    orc_polyblep_process(&v->modulator);
    v->carrier.frequency = in[2]; v->carrier.frequency_mod = v->modulator.output;
    orc_polyblep_process(&v->carrier);
    v->filter.cutoff = in[3]; v->filter.input = v->carrier.output;
    orc_tpt_process(&v->filter);
    if (v->with_adsr) {
        fm_env(&v->amp_env, gate, ng, in[4], in[5], in[6], in[7]);
        v->vca_gain = v->amp_env.output;
        v->audio_out = v->filter.output * v->vca_gain;
    } else {
        v->audio_out = v->filter.output;
    }
}

/* ==================================================================== bank == */
typedef struct { uint32_t frame; uint32_t input_id; float scalar; uint32_t seq; } vev;
typedef struct { uint32_t frame; uint8_t bytes[3]; uint8_t len; uint32_t seq; } mev;
typedef struct {          /* MidiVoiceHandler state, midi.rs:40-66 */
    int current_note; float current_frequency;
} midi_handler;

struct orc_bank {
    uint32_t kind, n_voices, vpi, n_inst, channels;
    float sr;
    int n_threads;
    size_t voice_size;
    unsigned char *voices;                 /* n_voices * voice_size */
    float *in;                             /* [voice][MAX_IN] */
    orc_ramp ramps[MAX_IN]; uint8_t ramped[MAX_IN]; uint8_t uniform[MAX_IN];
    /* per-instance MIDI front (fm-synth lib.rs:72-90, electric-piano main.rs:55-75) */
    orc_voice_alloc *alloc; midi_handler *handlers;
    ep_tremolo *trem;                      /* per instance (EPIANO) */
    /* pending (not yet rendered) events */
    vev **pend; uint32_t *npend, *cpend;   /* per voice */
    mev **mpend; uint32_t *nmpend, *cmpend; /* per instance */
    uint32_t seq;
    float *stream_in; uint32_t stream_frames;
    uint32_t last_frames;
    float *voice_out;                      /* [512][n_voices] */
    float *inst_out;                       /* [512][n_inst][channels] */
};

static size_t kind_voice_size(uint32_t k) {
    switch (k) {
    case K_FM: return sizeof(fm_voice);
    case K_TWIN: return sizeof(twin_voice);
    case K_SAT1: case K_SAT4: return sizeof(sat_voice);
    case K_EPIANO: return sizeof(ep_voice);
    case K_SYNTH: case K_SYNTH_ADSR: return sizeof(synth_voice);
    default: return sizeof(clip_voice);
    }
}
#define VOICE(b, v) ((void *)((b)->voices + (size_t)(v) * (b)->voice_size))

orc_bank *orc_bank_create(uint32_t kind, uint32_t n_voices, uint32_t vpi, float sr) {
    if (kind >= K_COUNT || n_voices == 0 || vpi == 0 || n_voices % vpi != 0 || vpi > ORC_MAX_VOICES)
        return NULL;
    orc_bank *b = calloc(1, sizeof *b);
    b->kind = kind; b->n_voices = n_voices; b->vpi = vpi; b->n_inst = n_voices / vpi; b->sr = sr;
    b->channels = (kind == K_EPIANO) ? 2 : 1;
    b->n_threads = 1;
    b->voice_size = kind_voice_size(kind);
    b->voices = calloc(n_voices, b->voice_size);
    b->in = calloc((size_t)n_voices * MAX_IN, sizeof(float));
    b->alloc = calloc(b->n_inst, sizeof *b->alloc);
    b->handlers = calloc(n_voices, sizeof *b->handlers);
    b->trem = calloc(b->n_inst, sizeof *b->trem);
    b->pend = calloc(n_voices, sizeof *b->pend);
    b->npend = calloc(n_voices, sizeof *b->npend); b->cpend = calloc(n_voices, sizeof *b->cpend);
    b->mpend = calloc(b->n_inst, sizeof *b->mpend);
    b->nmpend = calloc(b->n_inst, sizeof *b->nmpend); b->cmpend = calloc(b->n_inst, sizeof *b->cmpend);
    b->voice_out = calloc((size_t)512 * n_voices, sizeof(float));
    b->inst_out = calloc((size_t)512 * b->n_inst * b->channels, sizeof(float));
    for (uint32_t i = 0; i < MAX_IN; ++i) { orc_ramp_new(&b->ramps[i], DEFAULTS[kind][i]); b->uniform[i] = 1; }
    for (uint32_t v = 0; v < n_voices; ++v) {
        memcpy(b->in + (size_t)v * MAX_IN, DEFAULTS[kind], sizeof(float) * MAX_IN);
        b->handlers[v].current_note = -1; b->handlers[v].current_frequency = 440.0f;
        void *p = VOICE(b, v);
        switch (kind) {                                 /* new() then init(sr) */
        case K_FM: fm_voice_new(p); fm_voice_init(p, sr); break;
        case K_TWIN: twin_new(p); twin_init(p, sr); break;
        case K_SAT1: sat_new(p, 1); sat_init(p, sr); break;
        case K_SAT4: sat_new(p, 4); sat_init(p, sr); break;
        case K_EPIANO: ep_voice_new(p); ep_voice_init(p, sr); break;
        case K_SYNTH: synth_new(p, 0); synth_init(p, sr); break;
        case K_SYNTH_ADSR: synth_new(p, 1); synth_init(p, sr); break;
        default: clip_new(p, kind == K_CLIP_REF ? 0 : clip_policy(kind), clip_factor(kind)); break;
        }
    }
    for (uint32_t i = 0; i < b->n_inst; ++i) {
        orc_voice_alloc_new(&b->alloc[i], (int)vpi);
        ep_tremolo_new(&b->trem[i]); b->trem[i].sample_rate = sr;
    }
    return b;
}
void orc_bank_destroy(orc_bank *b) {
    if (!b) return;
    for (uint32_t v = 0; v < b->n_voices; ++v) free(b->pend[v]);
    for (uint32_t i = 0; i < b->n_inst; ++i) free(b->mpend[i]);
    free(b->voices); free(b->in); free(b->alloc); free(b->handlers); free(b->trem);
    free(b->pend); free(b->npend); free(b->cpend); free(b->mpend); free(b->nmpend); free(b->cmpend);
    free(b->stream_in); free(b->voice_out); free(b->inst_out); free(b);
}
uint32_t orc_bank_channels(const orc_bank *b) { return b->channels; }
/* latency_samples(): Down edges report at the source rate / factor (codegen/emit_struct.rs:533-566);
   SincDownFir<4> = 33 high-rate samples / factor 4 (integer division, emit_struct.rs:549-553) */
uint32_t orc_bank_latency(const orc_bank *b) {
    if (b->kind == K_SAT4) return 33 / 4;
    if (kind_is_clip(b->kind) && b->kind != K_CLIP_REF) {
        const clip_voice *v = (const clip_voice *)b->voices;
        int lat;
        switch (v->policy) {
        case ORC_POLICY_SINC: lat = orc_sinc_down_latency(&v->sdown); break;
        case ORC_POLICY_SINC_IIR: lat = orc_iir_down_latency(&v->idown); break;
        case ORC_POLICY_LINEAR: lat = orc_linear_down_latency(v->factor); break;
        default: lat = 0; break;
        }
        return (uint32_t)(lat / v->factor);
    }
    return 0;
}
void orc_bank_set_threads(orc_bank *b, int n) { b->n_threads = n < 1 ? 1 : n; }

int orc_bank_set_value(orc_bank *b, uint32_t id, uint32_t first, uint32_t count, const float *v, int bc) {
    if (id >= (uint32_t)N_INPUTS[b->kind] || first + count > b->n_voices) return -1;
    const int all = bc && first == 0 && count == b->n_voices;
    if (b->ramped[id] && !all) return -2;          /* a ramped graph input is one ValueRampState */
    for (uint32_t i = 0; i < count; ++i) b->in[(size_t)(first + i) * MAX_IN + id] = bc ? v[0] : v[i];
    if (all) { orc_ramp_set_immediate(&b->ramps[id], v[0]); b->uniform[id] = 1; }   /* set_<v>_immediate */
    else b->uniform[id] = 0;
    return 0;
}
/* set_<v>_with_ramp, codegen/mod.rs:948-957 */
int orc_bank_set_value_ramp(orc_bank *b, uint32_t id, float target, uint32_t frames) {
    if (id >= (uint32_t)N_INPUTS[b->kind]) return -1;
    if (!b->uniform[id]) return -2;
    for (uint32_t v = 0; v < b->n_voices; ++v)                    /* the mirror image of orc_bank_push_events' refusal */
        for (uint32_t k = 0; k < b->npend[v]; ++k) if (b->pend[v][k].input_id == id) return -2;
    b->ramped[id] = 1;
    if (target != b->ramps[id].target) orc_ramp_set_with_ramp(&b->ramps[id], target, frames);
    return 0;
}
int orc_bank_push_events(orc_bank *b, const orc_event *ev, size_t n) {
    /* a ramped graph input is ONE ValueRampState read through .current every frame (emit_node.rs:317-321): a per-voice
       write to it has no counterpart in the generated struct — refused, like a per-voice orc_bank_set_value on it */
    for (size_t i = 0; i < n; ++i)
        if (ev[i].input_id != ORC_EV_GATE && ev[i].input_id < MAX_IN && b->ramped[ev[i].input_id]) return -2;
    /* ElectricPianoGraph: vibrato_intensity / vibrato_speed feed the instance's Tremolo (main.rs:88-90), not a voice */
    for (size_t i = 0; i < n; ++i)
        if (b->kind == K_EPIANO && (ev[i].input_id == 7u || ev[i].input_id == 8u)) return -3;
    for (size_t i = 0; i < n; ++i) {
        uint32_t v = ev[i].voice;
        if (v >= b->n_voices) return -1;
        if (ev[i].input_id == ORC_EV_GATE) {
            /* the voice's `gate` input is a 32-slot StaticEventQueue: try_push on a full queue drops
               (graph/types.rs:18,183-188) */
            uint32_t g = 0;
            for (uint32_t k = 0; k < b->npend[v]; ++k) if (b->pend[v][k].input_id == ORC_EV_GATE) ++g;
            if (g >= MAXQ) continue;
        }
        if (b->npend[v] == b->cpend[v]) {
            b->cpend[v] = b->cpend[v] ? b->cpend[v] * 2 : 4;
            b->pend[v] = realloc(b->pend[v], b->cpend[v] * sizeof(vev));
        }
        vev e = { ev[i].frame_offset, ev[i].input_id, ev[i].scalar, b->seq++ };
        b->pend[v][b->npend[v]++] = e;
    }
    return 0;
}
int orc_bank_push_midi(orc_bank *b, uint32_t inst, uint32_t frame, const uint8_t *bytes, uint32_t len) {
    if (inst >= b->n_inst) return -1;
    if (b->nmpend[inst] >= MAXQ) return 0;                /* try_push on a full queue: dropped */
    if (b->nmpend[inst] == b->cmpend[inst]) {
        b->cmpend[inst] = b->cmpend[inst] ? b->cmpend[inst] * 2 : 4;
        b->mpend[inst] = realloc(b->mpend[inst], b->cmpend[inst] * sizeof(mev));
    }
    mev e; memset(&e, 0, sizeof e);
    e.frame = frame; e.len = (uint8_t)(len > 3 ? 3 : len);     /* RawMidiMessage::new, midi.rs:15-22 */
    memcpy(e.bytes, bytes, e.len); e.seq = b->seq++;
    b->mpend[inst][b->nmpend[inst]++] = e;
    return 0;
}
int orc_bank_stream_in(orc_bank *b, uint32_t id, const float *src, uint32_t frames) {
    if ((b->kind != K_TWIN && !kind_is_clip(b->kind)) || id != 0 || frames > 512) return -1;
    free(b->stream_in);
    b->stream_in = malloc((size_t)frames * b->n_voices * sizeof(float));
    memcpy(b->stream_in, src, (size_t)frames * b->n_voices * sizeof(float));
    b->stream_frames = frames;
    return 0;
}

static int vev_cmp(const void *a, const void *b) {
    const vev *x = a, *y = b;
    if (x->frame != y->frame) return x->frame < y->frame ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq);
}
static int mev_cmp(const void *a, const void *b) {
    const mev *x = a, *y = b;
    if (x->frame != y->frame) return x->frame < y->frame ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq);
}

/* one instance, one block: the generated outer-graph walk (fm-synth lib.rs:72-130 /
   electric-piano main.rs:55-96), frames in order, voices in index order */
static void render_instance(orc_bank *b, uint32_t inst, uint32_t frames, const float *ramp_tab) {
    const uint32_t vpi = b->vpi, v0 = inst * vpi;
    uint32_t cur[ORC_MAX_VOICES] = {0};
    uint32_t mcur = 0;
    for (uint32_t v = 0; v < vpi; ++v)
        if (b->npend[v0 + v] > 1) qsort(b->pend[v0 + v], b->npend[v0 + v], sizeof(vev), vev_cmp);
    if (b->nmpend[inst] > 1) qsort(b->mpend[inst], b->nmpend[inst], sizeof(mev), mev_cmp);

    for (uint32_t f = 0; f < frames; ++f) {
        float gates[ORC_MAX_VOICES][MAXQ]; int ngates[ORC_MAX_VOICES] = {0};
        /* --- MIDI front: parser -> allocator -> handlers (per frame, events at this frame) */
        if (mcur < b->nmpend[inst] && b->mpend[inst][mcur].frame == f) {
            /* MidiParser::on_midi_in, midi.rs:193-224 */
            uint8_t on_note[MAXQ]; float on_vel[MAXQ]; int n_on = 0;
            uint8_t off_note[MAXQ]; int n_off = 0;
            while (mcur < b->nmpend[inst] && b->mpend[inst][mcur].frame == f) {
                const mev *m = &b->mpend[inst][mcur++];
                uint8_t note = 0; float vel = 0.0f;
                int r = orc_midi_parse(m->bytes, m->len, &note, &vel);
                if (r == 1 && n_on < MAXQ) { on_note[n_on] = note; on_vel[n_on] = vel; ++n_on; }
                else if (r == 2 && n_off < MAXQ) { off_note[n_off++] = note; }
            }
            /* VoiceAllocator: note_on queue first, then note_off (derive order, oscen-macros lib.rs:266-295);
               per-voice output queue holds (kind, note, vel) in push order */
            struct { uint8_t kind, note; float vel; } vq[ORC_MAX_VOICES][MAXQ]; int nvq[ORC_MAX_VOICES] = {0};
            orc_voice_alloc *al = &b->alloc[inst];
            for (int i = 0; i < n_on; ++i) {                       /* voice_allocator.rs:114-124 */
                int vi = orc_voice_alloc_allocate(al, on_note[i]);
                if (vi < (int)vpi && nvq[vi] < MAXQ) { vq[vi][nvq[vi]].kind = 1; vq[vi][nvq[vi]].note = on_note[i]; vq[vi][nvq[vi]].vel = on_vel[i]; nvq[vi]++; }
            }
            for (int i = 0; i < n_off; ++i) {                      /* voice_allocator.rs:126-136 */
                int vi = orc_voice_alloc_find(al, off_note[i]);
                if (vi >= 0) {
                    if (nvq[vi] < MAXQ) { vq[vi][nvq[vi]].kind = 2; vq[vi][nvq[vi]].note = off_note[i]; vq[vi][nvq[vi]].vel = 0.0f; nvq[vi]++; }
                    orc_voice_alloc_release(al, vi);
                }
            }
            /* MidiVoiceHandler: all note_on events, then all note_off events (midi.rs:92-121) */
            for (uint32_t v = 0; v < vpi; ++v) {
                midi_handler *h = &b->handlers[v0 + v];
                for (int i = 0; i < nvq[v]; ++i) if (vq[v][i].kind == 1) {
                    h->current_note = vq[v][i].note;
                    h->current_frequency = orc_midi_note_to_freq(vq[v][i].note);
                    if (ngates[v] < MAXQ) gates[v][ngates[v]++] = vq[v][i].vel;
                }
                for (int i = 0; i < nvq[v]; ++i) if (vq[v][i].kind == 2) {
                    if (h->current_note == (int)vq[v][i].note) {
                        if (ngates[v] < MAXQ) gates[v][ngates[v]++] = 0.0f;
                        h->current_note = -1;
                    }
                }
                /* process(): frequency = current_frequency (midi.rs:82-88); value edge into the voice */
                b->in[(size_t)(v0 + v) * MAX_IN + 0] = h->current_frequency;
            }
        }
        float mix = 0.0f;
        for (uint32_t v = 0; v < vpi; ++v) {
            const uint32_t gv = v0 + v;
            float *in = b->in + (size_t)gv * MAX_IN;
            /* ramped graph inputs read via .current (emit_node.rs:317-321) */
            for (int i = 0; i < N_INPUTS[b->kind]; ++i) if (b->ramped[i]) in[i] = ramp_tab[(size_t)i * 512 + f];
            /* direct per-voice events at this frame (gate -> event queue, value -> field) */
            while (cur[v] < b->npend[gv] && b->pend[gv][cur[v]].frame == f) {
                const vev *e = &b->pend[gv][cur[v]++];
                if (e->input_id == 0xFFFFFFFFu) { if (ngates[v] < MAXQ) gates[v][ngates[v]++] = e->scalar; }
                else if (e->input_id < (uint32_t)N_INPUTS[b->kind]) in[e->input_id] = e->scalar;
            }
            void *p = VOICE(b, gv);
            float out;
            switch (b->kind) {
            case K_FM: fm_voice_process(p, in, gates[v], ngates[v]); out = ((fm_voice *)p)->audio_out; break;
            case K_TWIN: twin_process(p, in, b->stream_in ? b->stream_in[(size_t)f * b->n_voices + gv] : 0.0f);
                         out = ((twin_voice *)p)->audio_out; break;
            case K_SAT1: case K_SAT4: sat_process(p, in); out = ((sat_voice *)p)->audio_out; break;
            case K_EPIANO: ep_voice_process(p, in, gates[v], ngates[v]); out = ((ep_voice *)p)->output; break;
            case K_SYNTH: case K_SYNTH_ADSR: synth_process(p, in, gates[v], ngates[v]); out = ((synth_voice *)p)->audio_out; break;
            default: clip_process(p, b->stream_in ? b->stream_in[(size_t)f * b->n_voices + gv] : 0.0f);
                     out = ((clip_voice *)p)->out; break;
            }
            b->voice_out[(size_t)f * b->n_voices + gv] = out;
            /* array -> scalar: .iter().map().sum(), sequential in index order (emit_node.rs:463-466) */
            mix = (v == 0) ? (0.0f + out) : (mix + out);
        }
        float *io = b->inst_out + ((size_t)f * b->n_inst + inst) * b->channels;
        if (b->kind == K_EPIANO) {
            ep_tremolo *t = &b->trem[inst];
            const float *in0 = b->in + (size_t)v0 * MAX_IN;
            t->input = mix; t->depth = in0[7]; t->rate = in0[8];
            ep_tremolo_process(t);
            io[0] = t->out[0]; io[1] = t->out[1];
        } else {
            io[0] = mix;
        }
    }
    /* events beyond the block are dropped with the staged queue (codegen/mod.rs:836-872) */
    for (uint32_t v = 0; v < vpi; ++v) b->npend[v0 + v] = 0;
    b->nmpend[inst] = 0;
}

typedef struct { orc_bank *b; uint32_t i0, i1, frames; const float *ramp_tab; } render_job;
static void *render_thread(void *arg) {
    render_job *j = arg;
    for (uint32_t i = j->i0; i < j->i1; ++i) render_instance(j->b, i, j->frames, j->ramp_tab);
    return NULL;
}
int orc_bank_process_block(orc_bank *b, uint32_t frames) {
    if (frames > 512) return -1;
    if ((b->kind == K_TWIN || kind_is_clip(b->kind)) && b->stream_in && b->stream_frames < frames) return -1;
    /* tick_ramps once per frame (codegen/mod.rs:878-914); table of .current per frame */
    float *ramp_tab = calloc((size_t)MAX_IN * 512, sizeof(float));
    for (uint32_t f = 0; f < frames; ++f)
        for (int i = 0; i < N_INPUTS[b->kind]; ++i) if (b->ramped[i]) {
            orc_ramp_tick(&b->ramps[i]);
            ramp_tab[(size_t)i * 512 + f] = b->ramps[i].current;
        }
    int nt = b->n_threads;
    if (nt > (int)b->n_inst) nt = (int)b->n_inst;
    if (nt <= 1) {
        for (uint32_t inst = 0; inst < b->n_inst; ++inst) render_instance(b, inst, frames, ramp_tab);
    } else {                       /* instances statically partitioned per thread, no sharing */
        pthread_t *th = malloc((size_t)nt * sizeof *th);
        render_job *jobs = malloc((size_t)nt * sizeof *jobs);
        for (int t = 0; t < nt; ++t) {
            jobs[t].b = b; jobs[t].frames = frames; jobs[t].ramp_tab = ramp_tab;
            jobs[t].i0 = (uint32_t)((uint64_t)b->n_inst * (uint64_t)t / (uint64_t)nt);
            jobs[t].i1 = (uint32_t)((uint64_t)b->n_inst * (uint64_t)(t + 1) / (uint64_t)nt);
            pthread_create(&th[t], NULL, render_thread, &jobs[t]);
        }
        for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
        free(th); free(jobs);
    }
    free(ramp_tab);
    b->last_frames = frames;
    return 0;
}
int orc_bank_read_voice_out(const orc_bank *b, float *dst, uint32_t first, uint32_t count) {
    if (first + count > b->n_voices) return -1;
    for (uint32_t f = 0; f < b->last_frames; ++f)
        memcpy(dst + (size_t)f * count, b->voice_out + (size_t)f * b->n_voices + first, count * sizeof(float));
    return 0;
}
int orc_bank_read_instance_out(const orc_bank *b, float *dst, uint32_t first, uint32_t count) {
    if (first + count > b->n_inst) return -1;
    const uint32_t ch = b->channels;
    for (uint32_t f = 0; f < b->last_frames; ++f)
        memcpy(dst + (size_t)f * count * ch, b->inst_out + ((size_t)f * b->n_inst + first) * ch,
               (size_t)count * ch * sizeof(float));
    return 0;
}
int orc_bank_read_mix(const orc_bank *b, float *dst) {
    const uint32_t ch = b->channels;
    for (uint32_t f = 0; f < b->last_frames; ++f)
        for (uint32_t c = 0; c < ch; ++c) {
            float s = 0.0f;
            for (uint32_t i = 0; i < b->n_inst; ++i) s += b->inst_out[((size_t)f * b->n_inst + i) * ch + c];
            dst[(size_t)f * ch + c] = s;
        }
    return 0;
}
