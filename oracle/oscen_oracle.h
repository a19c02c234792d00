/*
 * oscen_oracle.h — CPU ORACLE for the oscen signal-graph hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * (reedrosenbluth/oscen @ 8061f7d5) per-sample graph walk and node process()
 * bodies.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` leg may load it.  The product (oscen_b200/) never links,
 * imports or calls anything in oracle/.
 *
 * Parity pin: the Rust reference cannot be compiled in this environment (no
 * cargo/rustc, nightly + un-vendored crates), so this restatement is pinned on
 * every known-answer vector the reference's own tests hold for the path
 * (see tests/test_oracle_golden.py).  End-to-end
 * waveforms of the example graphs are NOT pinned by the reference itself
 * (it has no golden waveforms): for those "parity unpinned" applies and the
 * oracle of record is this file.
 *
 * Arithmetic: IEEE f32, compiled -O2 -ffp-contract=off (rustc never contracts
 * a*b+c); transcendental functions are the platform libm (glibc), which is what
 * Rust's f32::sin/exp/tan/tanh/powf lower to on Linux.
 *
 * All citations are path:line under /root/reference.
 */
#ifndef OSCEN_ORACLE_H
#define OSCEN_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- ramps -- */
/* oscen-lib/src/graph/types.rs:300-373 */
typedef struct {
    float current, target, increment;
    uint32_t frames_remaining;
} orc_ramp;
void orc_ramp_new(orc_ramp *r, float initial);
void orc_ramp_set_immediate(orc_ramp *r, float v);
void orc_ramp_set_with_ramp(orc_ramp *r, float target, uint32_t frames);
int  orc_ramp_tick(orc_ramp *r);            /* 1 when the ramp just completed */
int  orc_ramp_is_ramping(const orc_ramp *r);

/* ----------------------------------------------------------------- ADSR -- */
/* oscen-lib/src/envelope/adsr.rs:20-306 */
enum { ORC_IDLE = 0, ORC_ATTACK = 1, ORC_DECAY = 2, ORC_SUSTAIN = 3, ORC_RELEASE = 4 };
typedef struct {
    float attack, decay, sustain, release;      /* value inputs */
    float output;
    int   stage;
    uint32_t attack_samples, decay_samples, release_samples, samples_remaining;
    float attack_coeff, decay_coeff, release_increment;
    float level, target_level, sustain_level, velocity;
    float sample_rate;
} orc_adsr;
void orc_adsr_new(orc_adsr *e, float a, float d, float s, float r);
void orc_adsr_set_sample_rate(orc_adsr *e, float sr);
void orc_adsr_prepare(orc_adsr *e);
void orc_adsr_on_gate(orc_adsr *e, float scalar);   /* EventPayload::Scalar(v) */
void orc_adsr_process(orc_adsr *e);

/* ------------------------------------------------------------ TptFilter -- */
/* oscen-lib/src/filters/tpt/mod.rs:12-138 (mono F = f32) */
typedef struct {
    float input, cutoff, q, f_mod, output;
    float current_cutoff, current_q;
    float z[2];
    float h, g, r, k;
    float sample_rate;
} orc_tpt;
void orc_tpt_new(orc_tpt *f, float cutoff, float q);
void orc_tpt_set_sample_rate(orc_tpt *f, float sr);
void orc_tpt_prepare(orc_tpt *f);
void orc_tpt_process(orc_tpt *f);

/* ---------------------------------------------------- PolyBlepOscillator -- */
/* oscen-lib/src/oscillators/mod.rs:79-233 */
enum { ORC_WAVE_SINE = 0, ORC_WAVE_SAW = 1, ORC_WAVE_SQUARE = 2, ORC_WAVE_TRIANGLE = 3 };
typedef struct {
    float phase, phase_mod, frequency, frequency_mod, amplitude, pulse_width, output;
    int   waveform;
    float sample_rate;
} orc_polyblep;
void orc_polyblep_new(orc_polyblep *o, float freq, float amp, int waveform);
void orc_polyblep_set_sample_rate(orc_polyblep *o, float sr);
void orc_polyblep_process(orc_polyblep *o);
float orc_poly_blep(float t, float dt);
float orc_poly_blamp(float t, float dt);

/* naive Oscillator (oscillators/mod.rs:7-76); waveform 0 sine 1 square 2 saw */
typedef struct {
    float phase, frequency, frequency_mod, amplitude, output;
    int   waveform;
    float sample_rate;
} orc_osc;
void orc_osc_new(orc_osc *o, float freq, float amp, int waveform);
void orc_osc_process(orc_osc *o);

/* ------------------------------------------------- fm-synth example nodes -- */
/* examples/fm-synth/src/nodes/fm_operator.rs:12-76 */
typedef struct {
    float phase, prev_output, sample_rate;
    float base_freq, ratio, phase_mod, feedback, envelope, level, output;
} orc_fmop;
void orc_fmop_new(orc_fmop *o);
void orc_fmop_process(orc_fmop *o);

/* ------------------------------------------------------------ LP18Filter -- */
/* examples/nih-twin-peaks/src/lp18_filter.rs:7-107 */
typedef struct {
    float input, cutoff, fmod, resonance;
    float z[3], g, h;
    float last_cutoff, last_fmod, last_resonance;
    float sample_rate, output;
} orc_lp18;
void orc_lp18_new(orc_lp18 *f, float cutoff, float resonance);
void orc_lp18_prepare(orc_lp18 *f);
void orc_lp18_process(orc_lp18 *f);

/* ----------------------------------------------------------- IirLowpass -- */
/* oscen-lib/src/filters/iir_lowpass/mod.rs:15-166 */
typedef struct {
    float input, cutoff, q, output;
    float b0, b1, b2, a1, a2, v1, v2;
    float sample_rate;
    size_t frame_counter, frames_per_update;
} orc_iirlp;
void orc_iirlp_new(orc_iirlp *f, float cutoff, float q);
void orc_iirlp_prepare(orc_iirlp *f);
void orc_iirlp_process(orc_iirlp *f);

/* ---------------------------------------------------- Delay / RingBuffer -- */
/* oscen-lib/src/delay/mod.rs:5-86, ring_buffer/mod.rs:29-202 */
typedef struct { float *buffer; size_t write_pos, capacity, mask; } orc_ring;
typedef struct {
    float input, delay_samples, feedback, output;
    orc_ring ring;
    float sample_rate;
    size_t frames_per_update, frame_counter;
} orc_delay;
void  orc_ring_init(orc_ring *r, size_t size);      /* RingBuffer::new(size): PowerOfTwo capacity */
void  orc_ring_push(orc_ring *r, float v);
float orc_ring_get(const orc_ring *r, float offset);
void  orc_ring_free(orc_ring *r);
void orc_delay_new(orc_delay *d, float delay_samples, float feedback);
void orc_delay_prepare(orc_delay *d);
void orc_delay_process(orc_delay *d);
void orc_delay_free(orc_delay *d);

/* -------------------------------------------------------- SamplePlayer -- */
/* oscen-lib/src/sample_player/mod.rs:64-127 (mono, F = f32).  `pending` stands for the handoff slot: a buffer published
   by orc_sample_player_publish is picked up by the next process(), which resets the playhead */
typedef struct {
    float output;
    const float *current; size_t current_len;
    const float *pending; size_t pending_len; int has_pending;
    size_t playhead;
} orc_sample_player;
void orc_sample_player_new(orc_sample_player *p);
/* SamplePlayerConsumer::build for F = f32 (mod.rs:36-49): channel 0 of an interleaved source; `dst` holds `frames` floats */
void orc_sample_player_build_mono(float *dst, const float *interleaved, size_t frames, size_t channels);
void orc_sample_player_publish(orc_sample_player *p, const float *samples, size_t frames);   /* the caller keeps `samples` alive */
void orc_sample_player_process(orc_sample_player *p);

/* -------------------------------------------------- halfband FIR cascade -- */
/* oscen-lib/src/resample/sinc_fir.rs:95-266, coeffs.rs:17-34 */
typedef struct { float history[24]; size_t head; } orc_hb_down;   /* Halfband2xDownStage */
typedef struct { float history[12]; size_t head; } orc_hb_up;     /* Halfband2xUpStage */
typedef struct { orc_hb_down stages[3]; int n_stages; int n; } orc_sinc_down;
typedef struct { orc_hb_up stages[3]; int n_stages; int n; } orc_sinc_up;
void  orc_sinc_down_new(orc_sinc_down *d, int n);      /* n in {1,2,4,8} */
float orc_sinc_down_process(orc_sinc_down *d, const float *xs);
int   orc_sinc_down_latency(const orc_sinc_down *d);
void  orc_sinc_down_reset(orc_sinc_down *d);
void  orc_sinc_up_new(orc_sinc_up *u, int n);
void  orc_sinc_up_process(orc_sinc_up *u, float x, float *out);
int   orc_sinc_up_latency(const orc_sinc_up *u);

/* ------------------------------- IIR halfband / linear / latch resamplers -- */
/* oscen-lib/src/resample/halfband_iir.rs:24-262, coeffs.rs:36-54 */
typedef struct { float a, x_prev, y_prev; } orc_allpass1;                       /* Allpass1 */
typedef struct { orc_allpass1 branch_a[2], branch_b[2]; float prev_odd_in; } orc_iir_hb;   /* IirHalfband2x */
typedef struct { orc_iir_hb stages[3]; int n_stages; int n; } orc_iir_up, orc_iir_down;
void  orc_iir_up_new(orc_iir_up *u, int n);            /* n in {1,2,4,8} */
void  orc_iir_up_process(orc_iir_up *u, float x, float *out);
int   orc_iir_up_latency(const orc_iir_up *u);
void  orc_iir_down_new(orc_iir_down *d, int n);
float orc_iir_down_process(orc_iir_down *d, const float *xs);
int   orc_iir_down_latency(const orc_iir_down *d);
/* oscen-lib/src/resample/linear.rs:11-74 */
typedef struct { float prev; int n; } orc_linear_up;
void  orc_linear_up_new(orc_linear_up *u, int n);
void  orc_linear_up_process(orc_linear_up *u, float x, float *out);
float orc_linear_down_process(int n, const float *xs);
int   orc_linear_up_latency(int n);
int   orc_linear_down_latency(int n);
/* oscen-lib/src/resample/latch.rs:6-54 */
void  orc_latch_up_process(int n, float x, float *out);
float orc_latch_down_process(int n, const float *xs);

/* ------------------------------------------------------- MIDI + voices --- */
/* oscen-lib/src/midi.rs:69-72, 147-224; voice_allocator.rs:29-150 */
float orc_midi_note_to_freq(uint8_t note);
/* returns 0 none, 1 note-on (note, velocity filled), 2 note-off (note filled) */
int   orc_midi_parse(const uint8_t *data, size_t len, uint8_t *note, float *velocity);

#define ORC_MAX_VOICES 24
typedef struct {
    int      num_voices;
    uint8_t  active[ORC_MAX_VOICES], released[ORC_MAX_VOICES];
    int      note[ORC_MAX_VOICES];          /* -1 == None */
    uint32_t age[ORC_MAX_VOICES];
    uint32_t current_age;
} orc_voice_alloc;
void orc_voice_alloc_new(orc_voice_alloc *a, int num_voices);
int  orc_voice_alloc_allocate(orc_voice_alloc *a, uint8_t note);
int  orc_voice_alloc_find(const orc_voice_alloc *a, uint8_t note);  /* -1 none */
void orc_voice_alloc_release(orc_voice_alloc *a, int idx);

/* ================================================================ banks == */
/*
 * A "bank" renders n_voices independent voices of one graph kind, grouped in
 * instances of voices_per_instance voices, exactly as n_voices/vpi separate
 * generated graph structs would: per frame, per voice, the topologically
 * ordered walk of SURVEY.md §3.2 / Appendix A.  It mirrors the C-ABI of
 * include/oscen_b200.h one-to-one so parity tests drive both with the same
 * script.  Ids (graph kinds, input ids) are the same as in oscen_b200.h.
 */
typedef struct orc_bank orc_bank;
typedef struct { uint32_t voice, frame_offset, input_id; float scalar; } orc_event;

orc_bank *orc_bank_create(uint32_t graph_kind, uint32_t n_voices, uint32_t voices_per_instance,
                          float sample_rate);
void   orc_bank_destroy(orc_bank *b);
int    orc_bank_set_value(orc_bank *b, uint32_t input_id, uint32_t first_voice, uint32_t count,
                          const float *v, int broadcast);
int    orc_bank_set_value_ramp(orc_bank *b, uint32_t input_id, float target, uint32_t frames);
int    orc_bank_push_events(orc_bank *b, const orc_event *ev, size_t n);
int    orc_bank_push_midi(orc_bank *b, uint32_t instance, uint32_t frame_offset,
                          const uint8_t *bytes, uint32_t len);
int    orc_bank_stream_in(orc_bank *b, uint32_t input_id, const float *src /* [frame][voice] */,
                          uint32_t frames);
int    orc_bank_process_block(orc_bank *b, uint32_t frames);
/* dst[frame * count + (v - first_voice)] */
int    orc_bank_read_voice_out(const orc_bank *b, float *dst, uint32_t first_voice, uint32_t count);
/* dst[(frame * count + (i - first_instance)) * channels + c] */
int    orc_bank_read_instance_out(const orc_bank *b, float *dst, uint32_t first_instance,
                                  uint32_t count);
/* dst[frame * channels + c]: sequential sum over instances in index order */
int    orc_bank_read_mix(const orc_bank *b, float *dst);
uint32_t orc_bank_channels(const orc_bank *b);
uint32_t orc_bank_latency(const orc_bank *b);
void   orc_bank_set_threads(orc_bank *b, int n_threads);
/* graph kinds 7..19: the cross-rate clipper graphs of oscen-lib/tests/multirate_graph.rs:282-325
   (ClipRef; ClipOversampled / ClipOversampledIir and their [linear] / [latch] policy variants) */
#define ORC_GRAPH_CLIP_REF 7u
#define ORC_GRAPH_CLIP_OS(policy, log2_factor) (8u + (policy) * 3u + ((log2_factor) - 1u))
enum { ORC_POLICY_SINC = 0, ORC_POLICY_SINC_IIR = 1, ORC_POLICY_LINEAR = 2, ORC_POLICY_LATCH = 3 };
#define ORC_EV_GATE 0xFFFFFFFFu   /* orc_event.input_id of a gate event; else a value input id */

#ifdef __cplusplus
}
#endif
#endif
