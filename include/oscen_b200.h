/*
 * oscen_b200.h — C ABI of the B200-native, voice-batched engine for Oscen's signal-graph
 * evaluation loop (the generated `process()` / `process_block()` walk).
 *
 * The reference (reedrosenbluth/oscen @ 8061f7d5) has NO FFI boundary on this path: nodes are
 * monomorphised and inlined into the struct `graph!` generates.  The boundary honoured here is
 * therefore the API surface that struct exposes to its host.  Each entry point
 * cites the generated/reference interface it replaces (paths relative to the reference repo):
 *
 *   oscen_b200_create          G::new() + G::init(sr)              oscen-graph-compiler/src/codegen/mod.rs:1309-1350
 *   oscen_b200_set_value       `g.<value> = v` / set_<v>_immediate  codegen/mod.rs:959-975
 *   oscen_b200_set_value_ramp  set_<v>_with_ramp(v, frames)         codegen/mod.rs:948-957
 *   oscen_b200_push_events     <event>.try_push(EventInstance{..})  oscen-lib/src/graph/types.rs:183-188
 *   oscen_b200_push_midi       midi_in.try_push(raw_midi_event(..)) examples/fm-synth/src/main.rs:157-165
 *   oscen_b200_stream_in       writing `<in>_block[..]`             codegen/emit_frame.rs:33-41
 *   oscen_b200_process_block   G::process_block(frames)             codegen/mod.rs:755-872
 *   oscen_b200_read_instance_out  reading `<out>_block[..]`         codegen/emit_frame.rs:44-52
 *   oscen_b200_read_voice_out  (parity hook: `voices[i].audio_out`) codegen/emit_node.rs:463-466
 *   oscen_b200_read_mix        Σ over instances of `<out>_block` (the bank's mix bus; new)
 *   oscen_b200_latency_samples G::latency_samples()                 codegen/mod.rs:644-716
 *
 * Conventions: every call returns 0 on success or a negative OSCEN_B200_E_* code and never
 * unwinds across the boundary; the caller owns host buffers, the engine owns device buffers;
 * no callbacks; one caller thread per engine at a time (mirrors `&mut self`: the engine is
 * Send, not Sync).  `process*` in the reference are infallible; here a failure can only be a
 * CUDA/NCCL error or an argument error, reported through the return code and last_error.
 *
 * Data layout: per-voice state is SoA `[field][voice]` in HBM; streams are `[frame][voice]`
 * (voice fastest) so a warp touches 128 contiguous bytes.
 */
#ifndef OSCEN_B200_H
#define OSCEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSCEN_B200_ABI_VERSION 1
#define OSCEN_B200_MAX_BLOCK 512u          /* DEFAULT_MAX_BLOCK_SIZE, oscen-lib/src/graph/types.rs:12 */
#define OSCEN_B200_MAX_EVENTS_PER_ENDPOINT 32u /* MAX_STATIC_EVENTS_PER_ENDPOINT, graph/types.rs:18 */

/* graph kinds: the reference graphs on the hot path */
enum {
    OSCEN_B200_GRAPH_FM_VOICE = 0,    /* FMVoice x N, instance = FMGraph voice sum   examples/fm-synth/src/fm_voice.rs:6-156 */
    OSCEN_B200_GRAPH_TWIN_PEAKS = 1,  /* TwinPeaksGraph (2 x LP18Filter, a - b)       examples/nih-twin-peaks/src/lib.rs:15-48 */
    OSCEN_B200_GRAPH_SAT_1X = 2,      /* SatGraph_1x                                  examples/oversampled-saturator/src/main.rs:64-80 */
    OSCEN_B200_GRAPH_SAT_4X = 3,      /* SatGraph_4x (4x inner loop + SincDownFir<4>) */
    OSCEN_B200_GRAPH_EPIANO = 4,      /* ElectricPianoVoiceNode x N, instance = sum -> Tremolo -> Frame<2>  examples/electric-piano/src/main.rs:33-97 */
    OSCEN_B200_GRAPH_SYNTH = 5,       /* README `Synth`: sine LFO -> saw PolyBLEP -> TptFilter   README.md:24-53 */
    OSCEN_B200_GRAPH_SYNTH_ADSR = 6,  /* README `Synth` + AdsrEnvelope -> Gain(vca) (wiring of benches/static_vs_runtime.rs:57-59) */
    /* cross-rate graphs: per-voice stream input -> Up edge -> HardClip `* N` -> Down edge
       (oscen-lib/tests/multirate_graph.rs:282-325).  HardClip here is
       clamp(+-0.7) with no drive (multirate_graph.rs:35-40). */
    OSCEN_B200_GRAPH_CLIP_REF = 7,    /* ClipRef: input -> clip -> out at 1x           multirate_graph.rs:282-293 */
    OSCEN_B200_GRAPH_CLIP_OS_BASE = 8,/* 8..19 = OSCEN_B200_GRAPH_CLIP_OS(policy, log2_factor), below */
    OSCEN_B200_GRAPH_CUSTOM = 20,     /* a graph compiled from `graph!` DSL text: oscen_b200_graph_compile + oscen_b200_create_from_graph */
    OSCEN_B200_GRAPH_COUNT = 21
};
/* connection policy of both cross-rate edges (`[sinc]`, `[sinc_iir]`, `[linear]`, `[latch]`,
   oscen-graph-compiler/src/ast.rs ConnectionPolicy; resampler kernels oscen-lib/src/resample/) */
enum {
    OSCEN_B200_POLICY_SINC = 0,       /* SincUpFir<N> / SincDownFir<N>          resample/sinc_fir.rs:146-266 */
    OSCEN_B200_POLICY_SINC_IIR = 1,   /* IirHalfbandUp<N> / IirHalfbandDown<N>  resample/halfband_iir.rs:146-262 */
    OSCEN_B200_POLICY_LINEAR = 2,     /* LinearUp<N> / LinearDown<N>            resample/linear.rs:11-74 */
    OSCEN_B200_POLICY_LATCH = 3       /* LatchUp<N> / LatchDown<N>              resample/latch.rs:6-54 */
};
/* ClipOversampled (`clip = HardClip::new() * N`, N = 2^log2_factor in {2,4,8}) with `policy` on both edges;
   OSCEN_B200_GRAPH_CLIP_OS(OSCEN_B200_POLICY_SINC, 2) is the reference's ClipOversampled,
   (OSCEN_B200_POLICY_SINC_IIR, 2) its ClipOversampledIir. */
#define OSCEN_B200_GRAPH_CLIP_OS(policy, log2_factor) (8u + (uint32_t)(policy) * 3u + ((uint32_t)(log2_factor) - 1u))

/* value-input ids (declaration order of the reference graph's `input x: value` lines) */
enum { /* FM_VOICE, fm_voice.rs:10-48 */
    OSCEN_B200_FM_FREQUENCY = 0, OSCEN_B200_FM_OP3_RATIO, OSCEN_B200_FM_OP3_LEVEL, OSCEN_B200_FM_OP3_FEEDBACK,
    OSCEN_B200_FM_OP3_ATTACK, OSCEN_B200_FM_OP3_DECAY, OSCEN_B200_FM_OP3_SUSTAIN, OSCEN_B200_FM_OP3_RELEASE,
    OSCEN_B200_FM_OP2_RATIO, OSCEN_B200_FM_OP2_LEVEL, OSCEN_B200_FM_OP2_FEEDBACK,
    OSCEN_B200_FM_OP2_ATTACK, OSCEN_B200_FM_OP2_DECAY, OSCEN_B200_FM_OP2_SUSTAIN, OSCEN_B200_FM_OP2_RELEASE,
    OSCEN_B200_FM_OP1_RATIO, OSCEN_B200_FM_OP1_ATTACK, OSCEN_B200_FM_OP1_DECAY, OSCEN_B200_FM_OP1_SUSTAIN,
    OSCEN_B200_FM_OP1_RELEASE, OSCEN_B200_FM_ROUTE, OSCEN_B200_FM_FILTER_CUTOFF, OSCEN_B200_FM_FILTER_RESONANCE,
    OSCEN_B200_FM_FILTER_ATTACK, OSCEN_B200_FM_FILTER_DECAY, OSCEN_B200_FM_FILTER_SUSTAIN,
    OSCEN_B200_FM_FILTER_RELEASE, OSCEN_B200_FM_FILTER_ENV_AMOUNT, OSCEN_B200_FM_N_INPUTS
};
enum { OSCEN_B200_TWIN_CUTOFF_A = 0, OSCEN_B200_TWIN_CUTOFF_B, OSCEN_B200_TWIN_RESONANCE, OSCEN_B200_TWIN_N_INPUTS };
enum { OSCEN_B200_SAT_FREQUENCY = 0, OSCEN_B200_SAT_AMPLITUDE, OSCEN_B200_SAT_N_INPUTS };   /* osc.frequency / osc.amplitude */
enum { /* EPIANO: [0] per voice; [1..6] voice params; [7..8] instance (tremolo), read from the instance's first voice */
    OSCEN_B200_EP_FREQUENCY = 0, OSCEN_B200_EP_BRIGHTNESS, OSCEN_B200_EP_VELOCITY_SCALING, OSCEN_B200_EP_DECAY_RATE,
    OSCEN_B200_EP_HARMONIC_DECAY, OSCEN_B200_EP_KEY_SCALING, OSCEN_B200_EP_RELEASE_RATE,
    OSCEN_B200_EP_VIBRATO_INTENSITY, OSCEN_B200_EP_VIBRATO_SPEED, OSCEN_B200_EP_N_INPUTS
};
enum { OSCEN_B200_SYNTH_MOD_FREQ = 0, OSCEN_B200_SYNTH_MOD_DEPTH, OSCEN_B200_SYNTH_CARRIER_FREQ, OSCEN_B200_SYNTH_CUTOFF,
       OSCEN_B200_SYNTH_ATTACK, OSCEN_B200_SYNTH_DECAY, OSCEN_B200_SYNTH_SUSTAIN, OSCEN_B200_SYNTH_RELEASE };
#define OSCEN_B200_MAX_INPUTS 28u

/* input_id of a gate event in oscen_b200_event; any other id must be a value-input id and
   means "set that value input of that voice at exactly that frame" (how MidiVoiceHandler's
   `frequency` value output reaches a voice, oscen-lib/src/midi.rs:82-88,92-104) */
#define OSCEN_B200_EVENT_GATE 0xFFFFFFFFu

/* EventInstance{frame_offset, payload: Scalar(v)} addressed to one voice
   (oscen-lib/src/graph/types.rs:133-137); gate: scalar > 0 = velocity / note on, <= 0 = off */
typedef struct {
    uint32_t voice;
    uint32_t frame_offset;
    uint32_t input_id;
    float    scalar;
} oscen_b200_event;

/* error codes */
enum {
    OSCEN_B200_OK = 0,
    OSCEN_B200_E_ARG = -1,          /* bad argument (id/range/size)                      */
    OSCEN_B200_E_STATE = -2,        /* call not valid in this state (e.g. ramp on a non-uniform input) */
    OSCEN_B200_E_CUDA = -3,         /* CUDA runtime error; see oscen_b200_last_error      */
    OSCEN_B200_E_NCCL = -4,         /* NCCL error / NCCL not loadable                     */
    OSCEN_B200_E_NOMEM = -5,
    OSCEN_B200_E_NO_DEVICE = -6,    /* no CUDA device: the engine has NO CPU fallback     */
    OSCEN_B200_E_JIT = -7           /* run-time compilation of a generated kernel failed (NVRTC missing / compile error) */
};

/* execution plans */
enum {
    OSCEN_B200_PLAN_FUSED = 0,      /* one kernel per graph: whole voice in registers       */
    OSCEN_B200_PLAN_NODES = 1       /* one kernel per node in topological order, streams in HBM, replayed as a CUDA graph */
};
/* output retention flags */
enum {
    OSCEN_B200_KEEP_VOICE_OUT = 1,     /* keep per-voice `audio_out` [frame][voice] (parity hook) */
    OSCEN_B200_KEEP_INSTANCE_OUT = 2   /* keep per-instance `<out>_block` [frame][instance][ch]   */
};

typedef struct oscen_b200_engine oscen_b200_engine;

typedef struct {
    uint32_t graph_kind;            /* OSCEN_B200_GRAPH_*                                    */
    uint32_t n_voices;              /* voices rendered by THIS engine (its shard)            */
    uint32_t voices_per_instance;   /* N of `voices = [Voice::new(); N]`; power of two <= 64 */
    float    sample_rate;           /* G::init(sample_rate)                                  */
    int32_t  device;                /* CUDA device ordinal                                   */
    uint32_t plan;                  /* OSCEN_B200_PLAN_*                                     */
    uint32_t keep;                  /* OSCEN_B200_KEEP_* flags                               */
    uint32_t reserved;
} oscen_b200_config;

int  oscen_b200_abi_version(void);
int  oscen_b200_device_count(void);

int  oscen_b200_create(const oscen_b200_config *cfg, oscen_b200_engine **out);
void oscen_b200_destroy(oscen_b200_engine *e);
const char *oscen_b200_last_error(const oscen_b200_engine *e);   /* e may be NULL: last create error */

/* value inputs.  broadcast != 0: v[0] to `count` voices; else v[i] to voice first_voice+i.
   Setting every voice to one value keeps the input "uniform" (a single ValueRampState can
   then ramp it, as for a graph-level `[ramp: N]` input). */
int  oscen_b200_set_value(oscen_b200_engine *e, uint32_t input_id, uint32_t first_voice, uint32_t count,
                          const float *v, int broadcast);
int  oscen_b200_set_value_ramp(oscen_b200_engine *e, uint32_t input_id, float target, uint32_t frames);

/* events for the NEXT process_block; frame_offset >= frames of that block are dropped, more
   than 32 events per voice per block are dropped (ArrayVec<_,32>::try_push).  A value event on an
   input that set_value_ramp has been used on is refused (OSCEN_B200_E_STATE): a ramped input is one
   ValueRampState for the whole bank, as for per-voice oscen_b200_set_value on it.  So is a value event on
   OSCEN_B200_GRAPH_EPIANO's instance-level inputs 7 / 8 (the graph's Tremolo): set_value / set_value_ramp. */
int  oscen_b200_push_events(oscen_b200_engine *e, const oscen_b200_event *ev, size_t n);
/* raw MIDI into instance `instance`'s MidiParser -> VoiceAllocator<N> -> MidiVoiceHandler chain
   (host-side integer logic; emits per-voice frequency + gate events) */
int  oscen_b200_push_midi(oscen_b200_engine *e, uint32_t instance, uint32_t frame_offset,
                          const uint8_t *bytes, uint32_t len);

/* Device-side MIDI front: the same semantics as n calls of oscen_b200_push_midi in
   array order, but MidiParser -> VoiceAllocator<N> -> MidiVoiceHandler run on the GPU, one thread
   per instance (oscen-lib/src/midi.rs:82-224, voice_allocator.rs:57-136), and write the per-voice
   gate / frequency events directly into the device event list — no host loop over instances.
   An engine uses EITHER push_midi (host front) OR push_midi_batch (device front) for its lifetime
   (the allocator state lives on one side); mixing returns OSCEN_B200_E_STATE, as does combining
   push_midi_batch with oscen_b200_push_events in one block.  voices_per_instance <= 16. */
typedef struct {
    uint32_t instance;
    uint32_t frame_offset;
    uint8_t  bytes[3];
    uint8_t  len;
} oscen_b200_midi_msg;
int  oscen_b200_push_midi_batch(oscen_b200_engine *e, const oscen_b200_midi_msg *msgs, size_t n);

/* stream input block [frames][n_voices] f32 for the NEXT process_block (TWIN_PEAKS audio_in,
   CLIP_* input).
   is_device != 0: `src` is a device pointer on the engine's device (no copy is made; it must
   stay valid until process_block returns). */
int  oscen_b200_stream_in(oscen_b200_engine *e, uint32_t input_id, const float *src, uint32_t frames,
                          int is_device);
/* synthetic stream input generated on device: audio_in[f][v] = U[-1,1) from a 64-bit LCG
   seeded (seed + global_voice_index), continuing across blocks */
int  oscen_b200_stream_in_lcg(oscen_b200_engine *e, uint32_t input_id, uint64_t seed, uint32_t voice_offset);

int  oscen_b200_process_block(oscen_b200_engine *e, uint32_t frames /* <= 512 */);
int  oscen_b200_synchronize(oscen_b200_engine *e);

/* outputs of the LAST block.  dst layouts: voice_out [frame][count]; instance_out
   [frame][count][channels]; mix [frame][channels] (sum over this engine's instances, and over
   all ranks once oscen_b200_comm_init was called). */
int  oscen_b200_read_voice_out(oscen_b200_engine *e, float *dst, uint32_t first_voice, uint32_t count);
int  oscen_b200_read_instance_out(oscen_b200_engine *e, float *dst, uint32_t first_instance, uint32_t count);
int  oscen_b200_read_mix(oscen_b200_engine *e, float *dst);

uint32_t oscen_b200_channels(const oscen_b200_engine *e);
uint32_t oscen_b200_latency_samples(const oscen_b200_engine *e);
uint32_t oscen_b200_n_inputs(uint32_t graph_kind);
size_t   oscen_b200_state_bytes_per_voice(const oscen_b200_engine *e);
/* raw SoA state download: dst[field * n_voices + voice], tests only */
int  oscen_b200_state_download(oscen_b200_engine *e, void *dst, size_t cap);

/* node plan of a graph kind as text, one launch per line ("k: NodeType [node, node]"): the
   topological order the plan compiler derives (Kahn over the stream edges, like
   oscen-graph-compiler/src/ir/lower.rs:1015-1085).  Returns the number of launches (>0), or a
   negative code for graphs that only have a fused kernel.  Needs no GPU. */
int  oscen_b200_plan_describe(uint32_t graph_kind, char *dst, size_t cap);

/* ---- the `graph!` DSL front end --------------------------------------------------
 * Replaces, at run time, what the reference does at Rust compile time: oscen_graph_compiler::compile
 * (oscen-graph-compiler/src/lib.rs:22-34) = parse_graph_def (parse.rs:195-241) -> ir::lower (ir/lower.rs)
 * -> dead-node pass (ir/passes/dead_nodes.rs) -> Kahn sort (ir/lower.rs:1015-1085) -> codegen.  The text
 * is the body of a `graph! { ... }` invocation (with or without the `graph! { }` wrapper): `name:`,
 * `input` / `output` declarations in both spellings, `node` / `nodes { }`, `connection` /
 * `connections { }`, compound sources (+ - * /, parentheses, .tanh() .abs() .clamp(a,b) .min(a) .max(a)).
 * The result is a node plan (one kernel launch per node type per ready wave, streams in HBM, replayed
 * as a CUDA graph) or one generated fused kernel.  Node types: AdsrEnvelope, FmOperator, TptFilter,
 * PolyBlepOscillator (sine/saw/square/triangle), LP18Filter, Gain, AddValue, Mixer, Crossfade, HardClip in both
 * back ends; Oscillator (sine/square/saw), IirLowpass, Delay, `* N` oversampled nodes with [sinc] / [sinc_iir] /
 * [linear] / [latch] cross-rate edges, and `-> [N] ->` / `-> [delay] ->` feedback edges in generated kernels
 * (OSCEN_B200_PLAN_FUSED) only.  What cannot be expressed is a compile ERROR with a line:column, never silently
 * dropped: `/ N` rates, node arrays outside the polyphonic shape (below), nested scalar graphs, `external`
 * assets, typed (array / Frame<N>) endpoints, free-function calls.
 * Value inputs get ids in declaration order (oscen_b200_graph_input_id); the graph's event input carries
 * gate events (OSCEN_B200_EVENT_GATE) to every AdsrEnvelope it is connected to; the graph's stream input
 * is oscen_b200_stream_in id 0.  Compiling needs no GPU. */
typedef struct oscen_b200_graph oscen_b200_graph;
/* err (optional) receives "line:col: message" on failure */
int  oscen_b200_graph_compile(const char *dsl_text, oscen_b200_graph **out, char *err, size_t err_cap);
/* the same, compiling the graph called `graph_name` instead of the last one of the text (NULL / "" = the last).
   `oversample_variants! { base_name: X; factors: [1, 4]; body: { ... {FACTOR} ... } }` expands to the graphs
   X_1x, X_4x, ... (oscen-macros/src/oversample_variants_macro.rs:94-139); graphs used as nodes of the chosen
   graph (`voice = SimpleVoice;`) are inlined. */
int  oscen_b200_graph_compile_named(const char *dsl_text, const char *graph_name, oscen_b200_graph **out, char *err, size_t err_cap);
void oscen_b200_graph_free(oscen_b200_graph *g);
/* text form: inputs with defaults / ranges / ramps, stream + event input names, output, state words per
   voice, then the launch plan "k: NodeType [node, node]".  Returns the number of launches. */
int  oscen_b200_graph_describe(const oscen_b200_graph *g, char *dst, size_t cap);
uint32_t oscen_b200_graph_n_inputs(const oscen_b200_graph *g);
/* Polyphonic graphs.  The text may hold several `graph! { }` blocks; the LAST one is compiled, earlier ones are
   the graphs it instantiates.  An outer graph of the reference's polyphonic shape (examples/fm-synth/src/lib.rs:
   22-131: `voices = [Voice::new(); N]` + MidiParser -> VoiceAllocator::<N> -> [MidiVoiceHandler; N], broadcast
   edges `value -> voices.input`, `voices.output -> out`) compiles to the VOICE graph's plan with
   voices_per_instance = N: value inputs keep the voice graph's ids but answer to the outer graph's names (with the
   outer declaration's default / ramp), the MIDI chain enables oscen_b200_push_midi / _push_midi_batch, and the
   voice sum is the engine's in-order instance fold.  cfg->voices_per_instance must be N or 0 for such a graph. */
uint32_t oscen_b200_graph_voices_per_instance(const oscen_b200_graph *g);   /* N of the voice array, 1 if not polyphonic */
uint32_t oscen_b200_graph_channels(const oscen_b200_graph *g);              /* lanes of the stream output: 1, or 2 for `output out: stream: Frame<2>;` (frame.rs:57-82) */
int  oscen_b200_graph_has_midi(const oscen_b200_graph *g);
int  oscen_b200_graph_input_id(const oscen_b200_graph *g, const char *name);          /* >= 0, or OSCEN_B200_E_ARG */
typedef struct {
    char     name[64];
    float    default_value;
    uint32_t has_range; float range_min, range_max; uint32_t logarithmic;   /* `[min..max, log]` */
    uint32_t ramp_frames;            /* `ramp: N`: frames set_<v>() ramps over (codegen/mod.rs:917-976); 0 = plain field */
} oscen_b200_input_info;
int  oscen_b200_graph_input_info(const oscen_b200_graph *g, uint32_t input_id, oscen_b200_input_info *out);
/* G::new() + G::init(sr) for n_voices instances of the compiled graph.  cfg->graph_kind must be
   OSCEN_B200_GRAPH_CUSTOM.  cfg->plan selects how it runs:
     OSCEN_B200_PLAN_NODES  the node plan (one kernel per node type per ready wave, CUDA graph);
     OSCEN_B200_PLAN_FUSED  ONE kernel generated for this graph (every node inlined in topological order,
                            streams in registers — what `graph!` emits as Rust, codegen/mod.rs:499-572, emitted
                            here as CUDA) and compiled at create time with NVRTC for sm_100a (needs
                            libnvrtc.so.12 and the engine's device headers next to the library).
   Both are bit-identical.  The graph object may be freed afterwards. */
int  oscen_b200_create_from_graph(const oscen_b200_graph *g, const oscen_b200_config *cfg, oscen_b200_engine **out);
/* PLAN_FUSED engines of a compiled graph: generate, compile (NVRTC) and load a SECOND variant of the graph's kernel for the
   value inputs that hold one value for every voice at the time of the call (defaults, broadcast oscen_b200_set_value,
   declared ramps; the MIDI front's `frequency` input is never taken) — the reference's `pub <value>: f32` fields of ONE graph
   instance are exactly that (codegen/mod.rs:1292-1392); per-voice values exist here only because a bank holds many
   instances.  The variant reads those inputs from the kernel-argument constant bank instead of one register each (+5 % on
   the generated FMVoice, profiles/r02_ab_generated_knobs.log).  It renders every block in which each of these inputs is
   still uniform and not ramping; any other block (a per-voice set_value, a value event or a running ramp on one of them)
   runs the general kernel — results are bit-identical either way, oscen_b200_last_kernel_name() says which ran.
   Call it after the parameters are set up (it takes an NVRTC compile, so not from the audio callback); calling it again
   re-specialises for the inputs uniform then.  *uniform_mask (may be NULL) receives the input ids taken, bit i = input i.
   OSCEN_B200_E_STATE for built-in kinds (FMVoice picks its variant itself, fm_select.hpp) and PLAN_NODES engines. */
int  oscen_b200_specialize(oscen_b200_engine *e, uint32_t *uniform_mask);
/* (inspection, host only) the tables of a BUILT-IN graph kind: value-input defaults, the state words of one voice after
   new() + init(sample_rate), output channels, dynamic shared memory of its fused kernel.  For tools and for the CPU suite's
   host-executed kernel check (tests/host_kernel.py); OSCEN_B200_GRAPH_CUSTOM graphs answer oscen_b200_graph_kernel_info. */
typedef struct {
    uint32_t n_inputs, n_state, channels, dyn_smem_bytes, has_stream_in, reserved;
} oscen_b200_kind_tables;
int  oscen_b200_kind_info(uint32_t graph_kind, float sample_rate, oscen_b200_kind_tables *info,
                          float *defaults /* [OSCEN_B200_MAX_INPUTS] or NULL */, uint32_t *state_words, size_t state_cap);

/* `external NAME: AudioAsset;` slots (oscen-graph-compiler/src/parse.rs:447-456; the generated graph's
   `graph.<name>.load_wav(..)` handle, oscen-lib/src/asset/mod.rs): the slots of a graph in declaration order, and the
   run-time load — `interleaved` is frames x channels f32 at the engine's sample rate (decoding / rate conversion stay with
   the caller); a mono SamplePlayer takes channel 0 (sample_player/mod.rs:36-49).  One buffer per slot per engine, shared
   by its voices; every player bound to the slot restarts at the next block (mod.rs:107-112); frames == 0 empties the slot. */
#define OSCEN_B200_MAX_ASSETS 4
uint32_t oscen_b200_graph_n_assets(const oscen_b200_graph *g);
int  oscen_b200_graph_asset_id(const oscen_b200_graph *g, const char *name);
int  oscen_b200_load_asset(oscen_b200_engine *e, uint32_t asset_id, const float *interleaved, uint32_t frames, uint32_t channels);
/* the generated CUDA source of PLAN_FUSED (returns its length), and a compile-only check of it with
   NVRTC (returns the sm_100a cubin size; the compiler log / error goes to `log`).  Neither needs a GPU. */
int  oscen_b200_graph_codegen(const oscen_b200_graph *g, char *dst, size_t cap);
int  oscen_b200_graph_jit_check(const oscen_b200_graph *g, char *log, size_t log_cap);
/* what an engine of this graph at `sample_rate` would hand to NVRTC and start from (Delay rings are sized by the sample
   rate, so the state layout depends on it): the generated source, the state words of one voice after new() + init(sr)
   (field order = the SoA rows of oscen_b200_state_download), the dynamic shared memory of the kernel.  Host only, no GPU:
   offline inspection, and the CPU suite's host-executed kernel check (tests/host_kernel.py).  Any output may be NULL.
   Returns 0, or OSCEN_B200_E_ARG when a buffer is too small (the needed sizes are still written). */
typedef struct {
    uint32_t source_len;        /* bytes of the generated source, without the NUL */
    uint32_t n_state;           /* state words per voice (stream-input LCG words included) */
    uint32_t dyn_smem_bytes;    /* dynamic shared memory per CTA */
    uint32_t inner_factor;      /* 1, or the graph's oversampling factor */
} oscen_b200_kernel_info;
int  oscen_b200_graph_kernel_info(const oscen_b200_graph *g, float sample_rate, oscen_b200_kernel_info *info,
                                  char *source, size_t source_cap, uint32_t *state_words, size_t state_cap);
/* the generated source of the variant oscen_b200_specialize() compiles for `uniform_mask` (0 = the general kernel) at this
   sample rate; returns its length (dst may be NULL to ask for it).  Host only. */
int  oscen_b200_graph_kernel_source(const oscen_b200_graph *g, float sample_rate, uint32_t uniform_mask, char *dst, size_t cap);

/* counters: kernels launched by this engine since create; bytes copied H2D / D2H */
typedef struct {
    uint64_t kernel_launches, graph_launches, h2d_bytes, d2h_bytes, blocks, events_dropped;
} oscen_b200_stats;
int  oscen_b200_get_stats(const oscen_b200_engine *e, oscen_b200_stats *out);
/* device time of the last block's compute kernels, CUDA events on the engine stream (ms) */
int  oscen_b200_last_block_ms(oscen_b200_engine *e, float *ms);
int  oscen_b200_enable_timing(oscen_b200_engine *e, int on);
/* device time of the last block's graph kernel alone (fused plan) / node sequence (node plan) */
int  oscen_b200_last_kernel_ms(oscen_b200_engine *e, float *ms);
/* the same two figures for EVERY block rendered since enable_timing(1) / the previous drain, oldest first, with one
   synchronisation for the whole batch (a timed loop then never waits on the device per step).  Either array may be
   NULL.  Returns the number of blocks written (<= cap); the engine keeps the last 4096 blocks. */
int  oscen_b200_timing_drain(oscen_b200_engine *e, float *block_ms, float *kernel_ms, uint32_t cap);
/* which graph kernel (variant) the last process_block launched, e.g. "fm_voice_seg_kernel<ramp=0,fmod=1>"; for the node
   plan the captured launch sequence's stateful kernels, e.g. "node plan (CUDA graph): adsr_node_plain_kernel
   fmop_node_tma_kernel ...".  Owned by the engine: valid until the next process_block / destroy */
const char *oscen_b200_last_kernel_name(const oscen_b200_engine *e);
/* benchmarking aid: overwrite a 256 MiB scratch buffer on the engine stream so the next block
   starts with a cold L2 (126 MB) */
int  oscen_b200_l2_flush(oscen_b200_engine *e);

/* multi-GPU: voices are sharded across ranks (one engine per GPU per process); the only
   exchange is the mix-bus all-reduce.  id is an ncclUniqueId (128 bytes) from rank 0. */
int  oscen_b200_comm_unique_id(void *id128);
int  oscen_b200_comm_init(oscen_b200_engine *e, int rank, int n_ranks, const void *id128);
/* NVLink peer-memory mix bus (preferred over NCCL for this <= 4 KB message): every rank exports
   a CUDA-IPC handle of its exchange buffer (64 bytes), the host all-gathers the handles, every
   rank imports all of them.  Afterwards process_block ends with ONE kernel that reduces the
   rank's partials, stores the result into every peer's buffer over NVLink, raises a per-block
   flag, waits for the peers' flags and folds the ranks' contributions in rank order
   (deterministic).  n_ranks <= 16. */
#define OSCEN_B200_IPC_HANDLE_BYTES 64
int  oscen_b200_comm_p2p_export(oscen_b200_engine *e, int n_ranks, void *handle64);
int  oscen_b200_comm_p2p_import(oscen_b200_engine *e, int rank, int n_ranks, const void *handles /* n_ranks x 64 */);
/* Leave the peer-memory exchange (closes the imported handles); the mix bus falls back to the NCCL
 * all-reduce of comm_init.  Collective by convention: every rank calls it or none does. */
int  oscen_b200_comm_p2p_close(oscen_b200_engine *e);

#ifdef __cplusplus
}
#endif
#endif /* OSCEN_B200_H */
