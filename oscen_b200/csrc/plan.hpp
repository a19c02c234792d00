/*
 * plan.hpp — the node-granular launch plan: a small host-side "graph compiler" back end.
 *
 * The reference's `graph!` macro lowers a graph description to IR (nodes, edges), sorts the nodes
 * topologically with Kahn's algorithm over the non-feedback edges
 * (oscen-graph-compiler/src/ir/lower.rs:1015-1085) and emits one `process()` call per node in
 * that order (codegen/mod.rs:499-514).  Here the same description (nodes with their stream
 * edges and value inputs, as written in the reference graph files) is sorted the same way and
 * becomes a CUDA-graph launch sequence: one kernel launch per node TYPE per ready wave (all
 * ready nodes of one type share a launch, blockIdx.y = node), streams in HBM between them.
 *
 * Any valid topological order is result-identical because every edge into a node is evaluated
 * immediately before that node runs; the reference's own order is not even
 * deterministic across compiles (its ready queue is seeded from a HashMap), ours is: ready
 * nodes are taken in declaration order.
 *
 * Host-only, no CUDA types: unit-tested on CPU through oscen_b200_plan_describe().
 */
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

namespace oscen {
namespace plan {

enum class NodeType : uint8_t {
    Adsr,          /* AdsrEnvelope                     value: attack decay sustain release; event: gate */
    FmOperator,    /* FmOperator                       stream: phase_mod, envelope; value: base_freq ratio feedback level */
    Tpt,           /* TptFilter                        stream: input, cutoff, f_mod; value: q */
    PolyBlepSine,  /* PolyBlepOscillator::sine         stream: frequency_mod, phase_mod; value: frequency amplitude pulse_width */
    PolyBlepSaw,   /* PolyBlepOscillator::saw */
    PolyBlepSquare,
    PolyBlepTriangle,
    Lp18,          /* LP18Filter                       stream: input; value: cutoff resonance fmod */
    Gain,          /* Gain: input * gain (gain = value input, constant, or a second stream) */
    AddValue,      /* AddValue: input + value */
    Mixer,         /* Mixer: a + b; also the auto-summing stream fan-in (emit_node.rs:153-174) */
    Sub,           /* compound source expression `a - b` (codegen/mod.rs:433-472) */
    Crossfade,     /* Crossfade: two outputs */
    HardClip,      /* HardClip */
    Expr,          /* one operator of a compound source expression; `op` says which (ExprOp) */
    IirLowpass,    /* IirLowpass (JUCE-style biquad)        stream: input; value: cutoff q */
    OscSine, OscSquare, OscSaw,   /* the naive Oscillator         stream: frequency_mod; value: frequency amplitude */
    Delay,         /* Delay over a power-of-two ring        stream: input; value: delay_samples feedback;
                      the one node type a feedback edge (`-> [N] ->`, `-> [name] ->`) may run through (AllowsFeedback);
                      graphs WITH a feedback edge run as generated kernels only */
    Up,            /* cross-rate stream edge outer -> inner (`* N` destination): StreamUpsampler of `policy`   */
    Down,          /* cross-rate stream edge inner -> outer: StreamDownsampler of `policy` (generated kernels only) */
    SamplePlayer,  /* SamplePlayer (mono): loops the buffer loaded into asset slot `asset` (`external s: AudioAsset; s -> p.buf`) */
    AmpSource,     /* electric-piano AmplitudeSource        value: frequency brightness velocity_scaling decay_rate harmonic_decay
                      key_scaling release_rate; event: gate; output `amplitudes`: an ARRAY-typed stream, [f32; 32] per frame */
    OscBank,       /* electric-piano OscillatorBank         stream: amplitudes ([f32; 32]); value: frequency; event: gate */
};
enum : int { kPolicySinc = 0, kPolicySincIir = 1, kPolicyLinear = 2, kPolicyLatch = 3 };   /* == OSCEN_B200_POLICY_* */

/* operators of compound connection sources (`a.x * gain + offset -> out`, `x.tanh() -> ...`):
   x = in[0], y = in[1] (stream), p = value input / literal (pid[0] / cst[0]), q = cst[1] */
enum class ExprOp : uint8_t {
    None,
    AddSS, SubSS, MulSS, DivSS,               /* stream (op) stream */
    AddSV, SubSV, MulSV, DivSV,               /* stream (op) value  */
    SubVS, DivVS,                             /* value (op) stream  */
    Value,                                    /* a value (graph input / literal) read as a stream */
    Tanh, Abs, Clamp, MinV, MaxV,             /* f32 methods: x.tanh(), x.abs(), x.clamp(p, q), x.min(p), x.max(p) */
    Sqrt, Sin, Cos, Exp,                      /* x.sqrt() (IEEE), x.sin() x.cos() x.exp() (libm sinf / cosf / expf, restated in om_math.h) */
};

constexpr int kMaxStreamIn = 3;
constexpr int kMaxParams = 8;

constexpr int kGraphInput = -2;      /* stream source: the graph's `audio_in` block */
constexpr int kNone = -1;
constexpr uint32_t kNoValue = 0xFFFFFFFFu;

struct StreamRef { int node = kNone; int port = 0; };   /* node index (or kGraphInput / kNone), output port */

struct Node {
    std::string name;
    NodeType type;
    StreamRef in[kMaxStreamIn];      /* stream inputs (meaning per type, see above) */
    bool fb[kMaxStreamIn] = {false, false, false};   /* the edge into in[k] is the outgoing leg of a feedback via: it reads the Delay's
                                        output of the PREVIOUS frame and does not order the nodes (ir/lower.rs:1015-1085 skips it) */
    uint32_t pid[kMaxParams];        /* value-input ids of the graph feeding this node's parameter slots (kNoValue: use cst).
                                        Slots [0, n_value) are the node's value endpoints; the slots after them are the
                                        stand-ins of its stream endpoints, read when in[k] is unconnected (constructor
                                        default) or fed by a graph VALUE input (Value -> Stream edge: a broadcast) */
    float cst[kMaxParams];           /* constructor value for unconnected slots */
    uint32_t state_base;             /* first SoA state field of this node (engine_types.h) */
    bool gate = true;                /* AdsrEnvelope: the graph's gate event input is connected */
    ExprOp op = ExprOp::None;
    float init[2] = {0.0f, 0.0f};    /* constructor arguments that seed the node's state (TptFilter / LP18Filter: cutoff, q) */
    int rate_log2 = 0;               /* 0: the graph's rate; L: the node runs 2^L times per frame (`= T::new() * N`) at sample rate sr * N */
    int policy = 0;                  /* Up / Down: connection policy (kPolicy*) */
    uint32_t ring_cap = 0;           /* Delay: ring capacity in samples (power of two), set by assign_state_layout() */
    float max_delay = 0.0f;          /* Delay: the reference ring's capacity - 1 (the clamp of apply_parameter_updates) */
    int asset = kNone;               /* SamplePlayer: index of the `external` asset slot connected to `buf` (kNone: never loaded, silent) */
};
/* lanes of a node's output stream: 1 for `f32`, N for an array-typed endpoint `[f32; N]` (oscen-lib/src/frame.rs, `#[output(stream)]
   pub amplitudes: [f32; NUM_HARMONICS]`); a wide stream only connects endpoints of the same width */
inline int stream_width(NodeType t, int /*port*/) { return t == NodeType::AmpSource ? 32 : 1; }
/* node types that exist in the generated (fused) kernel only: their streams / state do not fit the node plan's one-f32-per-voice buffers */
inline bool fused_only(NodeType t) { return t == NodeType::AmpSource || t == NodeType::OscBank; }

/* a value input of a DSL-compiled graph (`input cutoff: value = 3000.0 [20.0..20000.0, log, ramp: 1323];`) */
struct ValueInput {
    std::string name;
    float def = 0.0f;
    bool has_range = false; float lo = 0.0f, hi = 0.0f; bool log = false;
    uint32_t ramp = 0;               /* `ramp: N` frames of set_<v>() smoothing (codegen/mod.rs:917-976); 0 = plain f32 field */
};

struct Graph {
    std::string name;
    std::vector<Node> nodes;
    StreamRef output;                /* the stream connected to the graph's audio output */
    /* DSL-compiled graphs only (dsl.hpp); the built-in graph kinds keep their tables in engine.cu */
    std::vector<ValueInput> inputs;
    std::string stream_input, event_input, output_name;
    uint32_t n_state = 0;
    int inner_log2 = 0;              /* > 0: a multirate graph (one inner rate 2^L per graph); runs as a generated kernel only */
    /* polyphonic wrapper (`voices = [Voice::new(); N]` in an outer graph, e.g. FMGraph around FMVoice): the voice
       graph is what is planned; the outer graph contributes the polyphony, the MIDI front and the input names */
    uint32_t vpi = 0;                /* N of the voice array = the engine's voices_per_instance (0: not polyphonic) */
    bool midi = false;               /* midi_in -> MidiParser -> VoiceAllocator<N> -> [MidiVoiceHandler; N] feeds the voices */
    uint32_t midi_freq_input = 0;    /* the voice input `voice_handlers.frequency` is connected to */
    std::vector<std::pair<std::string, uint32_t>> aliases;   /* outer value-input name -> voice input id (broadcast edges) */
    std::vector<std::string> assets; /* `external NAME: AudioAsset;` slots in declaration order (engine: oscen_b200_load_asset) */
    /* instance-level tail of a polyphonic graph: `voices.output -> tremolo.input; tremolo.output -> out` with
       `output out: stream: Frame<2>` (examples/electric-piano/src/main.rs:33-97).  The voice sum of an instance runs through
       ONE Tremolo (tremolo.rs:5-61) whose output is a stereo frame: the engine's instance / mix outputs get 2 channels and
       one state word per instance (the LFO phase).  pid / cst: rate, depth (a graph value input, or the constructor's 5.0 / 0.5) */
    bool post_tremolo = false;
    uint32_t post_pid[2] = {kNoValue, kNoValue};
    float post_cst[2] = {5.0f, 0.5f};
    uint32_t channels = 1;           /* lanes of the graph's stream output: 1, or 2 for `stream: Frame<2>` */
};

struct Launch {
    NodeType type;
    std::vector<int> nodes;          /* 1..4 node indices rendered by this launch (blockIdx.y) */
};

struct Plan {
    std::vector<int> order;          /* topological order (node indices) */
    std::vector<Launch> launches;    /* grouped by ready wave and type */
    std::vector<int> first_stream;   /* per node: index of its first output stream buffer */
    int n_streams = 0;
};

inline int n_outputs(NodeType t) { return t == NodeType::Crossfade ? 2 : 1; }
/* SoA state words per voice of a node type (engine_types.h / nodes.cuh state structs) */
inline uint32_t n_state_words(const Node &nd) {
    const uint32_t l = (uint32_t)nd.rate_log2;     /* Up / Down: log2 of the edge's factor */
    switch (nd.type) {
    /* resampler state per connection policy (resample_edges.cuh; same layout as kernel_resample.cuh) */
    case NodeType::Up: return nd.policy == kPolicySinc ? 12u * l : nd.policy == kPolicySincIir ? 8u * l : nd.policy == kPolicyLinear ? 1u : 0u;
    case NodeType::Down: return nd.policy == kPolicySinc ? 24u * l : nd.policy == kPolicySincIir ? 9u * l : 0u;
    case NodeType::Adsr: return 4;            /* stage, rem, level, vel */
    case NodeType::FmOperator: return 2;      /* phase, prev */
    case NodeType::Tpt: return 7;             /* z0 z1 cur_cutoff cur_q h g k */
    case NodeType::PolyBlepSine: case NodeType::PolyBlepSaw: case NodeType::PolyBlepSquare: case NodeType::PolyBlepTriangle: return 1;
    case NodeType::Lp18: return 9;
    case NodeType::IirLowpass: return 8;      /* b0 b1 b2 a1 a2 v1 v2 frame_counter */
    case NodeType::OscSine: case NodeType::OscSquare: case NodeType::OscSaw: return 1;
    case NodeType::Delay: return 3u + nd.ring_cap;   /* write_pos frame_counter output + the ring */
    case NodeType::SamplePlayer: return 1;    /* playhead */
    case NodeType::AmpSource: return 130;     /* current / target / decay / release [32] + released + interpolation_step (nodes.cuh AS_*) */
    case NodeType::OscBank: return 129;       /* oscillators / multipliers re, im [32] + last_frequency (nodes.cuh OB_*) */
    default: return 0;
    }
}
inline bool node_plan_has_kernel(NodeType t) { return t != NodeType::Up && t != NodeType::Down; }
/* Delay::prepare (delay/mod.rs:59-69): RingBuffer::new(min((2 * sr) as usize, 88200)) in PowerOfTwo mode */
inline uint32_t delay_reference_capacity(float sample_rate) {
    const float t = 2.0f * sample_rate;
    uint32_t size = !(t > 0.0f) ? 0u : (t >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)t);
    if (size > 88200u) size = 88200u;
    if (size < 1u) size = 1u;
    uint32_t cap = 1u;
    while (cap < size) cap <<= 1;
    return cap;
}

/* Kahn's algorithm in waves: wave k holds every node whose producers all sit in earlier waves */
inline bool compile(const Graph &g, Plan &p, std::string &err) {
    const int n = (int)g.nodes.size();
    std::vector<int> indeg(n, 0);
    std::vector<std::vector<int>> users(n);
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < kMaxStreamIn; ++k) {
            const StreamRef &r = g.nodes[i].in[k];
            if (r.node >= 0) {
                if (r.node >= n || r.port >= n_outputs(g.nodes[r.node].type)) { err = "bad edge"; return false; }
                if (g.nodes[i].fb[k]) continue;          /* the outgoing leg of a feedback via does not order the walk */
                indeg[i]++; users[r.node].push_back(i);
            }
        }
    p = Plan();
    p.first_stream.assign(n, 0);
    for (int i = 0; i < n; ++i) { p.first_stream[i] = p.n_streams; p.n_streams += n_outputs(g.nodes[i].type); }
    std::vector<int> ready;
    for (int i = 0; i < n; ++i) if (indeg[i] == 0) ready.push_back(i);
    int done = 0;
    while (!ready.empty()) {
        std::vector<int> next;
        /* one launch per node type present in this wave (declaration order inside a launch) */
        std::vector<bool> taken(ready.size(), false);
        for (size_t a = 0; a < ready.size(); ++a) {
            if (taken[a]) continue;
            Launch l; l.type = g.nodes[ready[a]].type;
            for (size_t b = a; b < ready.size(); ++b)
                if (!taken[b] && g.nodes[ready[b]].type == l.type && l.nodes.size() < 4) { l.nodes.push_back(ready[b]); taken[b] = true; }
            p.launches.push_back(l);
        }
        for (int i : ready) {
            p.order.push_back(i); ++done;
            for (int u : users[i]) if (--indeg[u] == 0) next.push_back(u);
        }
        std::sort(next.begin(), next.end());
        ready.swap(next);
    }
    if (done != n) { err = "cycle in graph: close it through a delay (`a.output -> [N] -> b.input` or `-> [delay_node] ->`)"; return false; }
    if (g.output.node < 0 || g.output.node >= n) { err = "graph has no output"; return false; }
    return true;
}

inline const char *type_name(NodeType t) {
    switch (t) {
    case NodeType::Adsr: return "AdsrEnvelope";
    case NodeType::FmOperator: return "FmOperator";
    case NodeType::Tpt: return "TptFilter";
    case NodeType::PolyBlepSine: return "PolyBlepOscillator::sine";
    case NodeType::PolyBlepSaw: return "PolyBlepOscillator::saw";
    case NodeType::PolyBlepSquare: return "PolyBlepOscillator::square";
    case NodeType::PolyBlepTriangle: return "PolyBlepOscillator::triangle";
    case NodeType::Expr: return "Expr";
    case NodeType::IirLowpass: return "IirLowpass";
    case NodeType::OscSine: return "Oscillator::sine";
    case NodeType::OscSquare: return "Oscillator::square";
    case NodeType::OscSaw: return "Oscillator::saw";
    case NodeType::Delay: return "Delay";
    case NodeType::Up: return "UpEdge";
    case NodeType::Down: return "DownEdge";
    case NodeType::SamplePlayer: return "SamplePlayer";
    case NodeType::AmpSource: return "AmplitudeSource";
    case NodeType::OscBank: return "OscillatorBank";
    case NodeType::Lp18: return "LP18Filter";
    case NodeType::Gain: return "Gain";
    case NodeType::AddValue: return "AddValue";
    case NodeType::Mixer: return "Mixer";
    case NodeType::Sub: return "Sub(expr)";
    case NodeType::Crossfade: return "Crossfade";
    default: return "HardClip";
    }
}

/* SoA state layout (nodes in order) for a given sample rate: Delay rings are sized here — the reference capacity
   in general; a constant whole-sample delay only ever reads `delay` samples back through the exact-sample path
   (ring_buffer/mod.rs:178-190), so its ring shrinks to the next power of two above it.  Unconnected Delay
   parameters are clamped like the first apply_parameter_updates() does (delay/mod.rs:47-56). */
inline void assign_state_layout(Graph &g, float sample_rate) {
    uint32_t base = 0;
    for (Node &nd : g.nodes) {
        if (nd.type == NodeType::Delay) {
            const float sr = sample_rate * (float)(1 << nd.rate_log2);
            const uint32_t ref_cap = delay_reference_capacity(sr);
            nd.max_delay = (float)ref_cap - 1.0f;
            nd.ring_cap = ref_cap;
            if (nd.pid[1] == kNoValue) nd.cst[1] = nd.cst[1] < 0.0f ? 0.0f : (nd.cst[1] > 0.99f ? 0.99f : nd.cst[1]);
            if (nd.pid[0] == kNoValue) {
                float d = nd.cst[0];
                d = d < 0.0f ? 0.0f : (d > nd.max_delay ? nd.max_delay : d);     /* NaN stays NaN, as f32::clamp */
                nd.cst[0] = d;
                const float fr = d - (float)(uint32_t)d;
                if (d == d && (fr < 1e-6f || (1.0f - fr) < 1e-6f)) {
                    const uint32_t whole = (uint32_t)(d + 0.5f);
                    uint32_t cap = 1u;
                    while (cap < whole + 1u) cap <<= 1;
                    if (cap < ref_cap) nd.ring_cap = cap;
                }
            }
        }
        nd.state_base = base; base += n_state_words(nd);
    }
    g.n_state = base;
}

inline const char *op_name(ExprOp o) {
    static const char *const names[] = {"", "a + b", "a - b", "a * b", "a / b", "a + v", "a - v", "a * v", "a / v", "v - a", "v / a",
                                        "value", "tanh", "abs", "clamp", "min", "max", "sqrt", "sin", "cos", "exp"};
    return names[(int)o];
}

/* text form of a compiled plan: "launch k: Type node,node | ..." (for tests and DESIGN.md) */
inline std::string describe(const Graph &g, const Plan &p) {
    std::string s;
    for (size_t k = 0; k < p.launches.size(); ++k) {
        s += std::to_string(k) + ": " + type_name(p.launches[k].type) + " [";
        for (size_t j = 0; j < p.launches[k].nodes.size(); ++j) {
            if (j) s += ", ";
            s += g.nodes[p.launches[k].nodes[j]].name;
        }
        s += "]\n";
    }
    return s;
}

}  // namespace plan
}  // namespace oscen
