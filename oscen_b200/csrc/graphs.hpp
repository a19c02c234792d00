/*
 * graphs.hpp — the reference graphs on the hot path written as node/edge descriptions, the way
 * their `graph!` blocks declare them (node names, connection lists), for the plan compiler in
 * plan.hpp.  Multirate (SatGraph_4x) and array-stream (ElectricPiano) graphs run only as fused
 * kernels.
 */
#pragma once
#include "../../include/oscen_b200.h"
#include "engine_types.h"
#include "plan.hpp"

namespace oscen {
namespace plan {

inline Node make(const char *name, NodeType t, StreamRef a, StreamRef b, uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3,
                 float c0, float c1, float c2, float c3, uint32_t state_base) {
    Node n{};
    n.name = name; n.type = t; n.in[0] = a; n.in[1] = b;
    for (int k = 0; k < kMaxParams; ++k) { n.pid[k] = kNoValue; n.cst[k] = 0.0f; }
    n.pid[0] = p0; n.pid[1] = p1; n.pid[2] = p2; n.pid[3] = p3;
    n.cst[0] = c0; n.cst[1] = c1; n.cst[2] = c2; n.cst[3] = c3;
    if (t == NodeType::PolyBlepSine || t == NodeType::PolyBlepSaw) n.cst[2] = 0.5f;    /* pulse_width, oscillators/mod.rs:116 */
    if (t == NodeType::FmOperator) n.cst[5] = 1.0f;                                     /* envelope, fm_operator.rs:41 */
    n.state_base = state_base;
    return n;
}
constexpr uint32_t NV = kNoValue;
inline StreamRef out(int node, int port = 0) { StreamRef r; r.node = node; r.port = port; return r; }
inline StreamRef none() { return StreamRef(); }
inline StreamRef graph_in() { StreamRef r; r.node = kGraphInput; return r; }

/* FMVoice — examples/fm-synth/src/fm_voice.rs:52-155 (nodes in declaration order) */
inline Graph fm_voice() {
    enum { env3, env2, env1, env_filter, filter_env_gain, cutoff_mod, op3_osc, op2_osc, op1_osc, op3_route,
           op1_mod_mixer, filter, output_gain };
    Graph g;
    auto &N = g.nodes;
    N.push_back(make("env3", NodeType::Adsr, none(), none(), OSCEN_B200_FM_OP3_ATTACK, OSCEN_B200_FM_OP3_DECAY, OSCEN_B200_FM_OP3_SUSTAIN, OSCEN_B200_FM_OP3_RELEASE, 0, 0, 0, 0, FM_E3));
    N.push_back(make("env2", NodeType::Adsr, none(), none(), OSCEN_B200_FM_OP2_ATTACK, OSCEN_B200_FM_OP2_DECAY, OSCEN_B200_FM_OP2_SUSTAIN, OSCEN_B200_FM_OP2_RELEASE, 0, 0, 0, 0, FM_E2));
    N.push_back(make("env1", NodeType::Adsr, none(), none(), OSCEN_B200_FM_OP1_ATTACK, OSCEN_B200_FM_OP1_DECAY, OSCEN_B200_FM_OP1_SUSTAIN, OSCEN_B200_FM_OP1_RELEASE, 0, 0, 0, 0, FM_E1));
    N.push_back(make("env_filter", NodeType::Adsr, none(), none(), OSCEN_B200_FM_FILTER_ATTACK, OSCEN_B200_FM_FILTER_DECAY, OSCEN_B200_FM_FILTER_SUSTAIN, OSCEN_B200_FM_FILTER_RELEASE, 0, 0, 0, 0, FM_EF));
    /* env_filter.output -> filter_env_gain.input; filter_env_amount -> filter_env_gain.gain */
    N.push_back(make("filter_env_gain", NodeType::Gain, out(env_filter), none(), OSCEN_B200_FM_FILTER_ENV_AMOUNT, NV, NV, NV, 0, 0, 0, 0, 0));
    /* filter_env_gain.output -> cutoff_mod.input; filter_cutoff -> cutoff_mod.value */
    N.push_back(make("cutoff_mod", NodeType::AddValue, out(filter_env_gain), none(), OSCEN_B200_FM_FILTER_CUTOFF, NV, NV, NV, 0, 0, 0, 0, 0));
    /* op3: frequency, op3_ratio, op3_feedback, env3.output -> envelope, op3_level; phase_mod unconnected */
    N.push_back(make("op3_osc", NodeType::FmOperator, none(), out(env3), OSCEN_B200_FM_FREQUENCY, OSCEN_B200_FM_OP3_RATIO, OSCEN_B200_FM_OP3_FEEDBACK, OSCEN_B200_FM_OP3_LEVEL, 0, 0, 0, 0, FM_OP3));
    /* op2: phase_mod <- op3_route.output_a */
    N.push_back(make("op2_osc", NodeType::FmOperator, out(op3_route, 0), out(env2), OSCEN_B200_FM_FREQUENCY, OSCEN_B200_FM_OP2_RATIO, OSCEN_B200_FM_OP2_FEEDBACK, OSCEN_B200_FM_OP2_LEVEL, 0, 0, 0, 0, FM_OP2));
    /* op1: phase_mod <- op1_mod_mixer.output; feedback, level never connected (0.0, 1.0) */
    N.push_back(make("op1_osc", NodeType::FmOperator, out(op1_mod_mixer), out(env1), OSCEN_B200_FM_FREQUENCY, OSCEN_B200_FM_OP1_RATIO, NV, NV, 0, 0, 0.0f, 1.0f, FM_OP1));
    N.push_back(make("op3_route", NodeType::Crossfade, out(op3_osc), none(), OSCEN_B200_FM_ROUTE, NV, NV, NV, 0, 0, 0, 0, 0));
    N.push_back(make("op1_mod_mixer", NodeType::Mixer, out(op2_osc), out(op3_route, 1), NV, NV, NV, NV, 0, 0, 0, 0, 0));
    /* cutoff_mod.output -> filter.cutoff; op1_osc.output -> filter.input; filter_resonance -> filter.q */
    N.push_back(make("filter", NodeType::Tpt, out(op1_osc), out(cutoff_mod), OSCEN_B200_FM_FILTER_RESONANCE, NV, NV, NV, 0, 0, 0, 0, FM_TPT));
    N.push_back(make("output_gain", NodeType::Gain, out(filter), none(), NV, NV, NV, NV, 0.3f, 0, 0, 0, 0));
    g.output = out(output_gain);
    return g;
}

/* README `Synth` (README.md:24-53) and the +ADSR->vca variant (benches/static_vs_runtime.rs:57-59) */
inline Graph synth(bool with_adsr) {
    enum { modulator, carrier, filter, amp_env, vca };
    Graph g;
    auto &N = g.nodes;
    N.push_back(make("modulator", NodeType::PolyBlepSine, none(), none(), OSCEN_B200_SYNTH_MOD_FREQ, OSCEN_B200_SYNTH_MOD_DEPTH, NV, NV, 0, 0, 0, 0, SY_MOD));
    /* modulator.output -> carrier.frequency_mod; carrier amplitude is the constructor's 0.5 */
    N.push_back(make("carrier", NodeType::PolyBlepSaw, out(modulator), none(), OSCEN_B200_SYNTH_CARRIER_FREQ, NV, NV, NV, 0, 0.5f, 0, 0, SY_CAR));
    /* cutoff (value) -> filter.cutoff (stream endpoint): pid[1]; q stays 0.707 */
    N.push_back(make("filter", NodeType::Tpt, out(carrier), none(), NV, OSCEN_B200_SYNTH_CUTOFF, NV, NV, 0.707f, 0, 0, 0, SY_TPT));
    g.output = out(filter);
    if (with_adsr) {
        N.push_back(make("amp_env", NodeType::Adsr, none(), none(), OSCEN_B200_SYNTH_ATTACK, OSCEN_B200_SYNTH_DECAY, OSCEN_B200_SYNTH_SUSTAIN, OSCEN_B200_SYNTH_RELEASE, 0, 0, 0, 0, SY_ENV));
        /* filter.output -> vca.input; amp_env.output -> vca.gain */
        N.push_back(make("vca", NodeType::Gain, out(filter), out(amp_env), NV, NV, NV, NV, 0, 0, 0, 0, 0));
        g.output = out(vca);
    }
    return g;
}

/* TwinPeaksGraph — examples/nih-twin-peaks/src/lib.rs:15-48: `filter_a.output - filter_b.output -> audio_out` */
inline Graph twin_peaks() {
    enum { filter_a, filter_b, diff };
    Graph g;
    auto &N = g.nodes;
    N.push_back(make("filter_a", NodeType::Lp18, graph_in(), none(), OSCEN_B200_TWIN_CUTOFF_A, OSCEN_B200_TWIN_RESONANCE, NV, NV, 0, 0, 0, 0, TW_A));
    N.push_back(make("filter_b", NodeType::Lp18, graph_in(), none(), OSCEN_B200_TWIN_CUTOFF_B, OSCEN_B200_TWIN_RESONANCE, NV, NV, 0, 0, 0, 0, TW_B));
    N.push_back(make("filter_a.output - filter_b.output", NodeType::Sub, out(filter_a), out(filter_b), NV, NV, NV, NV, 0, 0, 0, 0, 0));
    g.output = out(diff);
    return g;
}

/* SatGraph_1x — examples/oversampled-saturator/src/main.rs:64-80 with FACTOR = 1 */
inline Graph sat_1x() {
    enum { osc, clip };
    Graph g;
    auto &N = g.nodes;
    N.push_back(make("osc", NodeType::PolyBlepSaw, none(), none(), OSCEN_B200_SAT_FREQUENCY, OSCEN_B200_SAT_AMPLITUDE, NV, NV, 0, 0, 0, 0, SAT_PHASE));
    N.push_back(make("clip", NodeType::HardClip, out(osc), none(), NV, NV, NV, NV, 0, 0, 0, 0, 0));
    g.output = out(clip);
    return g;
}

/* graphs that have a node plan; returns false for multirate / array-stream graphs */
inline bool for_kind(uint32_t kind, Graph &g) {
    switch (kind) {
    case OSCEN_B200_GRAPH_FM_VOICE: g = fm_voice(); return true;
    case OSCEN_B200_GRAPH_SYNTH: g = synth(false); return true;
    case OSCEN_B200_GRAPH_SYNTH_ADSR: g = synth(true); return true;
    case OSCEN_B200_GRAPH_TWIN_PEAKS: g = twin_peaks(); return true;
    case OSCEN_B200_GRAPH_SAT_1X: g = sat_1x(); return true;
    default: return false;
    }
}

}  // namespace plan
}  // namespace oscen
