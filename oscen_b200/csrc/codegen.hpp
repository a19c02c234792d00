/*
 * codegen.hpp — back end of the plan compiler for DSL graphs: plan::Graph -> CUDA source of ONE fused
 * kernel (one thread = one voice, every node's state in registers for the whole block, streams are
 * local floats), compiled at run time with NVRTC for sm_100a (jit.hpp).
 *
 * This is the B200 counterpart of what `graph!` does at Rust compile time
 * (oscen-graph-compiler/src/codegen/mod.rs:499-572, emit_node.rs:182-379): it emits, per frame,
 *   tick_ramps -> value edges -> process_event_inputs (gate handlers) -> node.process() in
 *   topological order -> graph output
 * with every node body inlined.  The emitted code calls the SAME device functions as the hand-written
 * fused kernels and the node kernels (nodes.cuh), is compiled with --fmad=false like they are, and
 * evaluates every edge expression with one f32 rounding per operator — so a generated kernel is
 * bit-identical to the node plan of the same graph (tests/test_dsl_gpu.py).
 *
 * Host-only C++; the generated text can be inspected through oscen_b200_graph_codegen().
 */
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include <cstring>
#include <string>
#include <vector>

#include "plan.hpp"

namespace oscen {
namespace codegen {

/* printf into a std::string: a line can never be cut short (an unterminated trailing comment would swallow the next line) */
inline std::string fmt(const char *f, ...) {
    va_list ap, ap2;
    va_start(ap, f); va_copy(ap2, ap);
    const int n = vsnprintf(nullptr, 0, f, ap);
    va_end(ap);
    std::string out((size_t)(n > 0 ? n : 0), '\0');
    if (n > 0) vsnprintf(&out[0], (size_t)n + 1, f, ap2);
    va_end(ap2);
    return out;
}

inline std::string lit(float v) {
    uint32_t u; memcpy(&u, &v, 4);
    char b[64];
    snprintf(b, sizeof b, "__uint_as_float(0x%08xu /* %.9g */)", u, (double)v);
    return b;
}

#ifndef OSCEN_B200_JIT_ARRAY_SMEM_DEFAULT
#define OSCEN_B200_JIT_ARRAY_SMEM_DEFAULT 2
#endif
struct Emitter {
    const plan::Graph &g;
    const plan::Plan &p;
    std::string s;
    explicit Emitter(const plan::Graph &g_, const plan::Plan &p_) : g(g_), p(p_) {}

    void line(int indent, const std::string &t) { s.append((size_t)indent * 4, ' '); s += t; s += '\n'; }
    std::string N(int i) const { return "n" + std::to_string(i); }
    std::string D(int i) const { return "d" + std::to_string(i); }
    std::string S(int i, int port = 0) const { return "s" + std::to_string(p.first_stream[i] + port); }
    /* parameter slot k of node i: a value input of the graph, or the constructor's literal */
    std::string P(int i, int k) const {
        const plan::Node &nd = g.nodes[i];
        if (nd.pid[k] == plan::kNoValue) return lit(nd.cst[k]);
        /* specialised variant: a uniform input is read from the kernel-argument constant bank at every use */
        if ((uniform_mask >> nd.pid[k]) & 1u) return "a.u.p[" + std::to_string(nd.pid[k]) + "]";
        return "P[" + std::to_string(nd.pid[k]) + "]";
    }
    bool P_is_const(int i, int k) const { return g.nodes[i].pid[k] == plan::kNoValue; }
    /* stream input k of node i; `standin` = the parameter slot read when no stream is connected */
    std::string IN(int i, int k, int standin) const {
        const plan::StreamRef &r = g.nodes[i].in[k];
        if (r.node >= 0) return S(r.node, r.port);
        if (r.node == plan::kGraphInput) return "x_in";
        return standin >= 0 ? P(i, standin) : std::string("0.0f");
    }
    static const char *wave(plan::NodeType t) {
        switch (t) {
        case plan::NodeType::PolyBlepSine: return "WAVE_SINE";
        case plan::NodeType::PolyBlepSaw: return "WAVE_SAW";
        case plan::NodeType::PolyBlepSquare: return "WAVE_SQUARE";
        default: return "WAVE_TRIANGLE";
        }
    }
    /* sample rate a node runs at: inner-rate (`* N`) nodes get sr * N (codegen/emit_struct.rs:578-581) */
    std::string SR(int i) const {
        const plan::Node &nd = g.nodes[i];
        return (nd.rate_log2 > 0 && nd.type != plan::NodeType::Up && nd.type != plan::NodeType::Down) ? "sri" : "sr";
    }
    /* multirate frame body phases (codegen/emit_frame.rs:114-176): 0 outer nodes before the inner loop, 1 Up edges,
       2 inner-rate nodes, 3 Down edges, 4 outer nodes that depend on a Down edge */
    std::vector<int> phases() const {
        const int n = (int)g.nodes.size();
        std::vector<char> tainted(n, 0);
        bool moved = true;
        while (moved) {
            moved = false;
            for (int i = 0; i < n; ++i) {
                if (tainted[i]) continue;
                bool t = g.nodes[i].type == plan::NodeType::Down;
                for (const plan::StreamRef &r : g.nodes[i].in) if (r.node >= 0 && tainted[r.node]) t = true;
                if (t) { tainted[i] = 1; moved = true; }
            }
        }
        std::vector<int> ph(n, 0);
        for (int i = 0; i < n; ++i) {
            const plan::Node &nd = g.nodes[i];
            if (nd.type == plan::NodeType::Up) ph[i] = 1;
            else if (nd.type == plan::NodeType::Down) ph[i] = 3;
            else if (nd.rate_log2 > 0) ph[i] = 2;
            else ph[i] = tainted[i] ? 4 : 0;
        }
        return ph;
    }
    std::string edge_type(int i) const {
        const plan::Node &nd = g.nodes[i];
        return std::string(nd.type == plan::NodeType::Up ? "UpEdge<" : "DownEdge<") + std::to_string(nd.policy) + ", " + std::to_string(nd.rate_log2) + ">";
    }
    static uint32_t edge_smem_floats(const plan::Node &) {
        return 0u;       /* [sinc] edges keep their FIR histories in registers since round 2 (nodes.cuh HbUpRegs / HbDownRegs) */
    }
    /* array-stream nodes (AmplitudeSource 130 words, OscillatorBank 129): their state is walked in place every frame, so for
       the block it lives in a [word][thread] slice of dynamic shared memory (conflict-free: consecutive threads, consecutive
       banks) and goes back to its SoA rows at the end — as long as the CTA's slices fit; beyond that the rows are walked in
       HBM (round-2 first form, ~5x slower) */
    /* OSCEN_B200_JIT_ARRAY_SMEM = 0: rows walked in HBM; 1: every word in shared memory; 2: the hot words (nodes.cuh AS_NHOT / OB_NHOT:
       what each frame reads AND writes) in shared memory, the read-mostly cold words (decay / release factors, rotation
       multipliers) in their HBM rows behind L1 — half the shared memory per CTA, so a third CTA fits an SM */
    static int array_mode() { const char *e = getenv("OSCEN_B200_JIT_ARRAY_SMEM"); return e && *e ? atoi(e) : OSCEN_B200_JIT_ARRAY_SMEM_DEFAULT; }
    static uint32_t array_state_words(const plan::Node &nd) {
        const bool hot_only = array_mode() == 2;
        return nd.type == plan::NodeType::AmpSource ? (hot_only ? 66u : 130u) : (nd.type == plan::NodeType::OscBank ? (hot_only ? 65u : 129u) : 0u);
    }
    static uint32_t array_smem_floats(const plan::Graph &g) {
        uint32_t w = 0;
        for (const plan::Node &nd : g.nodes) w += array_state_words(nd);
        if (array_mode() == 0) return 0u;
        return (size_t)w * 64u * sizeof(float) <= 160u * 1024u ? w : 0u;
    }
    static bool is_polyblep(plan::NodeType t) {
        return t == plan::NodeType::PolyBlepSine || t == plan::NodeType::PolyBlepSaw || t == plan::NodeType::PolyBlepSquare || t == plan::NodeType::PolyBlepTriangle;
    }

    /* tuning knobs (A/B measurements: tools/bench_dsl.py under different environments) */
    static int env_int(const char *name, int def) { const char *e = getenv(name); return e && *e ? atoi(e) : def; }
    bool hoist_adsr = true;
    int min_blocks = 8;
    uint32_t uniform_mask = 0;       /* the SPECIALISED variant (oscen_b200_specialize): these value inputs hold one value for every
                                        voice, and are neither ramped nor written by value events in the blocks this kernel renders (the
                                        engine checks that per block and falls back to the general variant); they are read from the
                                        kernel-argument constant bank at every use instead of occupying a register each.
                                        OSCEN_B200_JIT_UNIFORM_MASK=0x... ORs bits in for A/B runs (tools/ab_uniform.py) */

    std::string run(const char *kernel_name) {
        using T = plan::NodeType;
        hoist_adsr = env_int("OSCEN_B200_JIT_ADSR_HOIST", 1) != 0;
        min_blocks = env_int("OSCEN_B200_JIT_MIN_BLOCKS", 8);
        { const char *e = getenv("OSCEN_B200_JIT_UNIFORM_MASK"); if (e && *e) uniform_mask |= (uint32_t)strtoul(e, nullptr, 0); }
        const int n_nodes = (int)g.nodes.size();
        const uint32_t ni = (uint32_t)g.inputs.size();
        const bool has_in = !g.stream_input.empty();
        line(0, "/* generated by oscen_b200 codegen.hpp from graph `" + (g.name.empty() ? std::string("<anonymous>") : g.name) + "` — do not edit */");
        const bool multirate = g.inner_log2 > 0;
        const int NI = 1 << g.inner_log2;
        line(0, multirate ? "#include \"resample_edges.cuh\"" : "#include \"state_io.cuh\"");
        line(0, "using namespace oscen;");
        /* min_blocks CTAs of 64 threads per SM: 8 caps the kernel at 128 registers per thread, so that the 1024 CTAs of a
           65 536-voice bank (6.9 per SM) are resident in one wave */
        const uint32_t arr_floats = array_smem_floats(g);
        const bool arr_smem = arr_floats > 0u;
        const std::string arr_stride = arr_smem ? "(size_t)kCta" : "(size_t)n";
        const std::string cold_stride = (arr_smem && array_mode() != 2) ? "(size_t)kCta" : "(size_t)n";
        auto cold = [&](int i) { return "c" + std::to_string(i) + ", " + cold_stride; };
        if (arr_smem) {            /* CTAs per SM are set by the shared-memory slices: let the register allocation follow */
            const int fit = (int)((220u * 1024u) / ((size_t)arr_floats * 64u * sizeof(float) + 26u * 1024u));
            min_blocks = fit < 1 ? 1 : (fit < min_blocks ? fit : min_blocks);
        }
        line(0, fmt("extern \"C\" __global__ void __launch_bounds__(kCta, %d) %s(const __grid_constant__ KernelArgs a) {", min_blocks, kernel_name));
        line(1, "__shared__ float tile[kTileF * kTileW];");
        /* (a name of its own for the stereo scratch: the host-executed test harness loads many generated kernels into one
           process, where a function-local static of an inline function is ONE object per name) */
        line(1, g.channels == 2 ? "__shared__ float isum2[kTileF * kCta * 2];    /* stereo instance frames */" : "__shared__ float isum[kTileF * kCta];");
        line(1, "const uint32_t n = a.n_voices;");
        line(1, "const uint32_t v = blockIdx.x * kCta + threadIdx.x;");
        line(1, "const bool active = v < n;");
        line(1, "const uint32_t vs = active ? v : 0u;");
        line(1, "const uint32_t frames = a.ctl->frames, ramp_mask = a.ctl->ramp_mask;");
        line(1, "const float sr = a.u.sr;");
        line(1, "const uint32_t col = tile_col(threadIdx.x, a.vpi);");
        line(1, "(void)sr; (void)ramp_mask;");
        if (arr_smem && !multirate) line(1, "extern __shared__ float dyn[];                 /* this block's state of the array-stream nodes, [word][thread] */");
        if (multirate) {
            line(1, "extern __shared__ float dyn[];                 /* FIR rings of the [sinc] cross-rate edges */");
            line(1, "const float sri = sr * " + std::to_string(NI) + ".0f;    /* inner-rate nodes: set_sample_rate(sr * N) */");
            line(1, "(void)sri; (void)dyn;");
        }
        /* ---- value inputs */
        line(1, fmt("float P[%u];", ni ? ni : 1u));
        for (uint32_t i = 0; i < ni; ++i) {
            if ((uniform_mask >> i) & 1u) { line(1, fmt("P[%u] = 0.0f;    /* %s: uniform, read from a.u.p[%u] */", i, g.inputs[i].name.c_str(), i)); continue; }
            line(1, fmt("P[%u] = a.pv[(size_t)%uu * n + vs];    /* %s */", i, i, g.inputs[i].name.c_str()));
        }
        if (!ni) line(1, "P[0] = 0.0f;");
        /* ---- node state: HBM -> registers */
        uint32_t smem_off = 0;
        for (int i = 0; i < n_nodes; ++i) {
            const plan::Node &nd = g.nodes[i];
            const std::string base = std::to_string(nd.state_base) + "u";
            line(1, "/* " + nd.name + " */");
            switch (nd.type) {
            case T::Adsr:
                line(1, "AdsrState " + N(i) + "; adsr_load(" + N(i) + ", a.state, " + base + ", n, vs);");
                line(1, "AdsrDerived " + D(i) + "; adsr_derive_r(" + D(i) + ", " + P(i, 0) + ", " + P(i, 1) + ", " + P(i, 3) + ", " + SR(i) + ");");
                /* update_sustain_level()'s sustain-level / countdown-clamp part only changes with its inputs: refreshed at block
                   start and on frames where a value input moved or a gate arrived (nodes.cuh adsr_refresh), not every sample */
                if (hoist_adsr) line(1, "AdsrHot h" + std::to_string(i) + "; adsr_refresh<false>(" + N(i) + ", h" + std::to_string(i) + ", " + P(i, 2) + ", " + D(i) + ");");
                break;
            case T::FmOperator:
                line(1, "FmOpState " + N(i) + "; " + N(i) + ".phase = a.state[(size_t)(" + base + " + 0u) * n + vs]; " + N(i) + ".prev = a.state[(size_t)(" + base + " + 1u) * n + vs];");
                line(1, N(i) + ".inc = fmop_phase_inc(" + P(i, 0) + ", " + P(i, 1) + ", " + SR(i) + ");");
                break;
            case T::Tpt: line(1, "TptState " + N(i) + "; tpt_load(" + N(i) + ", a.state, " + base + ", n, vs);"); break;
            case T::Lp18: line(1, "Lp18State " + N(i) + "; lp18_load(" + N(i) + ", a.state, " + base + ", n, vs);"); break;
            case T::IirLowpass:
                line(1, "IirLpState " + N(i) + "; { const float *st_ = a.state + (size_t)" + base + " * n + vs; " + N(i) + ".b0 = st_[0]; " + N(i) + ".b1 = st_[n]; " + N(i) +
                        ".b2 = st_[2 * (size_t)n]; " + N(i) + ".a1 = st_[3 * (size_t)n]; " + N(i) + ".a2 = st_[4 * (size_t)n]; " + N(i) + ".v1 = st_[5 * (size_t)n]; " + N(i) +
                        ".v2 = st_[6 * (size_t)n]; " + N(i) + ".counter = __float_as_uint(st_[7 * (size_t)n]); }");
                break;
            case T::OscSine: case T::OscSquare: case T::OscSaw:
                line(1, "float " + N(i) + " = a.state[(size_t)" + base + " * n + vs];");
                break;
            case T::Delay:
                line(1, "DelayState " + N(i) + "; " + N(i) + ".write_pos = __float_as_uint(a.state[(size_t)(" + base + " + 0u) * n + vs]); " + N(i) +
                        ".counter = __float_as_uint(a.state[(size_t)(" + base + " + 1u) * n + vs]); " + N(i) + ".output = a.state[(size_t)(" + base + " + 2u) * n + vs];");
                line(1, "float *ring" + std::to_string(i) + " = a.state + (size_t)(" + base + " + 3u) * n + vs;    /* [" + std::to_string(nd.ring_cap) + " slots][voices] */");
                line(1, "float " + S(i) + " = " + N(i) + ".output;    /* persists across frames: a feedback edge reads the previous frame's value */");
                break;
            case T::SamplePlayer: {
                /* the playhead restarts when the slot was (re)loaded since the last block; a slot that shrank below an old
                   playhead cannot be read out of range either way */
                const std::string k = std::to_string(nd.asset);
                if (nd.asset >= 0) {
                    line(1, "const float *buf" + std::to_string(i) + " = a.assets->data[" + k + "]; const uint32_t len" + std::to_string(i) + " = a.assets->frames[" + k + "];");
                    line(1, "uint32_t " + N(i) + " = a.assets->reset[" + k + "] ? 0u : __float_as_uint(a.state[(size_t)" + base + " * n + vs]);");
                    line(1, "if (" + N(i) + " >= len" + std::to_string(i) + ") " + N(i) + " = 0u;");
                } else {
                    line(1, "const float *buf" + std::to_string(i) + " = nullptr; const uint32_t len" + std::to_string(i) + " = 0u; uint32_t " + N(i) + " = 0u;    /* `buf` is not connected: silence */");
                }
            } break;
            case T::AmpSource: case T::OscBank:
                /* 130 / 129 state words, walked in place (nodes.cuh), not held in registers */
                if (arr_smem) {
                    const std::string w = std::to_string(array_state_words(nd)) + "u";
                    line(1, "float *" + N(i) + " = dyn + " + std::to_string(smem_off) + "u * kCta + threadIdx.x;    /* [word][thread] slice of shared memory */");
                    line(1, "_Pragma(\"unroll 8\") for (uint32_t w_ = 0; w_ < " + w + "; ++w_) " + N(i) + "[(size_t)w_ * kCta] = a.state[(size_t)(" + base + " + w_) * n + vs];");
                    smem_off += array_state_words(nd);
                    if (array_mode() == 2) line(1, "float *c" + std::to_string(i) + " = a.state + (size_t)" + base + " * n + vs;    /* cold words: their SoA rows in HBM */");
                    else line(1, "float *c" + std::to_string(i) + " = " + N(i) + ";");
                } else {
                    line(1, "float *" + N(i) + " = a.state + (size_t)" + base + " * n + vs;    /* walked in its SoA rows in HBM */");
                    line(1, "float *c" + std::to_string(i) + " = " + N(i) + ";");
                }
                break;
            case T::Up: case T::Down:
                line(1, edge_type(i) + " " + N(i) + "; " + N(i) + ".load(a.state, " + base + ", n, vs, dyn + " + std::to_string(smem_off) + "u * kCta);");
                smem_off += edge_smem_floats(nd);
                break;
            default:
                if (is_polyblep(nd.type)) line(1, "float " + N(i) + " = a.state[(size_t)" + base + " * n + vs];");
                break;
            }
        }
        if (has_in) {
            const std::string f = std::to_string(g.n_state) + "u";
            line(1, "const bool lcg_on = a.ctl->lcg_on != 0u;");
            line(1, "unsigned long long lcg = (unsigned long long)__float_as_uint(a.state[(size_t)(" + f + " + 0u) * n + vs]) |");
            line(1, "                         ((unsigned long long)__float_as_uint(a.state[(size_t)(" + f + " + 1u) * n + vs]) << 32);");
        }
        if (g.post_tremolo) {
            line(1, "/* instance-level tail: voice sum -> Tremolo -> Frame<2>; one LFO phase per instance */");
            line(1, "const uint32_t my_inst = blockIdx.x * (kCta / a.vpi) + threadIdx.x;");
            line(1, "float trem_phase = (threadIdx.x < kCta / a.vpi && my_inst < a.n_inst) ? a.inst_state[my_inst] : 0.0f;");
        }
        {   /* FmOperator: the sine's three addend constants are loaded once and stay in registers (om_math.h om_sinf_rk) */
            bool any_fm = false;
            for (int i = 0; i < n_nodes; ++i) any_fm |= g.nodes[i].type == T::FmOperator;
            if (any_fm) line(1, "const OmSinRegs sk = om_sin_regs();");
        }
        line(1, "EventCursor ec; ec.init(a.events, active ? a.ctl->n_events : 0u, vs);");
        {   /* a gate-off at frame 0 starts from the release length the previous block ended with (AdsrState::r_gate) */
            bool any = false;
            for (int i = 0; i < n_nodes; ++i) if (g.nodes[i].type == T::Adsr) {
                if (!any) line(1, "const bool ev0 = ec.due(0u);");
                any = true;
                line(1, "adsr_begin_block(" + N(i) + ", " + D(i) + ", ev0);");
            }
        }
        /* ---- the walk: nodes in topological order (multirate graphs: per phase, emit_frame.rs:114-176) */
        bool plain_mode = false;      /* emitting the plain-segment copy of the walk (segmented frame loop, below) */
        auto emit_node = [&](int i, int ind) {
            const plan::Node &nd = g.nodes[i];
            const std::string o = "const float " + S(i) + " = ";
            const std::string cm = "    /* " + nd.name + " */";
            switch (nd.type) {
            case T::Adsr:
                if (plain_mode) { const std::string k = std::to_string(i);
                    line(ind, o + "any_rl ? adsr_plain_rel(" + N(i) + ", ad" + k + ", rl" + k + ", tg" + k + ", kk" + k + ") : adsr_ad(" + N(i) + ", ad" + k + ", tg" + k + ", kk" + k + ");" + cm); }
                else if (hoist_adsr) line(ind, o + "adsr_process_t<false>(" + N(i) + ", h" + std::to_string(i) + ", " + P(i, 2) + ", " + D(i) + ");" + cm);
                else line(ind, o + "adsr_process(" + N(i) + ", " + P(i, 2) + ", " + D(i) + ");" + cm);
                break;
            case T::FmOperator: line(ind, o + "fmop_process_k(" + N(i) + ", " + IN(i, 0, 4) + ", " + P(i, 2) + ", " + IN(i, 1, 5) + ", " + P(i, 3) + ", sk);" + cm); break;
            case T::Tpt: line(ind, o + "tpt_process(" + N(i) + ", " + IN(i, 0, 3) + ", " + IN(i, 1, 1) + ", " + P(i, 0) + ", " + IN(i, 2, 2) + ", " + SR(i) + ");" + cm); break;
            case T::Lp18: line(ind, o + "lp18_process(" + N(i) + ", " + IN(i, 0, 3) + ", " + P(i, 0) + ", " + P(i, 2) + ", " + P(i, 1) + ", " + SR(i) + ");" + cm); break;
            case T::IirLowpass: line(ind, o + "iirlp_process(" + N(i) + ", " + IN(i, 0, 2) + ", " + P(i, 0) + ", " + P(i, 1) + ", " + SR(i) + ");" + cm); break;
            case T::OscSine: case T::OscSquare: case T::OscSaw:
                line(ind, o + "osc_process<" + (nd.type == T::OscSine ? "OSC_SINE" : nd.type == T::OscSquare ? "OSC_SQUARE" : "OSC_SAW") + ">(" + N(i) + ", " + P(i, 0) + ", " +
                          IN(i, 0, 2) + ", " + P(i, 1) + ", " + SR(i) + ");" + cm);
                break;
            case T::Delay:
                line(ind, S(i) + " = delay_process<" + (P_is_const(i, 0) ? "false" : "true") + ", " + (P_is_const(i, 1) ? "false" : "true") + ">(" + N(i) + ", ring" + std::to_string(i) +
                          ", (size_t)n, " + std::to_string(nd.ring_cap) + "u, " + lit(nd.max_delay) + ", " + IN(i, 0, 2) + ", " + P(i, 0) + ", " + P(i, 1) + ");" + cm);
                break;
            case T::AmpSource:
                line(ind, "float " + S(i) + "[32]; ampsrc_process(" + N(i) + ", " + arr_stride + ", " + cold(i) + ", " + S(i) + ");" + cm + "    /* amplitudes: [f32; 32] */");
                break;
            case T::OscBank:
                if (nd.in[0].node >= 0) line(ind, o + "oscbank_process(" + N(i) + ", " + arr_stride + ", " + cold(i) + ", " + P(i, 0) + ", " + SR(i) + ", " + S(nd.in[0].node, nd.in[0].port) + ");" + cm);
                else { line(ind, "const float z" + std::to_string(i) + "[32] = {};    /* `amplitudes` is not connected: [0.0; 32] */");
                       line(ind, o + "oscbank_process(" + N(i) + ", " + arr_stride + ", " + cold(i) + ", " + P(i, 0) + ", " + SR(i) + ", z" + std::to_string(i) + ");" + cm); }
                break;
            case T::SamplePlayer: line(ind, o + "sample_player_process(" + N(i) + ", buf" + std::to_string(i) + ", len" + std::to_string(i) + ");" + cm); break;
            case T::Gain: line(ind, o + IN(i, 0, -1) + " * " + (nd.in[1].node != plan::kNone ? IN(i, 1, 0) : P(i, 0)) + ";" + cm); break;
            case T::AddValue: line(ind, o + IN(i, 0, -1) + " + " + P(i, 0) + ";" + cm); break;
            case T::Mixer: line(ind, o + IN(i, 0, -1) + " + " + IN(i, 1, -1) + ";" + cm); break;
            case T::Sub: line(ind, o + IN(i, 0, -1) + " - " + IN(i, 1, -1) + ";" + cm); break;
            case T::HardClip: line(ind, o + "hardclip_process(" + IN(i, 0, -1) + ");" + cm); break;
            case T::Crossfade:
                line(ind, "const float m" + std::to_string(i) + " = clampf(" + P(i, 0) + ", 0.0f, 1.0f);");
                line(ind, "const float " + S(i, 0) + " = " + IN(i, 0, -1) + " * (1.0f - m" + std::to_string(i) + ");" + cm);
                line(ind, "const float " + S(i, 1) + " = " + IN(i, 0, -1) + " * m" + std::to_string(i) + ";");
                break;
            case T::Expr: {
                using O = plan::ExprOp;
                const std::string x = IN(i, 0, -1), yv = IN(i, 1, -1), pv = P(i, 0), q = lit(nd.cst[1]);
                std::string e;
                switch (nd.op) {
                case O::AddSS: e = x + " + " + yv; break;
                case O::SubSS: e = x + " - " + yv; break;
                case O::MulSS: e = x + " * " + yv; break;
                case O::DivSS: e = x + " / " + yv; break;
                case O::AddSV: e = x + " + " + pv; break;
                case O::SubSV: e = x + " - " + pv; break;
                case O::MulSV: e = x + " * " + pv; break;
                case O::DivSV: e = x + " / " + pv; break;
                case O::SubVS: e = pv + " - " + x; break;
                case O::DivVS: e = pv + " / " + x; break;
                case O::Value: e = pv; break;
                case O::Tanh: e = "om_tanhf_u(" + x + ")"; break;
                case O::Abs: e = "fabsf(" + x + ")"; break;
                case O::Clamp: e = "clampf(" + x + ", " + pv + ", " + q + ")"; break;
                case O::MinV: e = "fminf(" + x + ", " + pv + ")"; break;
                case O::Sqrt: e = "sqrtf(" + x + ")"; break;
                case O::Sin: e = "om_sinf(" + x + ")"; break;
                case O::Cos: e = "om_cosf(" + x + ")"; break;
                case O::Exp: e = "om_expf(" + x + ")"; break;
                default: e = "fmaxf(" + x + ", " + pv + ")"; break;
                }
                line(ind, o + e + ";" + cm);
            } break;
            default:
                line(ind, o + "polyblep_process<" + wave(nd.type) + ">(" + N(i) + ", " + IN(i, 1, 4) + ", " + P(i, 0) + ", " + IN(i, 0, 3) + ", " + P(i, 1) + ", " + P(i, 2) + ", " + SR(i) + ");" + cm);
                break;
            }
        };
        /* the walk of one frame: nodes in topological order (multirate graphs: per phase, emit_frame.rs:114-176) */
        auto emit_walk = [&](int ind) {
            if (!multirate) {
                for (int i : p.order) emit_node(i, ind);
            } else {
                const std::vector<int> ph = phases();
                const std::string ns = std::to_string(NI);
                for (int i : p.order) if (ph[i] == 0) emit_node(i, ind);
                for (int i : p.order) if (ph[i] == 1) {                       /* step 4: Up-edge warm-up */
                    line(ind, "float u" + std::to_string(i) + "[" + ns + "]; " + N(i) + ".up(" + IN(i, 0, -1) + ", u" + std::to_string(i) + ");    /* " + g.nodes[i].name + " */");
                }
                for (int i : p.order) if (ph[i] == 3) line(ind, "float w" + std::to_string(i) + "[" + ns + "];");      /* step 5: Down accumulators */
                line(ind, "_Pragma(\"unroll\") for (int inner = 0; inner < " + ns + "; ++inner) {    /* step 6: N inner ticks */");
                for (int i : p.order) if (ph[i] == 1) line(ind + 1, "const float " + S(i) + " = u" + std::to_string(i) + "[inner];");
                for (int i : p.order) if (ph[i] == 2) emit_node(i, ind + 1);
                for (int i : p.order) if (ph[i] == 3) line(ind + 1, "w" + std::to_string(i) + "[inner] = " + IN(i, 0, -1) + ";");
                line(ind, "}");
                for (int i : p.order) if (ph[i] == 3)                         /* step 7: downsample once per outer tick */
                    line(ind, "const float " + S(i) + " = " + N(i) + ".down(w" + std::to_string(i) + ");    /* " + g.nodes[i].name + " */");
                for (int i : p.order) if (ph[i] == 4) emit_node(i, ind);        /* step 7.5: post-inner outer nodes */
            }
        };
        /* Segmented frame loop: per WARP, L = frames until the first lane's next event or envelope
           stage completion (one REDUX.MIN); L plain frames walk the nodes with no ramp / event / `changed` code and the envelopes
           in their segment form (Attack / Decay: one-pole step, Sustain / Idle: a constant, Release behind a warp-uniform test) —
           the scheme of kernel_fm_seg.cuh, which itself follows the reference's sub-block split at event frames
           (codegen/mod.rs:836-872); then ONE general frame (the full per-frame body below).  Frame 0 of a block and every
           frame of a block with a ramping input are general.  Same operations per frame either way: the same bits. */
        bool any_adsr_node = false, inner_adsr = false;
        for (int i = 0; i < n_nodes; ++i) if (g.nodes[i].type == T::Adsr) { any_adsr_node = true; inner_adsr |= g.nodes[i].rate_log2 > 0; }
        /* (an envelope inside an oversampled section counts its stage down in inner ticks: such graphs keep the per-frame loop) */
        const bool segmented = hoist_adsr && !inner_adsr && env_int("OSCEN_B200_JIT_SEGMENTS", 1) != 0;
        line(1, "for (uint32_t f0 = 0; f0 < frames; f0 += kTileF) {");
        line(2, "const uint32_t fe = min(kTileF, frames - f0);");
        /* Blocks with a ramping input: segments too, as long as no ramped input feeds a quantity this kernel derives outside
           process() and keeps across frames — the envelopes' step data / clamped sustain (adsr_derive_r, adsr_refresh) and the FM
           operators' phase increment (fmop_phase_inc).  Such a segment runs the plain walk with the frame's ramp row loaded
           first (`tick_ramps`, codegen/mod.rs:878-914); every other node reads its parameters inside process() each frame
           anyway.  OSCEN_B200_JIT_RAMP_SEGMENTS=0: ramp blocks on the per-frame path (A/B). */
        uint32_t hoist_mask = 0;
        for (int i = 0; i < n_nodes; ++i) {
            const plan::Node &nd = g.nodes[i];
            const int slots = nd.type == T::Adsr ? 4 : (nd.type == T::FmOperator ? 2 : 0);
            for (int k = 0; k < slots; ++k) if (nd.pid[k] != plan::kNoValue) hoist_mask |= 1u << nd.pid[k];
        }
        const bool ramp_segments = segmented && ni > 0 && env_int("OSCEN_B200_JIT_RAMP_SEGMENTS", 1) != 0;
        if (segmented) {
            line(2, "uint32_t ff = 0;");
            line(2, "while (ff < fe) {");
            line(3, "const uint32_t f = f0 + ff;");
            line(3, "{");
            if (ramp_segments) {
                line(4, fmt("const bool ramp_seg = (ramp_mask & 0x%xu) == 0u;    /* no ramp on an input behind hoisted state */", hoist_mask));
                line(4, "uint32_t np = ((ramp_mask != 0u && !ramp_seg) || f == 0u) ? 0u : (active ? ec.next_frame - f : 0x7FFFFFFFu);");
            } else
            line(4, "uint32_t np = (ramp_mask != 0u || f == 0u) ? 0u : (active ? ec.next_frame - f : 0x7FFFFFFFu);");
            for (int i = 0; i < n_nodes; ++i) if (g.nodes[i].type == T::Adsr)
                line(4, "np = oscen::umin(np, adsr_plain_frames(" + N(i) + "));");
            line(4, "const uint32_t L = oscen::umin(__reduce_min_sync(0xffffffffu, np), fe - ff);");
            line(4, "if (L > 0u) {");
            std::string rl_any;
            for (int i = 0; i < n_nodes; ++i) if (g.nodes[i].type == T::Adsr) {
                const std::string k = std::to_string(i), h = "h" + k;
                line(5, "const bool ad" + k + " = (" + N(i) + ".stage == ST_ATTACK) | (" + N(i) + ".stage == ST_DECAY), rl" + k + " = " + N(i) + ".stage == ST_RELEASE;");
                line(5, "const float tg" + k + " = " + N(i) + ".stage == ST_ATTACK ? 1.0f : " + h + ".sus, kk" + k + " = " + N(i) + ".stage == ST_ATTACK ? " + D(i) + ".a_k : " + D(i) + ".d_k;");
                rl_any += (rl_any.empty() ? "" : " | ") + std::string("rl") + k;
            }
            if (any_adsr_node) line(5, "const bool any_rl = __any_sync(0xffffffffu, " + rl_any + ");");
            for (int ramp_form = 0; ramp_form < (ramp_segments ? 2 : 1); ++ramp_form) {
                /* two copies of the plain loop: without ramps nothing in it writes a parameter (so the compiler hoists every
                   parameter-derived computation out of it); with ramps the frame's row is loaded first */
                if (ramp_segments) line(5, ramp_form == 0 ? "if (ramp_mask == 0u) {" : "} else {");
                line(5, "for (uint32_t i_ = 0; i_ < L; ++i_) {");
                line(6, "const uint32_t f = f0 + ff + i_;");
                line(6, "float y = 0.0f;");
                line(6, "if (active) {");
                if (ramp_form == 1) {
                    line(7, "const float *row = a.ptab + (size_t)f * kMaxInputs;");
                    for (uint32_t i = 0; i < ni; ++i) if (!((uniform_mask >> i) & 1u) && !((hoist_mask >> i) & 1u))
                        line(7, fmt("if ((ramp_mask >> %u) & 1u) P[%u] = row[%u];", i, i, i));
                }
                if (has_in) line(7, "const float x_in = lcg_on ? lcg_next(lcg) : ((a.stream_in != nullptr) ? __ldcs(a.stream_in + (size_t)f * n + v) : 0.0f);");
                plain_mode = true;
                emit_walk(7);
                plain_mode = false;
                line(7, "y = " + S(g.output.node, g.output.port) + ";    /* -> " + g.output_name + " */");
                line(7, "if (a.voice_out != nullptr) a.voice_out[(size_t)f * n + v] = y;");
                line(6, "}");
                line(6, "tile[(ff + i_) * kTileW + col] = y;");
                line(5, "}");
            }
            if (ramp_segments) line(5, "}");
            for (int i = 0; i < n_nodes; ++i) if (g.nodes[i].type == T::Adsr)
                line(5, "if (!any_rl && ad" + std::to_string(i) + ") " + N(i) + ".rem -= L;    /* the Release form counts down per frame itself */");
            line(5, "ff += L;");
            line(5, "continue;");
            line(4, "}");
            line(3, "}");
        } else {
            line(2, "for (uint32_t ff = 0; ff < fe; ++ff) {");
            line(3, "const uint32_t f = f0 + ff;");
        }
        line(3, "float y = 0.0f;");
        line(3, "if (active) {");
        /* ---- tick_ramps: ramped inputs are read through `.current` (codegen/emit_node.rs:317-321) */
        if (ni) {
            line(4, "if (ramp_mask) {");
            line(5, "const float *row = a.ptab + (size_t)f * kMaxInputs;");
            for (uint32_t i = 0; i < ni; ++i) if (!((uniform_mask >> i) & 1u)) { line(5, fmt("if ((ramp_mask >> %u) & 1u) P[%u] = row[%u];", i, i, i)); }
            line(4, "}");
        }
        /* ---- this frame's events: value edges are assigned before process_event_inputs() */
        line(4, "const bool due = ec.due(f);");
        line(4, "uint32_t ev_end = ec.i;");
        line(4, "if (due) {");
        line(5, "ev_end = ec.end_of_frame(f);");
        line(5, "for (uint32_t q = ec.i; q < ev_end; ++q) {");
        line(6, "const Event e = ec.ev[q];");
        line(6, "if (e.input_id == kEventGate) continue;");
        /* select form: `if (id == k) P[k] = x` chains are turned into ONE store with a run-time index by the compiler, which moves P
           from registers to local memory for the whole kernel */
        for (uint32_t i = 0; i < ni; ++i) if (!((uniform_mask >> i) & 1u)) { line(6, fmt("P[%u] = (e.input_id == %uu) ? e.scalar : P[%u];", i, i, i)); }
        line(6, fmt("if (e.input_id < %uu) a.pv[(size_t)e.input_id * n + v] = e.scalar;", ni));
        line(5, "}");
        line(4, "}");
        line(4, "const bool changed = (ramp_mask != 0u) || due;    /* a value input may have moved this frame */");
        line(4, "(void)changed;");
        for (int i = 0; i < n_nodes; ++i) {
            const plan::Node &nd = g.nodes[i];
            if (nd.type == T::Adsr && !(P_is_const(i, 0) && P_is_const(i, 1) && P_is_const(i, 3)))
                line(4, "if (changed && adsr_derived_stale(" + D(i) + ", " + P(i, 0) + ", " + P(i, 1) + ", " + P(i, 3) + ")) adsr_derive_r(" + D(i) + ", " + P(i, 0) + ", " + P(i, 1) + ", " + P(i, 3) + ", " + SR(i) + ");");
            if (nd.type == T::FmOperator && !(P_is_const(i, 0) && P_is_const(i, 1)))
                line(4, "if (changed) " + N(i) + ".inc = fmop_phase_inc(" + P(i, 0) + ", " + P(i, 1) + ", " + SR(i) + ");");    /* an inner-rate node runs at sr * N */
        }
        bool any_gate = false;
        for (int i = 0; i < n_nodes; ++i) any_gate |= (g.nodes[i].type == T::Adsr || g.nodes[i].type == T::AmpSource || g.nodes[i].type == T::OscBank) && g.nodes[i].gate;
        line(4, "if (due) {");
        if (any_gate) {
            line(5, "for (uint32_t q = ec.i; q < ev_end; ++q) {");
            line(6, "if (ec.ev[q].input_id != kEventGate) continue;");
            line(6, "const float gv = ec.ev[q].scalar;");
            for (int i = 0; i < n_nodes; ++i) if (g.nodes[i].type == T::Adsr && g.nodes[i].gate)
                line(6, "adsr_on_gate(" + N(i) + ", gv, " + P(i, 0) + ", " + P(i, 3) + ", " + D(i) + ");    /* " + g.event_input + " -> " + g.nodes[i].name + ".gate */");
            for (int i = 0; i < n_nodes; ++i) if (g.nodes[i].gate) {
                if (g.nodes[i].type == T::AmpSource)
                    line(6, "ampsrc_on_gate(" + N(i) + ", " + arr_stride + ", " + cold(i) + ", gv, " + P(i, 1) + ", " + P(i, 2) + ", " + P(i, 3) + ", " + P(i, 4) + ", " + P(i, 5) + ", " + P(i, 6) + ");    /* " + g.event_input + " -> " + g.nodes[i].name + ".gate */");
                if (g.nodes[i].type == T::OscBank)
                    line(6, "oscbank_on_gate(" + N(i) + ", " + arr_stride + ", gv);    /* " + g.event_input + " -> " + g.nodes[i].name + ".gate */");
            }
            line(5, "}");
        }
        line(5, "ec.advance(ev_end);");
        line(4, "}");
        {
            bool any_adsr = false;
            for (int i = 0; i < n_nodes; ++i) any_adsr |= g.nodes[i].type == T::Adsr;
            if (any_adsr && hoist_adsr) {
                line(4, "if (changed) {");
                for (int i = 0; i < n_nodes; ++i) if (g.nodes[i].type == T::Adsr)
                    line(5, "adsr_refresh(" + N(i) + ", h" + std::to_string(i) + ", " + P(i, 2) + ", " + D(i) + ");");
                line(4, "}");
            }
        }
        if (has_in) line(4, "const float x_in = lcg_on ? lcg_next(lcg) : ((a.stream_in != nullptr) ? __ldcs(a.stream_in + (size_t)f * n + v) : 0.0f);");
        emit_walk(4);
        line(4, "y = " + S(g.output.node, g.output.port) + ";    /* -> " + g.output_name + " */");
        line(4, "if (a.voice_out != nullptr) a.voice_out[(size_t)f * n + v] = y;");
        line(3, "}");
        line(3, "tile[ff * kTileW + col] = y;");
        if (segmented) line(3, "ff += 1u;");
        line(2, "}");
        if (g.post_tremolo) {
            auto pid = [&](int k) { return g.post_pid[k] == plan::kNoValue ? std::string("kNoPostInput") : std::to_string(g.post_pid[k]) + "u"; };
            line(2, "tile_flush_tremolo(a, tile, isum2, f0, fe, trem_phase, " + pid(0) + ", " + lit(g.post_cst[0]) + ", " + pid(1) + ", " + lit(g.post_cst[1]) + ");    /* rate, depth */");
        } else {
            line(2, "tile_flush_mono(a, tile, isum, f0, fe);");
        }
        line(1, "}");
        /* ---- registers -> HBM */
        if (g.post_tremolo) line(1, "if (threadIdx.x < kCta / a.vpi && my_inst < a.n_inst) a.inst_state[my_inst] = trem_phase;");
        line(1, "if (active) {");
        for (int i = 0; i < n_nodes; ++i) {
            const plan::Node &nd = g.nodes[i];
            const std::string base = std::to_string(nd.state_base) + "u";
            switch (nd.type) {
            case T::Adsr: line(2, "adsr_store(" + N(i) + ", a.state, " + base + ", n, v);"); break;
            case T::AmpSource: case T::OscBank:
                if (arr_smem) line(2, "_Pragma(\"unroll 8\") for (uint32_t w_ = 0; w_ < " + std::to_string(array_state_words(nd)) + "u; ++w_) a.state[(size_t)(" + base + " + w_) * n + v] = " + N(i) + "[(size_t)w_ * kCta];");
                break;
            case T::FmOperator:
                line(2, "a.state[(size_t)(" + base + " + 0u) * n + v] = " + N(i) + ".phase; a.state[(size_t)(" + base + " + 1u) * n + v] = " + N(i) + ".prev;");
                break;
            case T::Tpt: line(2, "tpt_store(" + N(i) + ", a.state, " + base + ", n, v);"); break;
            case T::Lp18: line(2, "lp18_store(" + N(i) + ", a.state, " + base + ", n, v);"); break;
            case T::IirLowpass:
                line(2, "{ float *st_ = a.state + (size_t)" + base + " * n + v; st_[0] = " + N(i) + ".b0; st_[n] = " + N(i) + ".b1; st_[2 * (size_t)n] = " + N(i) + ".b2; st_[3 * (size_t)n] = " +
                        N(i) + ".a1; st_[4 * (size_t)n] = " + N(i) + ".a2; st_[5 * (size_t)n] = " + N(i) + ".v1; st_[6 * (size_t)n] = " + N(i) + ".v2; st_[7 * (size_t)n] = __uint_as_float(" + N(i) + ".counter); }");
                break;
            case T::OscSine: case T::OscSquare: case T::OscSaw: line(2, "a.state[(size_t)" + base + " * n + v] = " + N(i) + ";"); break;
            case T::Delay:
                line(2, "a.state[(size_t)(" + base + " + 0u) * n + v] = __uint_as_float(" + N(i) + ".write_pos); a.state[(size_t)(" + base + " + 1u) * n + v] = __uint_as_float(" + N(i) +
                        ".counter); a.state[(size_t)(" + base + " + 2u) * n + v] = " + N(i) + ".output;");
                break;
            case T::SamplePlayer: line(2, "a.state[(size_t)" + base + " * n + v] = __uint_as_float(" + N(i) + ");"); break;
            case T::Up: case T::Down: line(2, N(i) + ".store(a.state, " + base + ", n, v);"); break;
            default:
                if (is_polyblep(nd.type)) line(2, "a.state[(size_t)" + base + " * n + v] = " + N(i) + ";");
                break;
            }
        }
        if (has_in) {
            const std::string f = std::to_string(g.n_state) + "u";
            line(2, "a.state[(size_t)(" + f + " + 0u) * n + v] = __uint_as_float((uint32_t)(lcg & 0xffffffffull));");
            line(2, "a.state[(size_t)(" + f + " + 1u) * n + v] = __uint_as_float((uint32_t)(lcg >> 32));");
        }
        line(1, "}");
        line(0, "}");
        return s;
    }
};

/* dynamic shared memory of the generated kernel: the block-resident state of its array-stream nodes (the FIR histories of
   [sinc] cross-rate edges moved to registers in round 2) */
inline size_t dyn_smem_bytes(const plan::Graph &g) {
    size_t floats = 0;
    for (const plan::Node &nd : g.nodes) floats += Emitter::edge_smem_floats(nd);
    floats += Emitter::array_smem_floats(g);  /* the block-resident state of AmplitudeSource / OscillatorBank nodes */
    return floats * 64u * sizeof(float);      /* x kCta threads */
}

inline std::string fused_kernel_source(const plan::Graph &g, const plan::Plan &p, const char *kernel_name = "oscen_b200_generated_kernel",
                                       uint32_t uniform_mask = 0) {
    Emitter e(g, p);
    e.uniform_mask = uniform_mask;
    return e.run(kernel_name);
}

}  // namespace codegen
}  // namespace oscen
