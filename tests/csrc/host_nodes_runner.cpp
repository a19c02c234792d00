/*
 * tests/csrc/host_nodes_runner.cpp — TEST INFRASTRUCTURE.  Runs one block of the NODE PLAN (OSCEN_B200_PLAN_NODES) on
 * the host: the node kernels of oscen_b200/csrc/kernel_nodes.cuh, compiled by g++ against
 * tests/csrc/host_cuda/cuda_runtime.h, launched in the order of the plan that dsl.hpp + plan.hpp (host-only product
 * headers, used as they are) compile from the graph text.  The launch sequence restates engine.cu's launch_plan():
 * one launch per node type per ready wave (blockIdx.y = node), stateless nodes one stream kernel each (float4
 * grid-stride form when the block carries no value events and n_voices % 4 == 0), then the instance fold + mix partials
 * and the value-event commit.  Streams live in host memory as [frame][voice] rows, like the engine's HBM buffers.
 */
#include "dsl.hpp"
#include "kernel_nodes.cuh"

namespace {
using namespace oscen;

struct HostPlan {
    plan::Graph g;
    plan::Plan p;
    std::vector<std::vector<float>> streams;
    uint32_t n_voices = 0;
};

template <int OP>
void launch_ew(const NodeArgs &n, uint32_t grid, bool vec) {
    if (vec) host_launch(3, 1, 64, [&]() { ew_node_vec_kernel<OP>(n); });            /* any grid: a grid-stride loop */
    else { NodeArgs4 a4{}; a4.n[0] = n; host_launch(grid, 1, kCta, [&]() { ew_node_kernel<OP>(a4); }); }
}
}  // namespace


/* stream_in_lcg: the graph's stream input is generated on the device from the per-voice LCG state (BlockCtl.lcg_on) */
static uint32_t g_lcg_on = 0, g_lcg_field = 0;
extern "C" void oscen_host_set_lcg(uint32_t on, uint32_t field) { g_lcg_on = on; g_lcg_field = field; }

extern "C" void *oscen_host_plan_create(const char *text, float sample_rate, char *err, size_t err_cap) {
    HostPlan *h = new HostPlan();
    std::string e;
    if (!dsl::compile(text, h->g, e) || !plan::compile(h->g, h->p, e)) {
        if (err && err_cap) { strncpy(err, e.c_str(), err_cap - 1); err[err_cap - 1] = 0; }
        delete h;
        return nullptr;
    }
    plan::assign_state_layout(h->g, sample_rate);
    return h;
}
extern "C" void oscen_host_plan_free(void *h) { delete (HostPlan *)h; }

extern "C" int oscen_host_plan_run_block(void *handle, uint32_t n, uint32_t vpi, float sample_rate, float *state, float *pv,
                                         const Event *events, uint32_t n_events, const float *ptab, uint32_t ramp_mask,
                                         uint32_t frames, const float *stream_in, int has_value_events, float *voice_out,
                                         float *inst_out, float *mix_partial, const float *const *asset_data,
                                         const uint32_t *asset_frames, const uint32_t *asset_reset) {
    HostPlan *h = (HostPlan *)handle;
    const plan::Graph &g = h->g;
    const plan::Plan &p = h->p;
    if (!h || n == 0 || vpi == 0 || kCta % vpi != 0 || n % vpi != 0 || frames > kMaxBlock) return -1;
    if (g.inner_log2 > 0) return -2;                                               /* multirate graphs have no node plan */
    for (const plan::Node &nd : g.nodes) {
        if (!plan::node_plan_has_kernel(nd.type)) return -2;
        for (int k = 0; k < plan::kMaxStreamIn; ++k) if (nd.in[k].node >= 0 && nd.fb[k]) return -2;   /* feedback: per-sample walk only */
    }
    if (h->n_voices != n) {
        h->streams.assign((size_t)p.n_streams + 1, std::vector<float>((size_t)kMaxBlock * n, 0.0f));
        h->n_voices = n;
    }
    BlockCtl ctl{};
    ctl.frames = frames; ctl.n_events = n_events; ctl.ramp_mask = ramp_mask; ctl.lcg_on = g_lcg_on;
    AssetTable assets{};
    for (uint32_t k = 0; k < kMaxAssets && asset_data; ++k) { assets.data[k] = asset_data[k]; assets.frames[k] = asset_frames[k]; assets.reset[k] = asset_reset[k]; }
    const uint32_t grid = (n + kCta - 1) / kCta;
    const bool vec = !has_value_events && (n % 4u == 0u);
    const bool plain = n_events == 0 && ramp_mask == 0;
    auto S = [&](int i) { return h->streams[(size_t)i].data(); };
    const float *gin = stream_in;
    if (g_lcg_on && !g.stream_input.empty()) {                 /* launch_plan(): the LCG source node fills a stream of its own */
        float *dst = S(p.n_streams);
        host_launch(grid, 1, kCta, [&]() { lcg_source_kernel(state, dst, &ctl, n, g_lcg_field); });
        gin = dst;
    }
    auto resolve = [&](const plan::StreamRef &r) -> float * {
        if (r.node == plan::kNone) return nullptr;
        if (r.node == plan::kGraphInput) return const_cast<float *>(gin);
        return S(p.first_stream[r.node] + r.port);
    };
    auto args_of = [&](int idx) {
        const plan::Node &nd = g.nodes[idx];
        NodeArgs a{};
        a.n_voices = n; a.state_base = nd.state_base; a.ctl = &ctl; a.events = events;
        a.ptab = ptab; a.pv = pv; a.state = state; a.sr = sample_rate;
        a.in0 = resolve(nd.in[0]); a.in1 = resolve(nd.in[1]); a.in2 = resolve(nd.in[2]);
        a.out0 = S(p.first_stream[idx]);
        a.out1 = plan::n_outputs(nd.type) > 1 ? S(p.first_stream[idx] + 1) : nullptr;
        for (int k = 0; k < plan::kMaxParams; ++k) { a.pid[k] = nd.pid[k]; a.cst[k] = nd.cst[k]; }
        a.gate = nd.gate ? 1u : 0u; a.q = nd.cst[1];
        if (nd.type == plan::NodeType::Delay) { a.gate = nd.ring_cap; a.q = nd.max_delay; }
        if (nd.type == plan::NodeType::SamplePlayer) {
            a.in1 = reinterpret_cast<const float *>(&assets); a.gate = nd.asset >= 0 ? (uint32_t)nd.asset : 0xFFFFFFFFu;
        }
        return a;
    };
    using T = plan::NodeType;
    for (const plan::Launch &l : p.launches) {
        NodeArgs4 a4{};
        const uint32_t cnt = (uint32_t)l.nodes.size();
        for (uint32_t k = 0; k < cnt; ++k) a4.n[k] = args_of(l.nodes[k]);
#define NODE_LAUNCH(KERNEL) host_launch(grid, cnt, kCta, [&]() { KERNEL(a4); })
        switch (l.type) {
        /* launch_plan(): a block without events and ramps runs the plain variants */
        case T::Adsr: if (plain) NODE_LAUNCH(adsr_node_plain_kernel); else NODE_LAUNCH(adsr_node_kernel); break;
        case T::FmOperator: if (plain) NODE_LAUNCH(fmop_node_plain_kernel); else NODE_LAUNCH(fmop_node_kernel); break;
        case T::Tpt: {
            bool zero_fmod = true;
            for (uint32_t k = 0; k < cnt; ++k) zero_fmod = zero_fmod && a4.n[k].in2 == nullptr && a4.n[k].pid[2] == kNoInput && a4.n[k].cst[2] == 0.0f;
            if (plain && zero_fmod) NODE_LAUNCH(tpt_node_plain_kernel<true>);
            else if (plain) NODE_LAUNCH(tpt_node_plain_kernel<false>);
            else NODE_LAUNCH(tpt_node_kernel);
        } break;
        case T::PolyBlepSine: NODE_LAUNCH(polyblep_node_kernel<WAVE_SINE>); break;
        case T::PolyBlepSaw: NODE_LAUNCH(polyblep_node_kernel<WAVE_SAW>); break;
        case T::PolyBlepSquare: NODE_LAUNCH(polyblep_node_kernel<WAVE_SQUARE>); break;
        case T::PolyBlepTriangle: NODE_LAUNCH(polyblep_node_kernel<WAVE_TRIANGLE>); break;
        case T::Lp18: NODE_LAUNCH(lp18_node_kernel); break;
        case T::OscSine: NODE_LAUNCH(osc_node_kernel<OSC_SINE>); break;
        case T::OscSquare: NODE_LAUNCH(osc_node_kernel<OSC_SQUARE>); break;
        case T::OscSaw: NODE_LAUNCH(osc_node_kernel<OSC_SAW>); break;
        case T::IirLowpass: NODE_LAUNCH(iirlp_node_kernel); break;
        case T::Delay: NODE_LAUNCH(delay_node_kernel); break;
        case T::SamplePlayer: NODE_LAUNCH(sample_player_node_kernel); break;
#undef NODE_LAUNCH
        default:
            for (uint32_t k = 0; k < cnt; ++k) {
                const NodeArgs &a = a4.n[k];
                switch (l.type) {
                case T::Gain: if (a.in1 != nullptr) launch_ew<EW_MUL>(a, grid, vec); else launch_ew<EW_GAIN>(a, grid, vec); break;
                case T::AddValue: launch_ew<EW_ADD_VALUE>(a, grid, vec); break;
                case T::Mixer: launch_ew<EW_MIXER>(a, grid, vec); break;
                case T::Sub: launch_ew<EW_SUB>(a, grid, vec); break;
                case T::Crossfade: launch_ew<EW_CROSSFADE>(a, grid, vec); break;
                case T::Expr:
                    switch (g.nodes[l.nodes[k]].op) {
                    case plan::ExprOp::AddSS: launch_ew<EW_MIXER>(a, grid, vec); break;
                    case plan::ExprOp::SubSS: launch_ew<EW_SUB>(a, grid, vec); break;
                    case plan::ExprOp::MulSS: launch_ew<EW_MUL>(a, grid, vec); break;
                    case plan::ExprOp::DivSS: launch_ew<EW_DIV>(a, grid, vec); break;
                    case plan::ExprOp::AddSV: launch_ew<EW_ADD_VALUE>(a, grid, vec); break;
                    case plan::ExprOp::SubSV: launch_ew<EW_SUBV>(a, grid, vec); break;
                    case plan::ExprOp::MulSV: launch_ew<EW_GAIN>(a, grid, vec); break;
                    case plan::ExprOp::DivSV: launch_ew<EW_DIVV>(a, grid, vec); break;
                    case plan::ExprOp::SubVS: launch_ew<EW_RSUBV>(a, grid, vec); break;
                    case plan::ExprOp::DivVS: launch_ew<EW_RDIVV>(a, grid, vec); break;
                    case plan::ExprOp::Value: launch_ew<EW_VALUE>(a, grid, vec); break;
                    case plan::ExprOp::Tanh: launch_ew<EW_TANH>(a, grid, vec); break;
                    case plan::ExprOp::Abs: launch_ew<EW_ABS>(a, grid, vec); break;
                    case plan::ExprOp::Clamp: launch_ew<EW_CLAMP>(a, grid, vec); break;
                    case plan::ExprOp::MinV: launch_ew<EW_MINV>(a, grid, vec); break;
                    case plan::ExprOp::Sqrt: launch_ew<EW_SQRT>(a, grid, vec); break;
                    case plan::ExprOp::Sin: launch_ew<EW_SIN>(a, grid, vec); break;
                    case plan::ExprOp::Cos: launch_ew<EW_COS>(a, grid, vec); break;
                    case plan::ExprOp::Exp: launch_ew<EW_EXP>(a, grid, vec); break;
                    default: launch_ew<EW_MAXV>(a, grid, vec); break;
                    }
                    break;
                default: launch_ew<EW_HARDCLIP>(a, grid, vec); break;
                }
            }
            break;
        }
    }
    const float *final_stream = S(p.first_stream[g.output.node] + g.output.port);
    {
        KernelArgs a;
        memset(&a, 0, sizeof a);
        a.n_voices = n; a.vpi = vpi; a.n_inst = n / vpi; a.channels = 1; a.ctl = &ctl;
        a.inst_out = inst_out; a.mix_partial = mix_partial; a.u.sr = sample_rate;
        host_launch(grid, 1, kCta, [&]() { instance_mix_node_kernel(a, final_stream); });
    }
    if (n_events) host_launch((n_events + 63u) / 64u, 1, 64, [&]() { commit_value_events_kernel(pv, events, &ctl, n, (uint32_t)g.inputs.size()); });
    memcpy(voice_out, final_stream, (size_t)frames * n * sizeof(float));
    return 0;
}
