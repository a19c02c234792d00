/*
 * engine.cu — the B200-native voice-batched engine behind include/oscen_b200.h.
 *
 * One engine = one GPU = one shard of voices.  Host side: value inputs / ramps / events / MIDI
 * front bookkeeping that mirrors what the generated graph struct does between blocks
 * (oscen-graph-compiler/src/codegen/mod.rs:755-976).  Device side: the fused per-graph kernels
 * (kernel_fm.cuh, kernel_others.cuh) or the node-granular CUDA-graph plan (kernel_nodes.cuh).
 * There is NO CPU rendering path: without a CUDA device create() fails with E_NO_DEVICE.
 */
#include <cuda_runtime.h>
#include <cuda.h>              /* CUtensorMap and the cuTensorMapEncodeTiled prototype only: the entry point is fetched at run time */
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/oscen_b200.h"
#include "engine_types.h"
#include "kernel_fm.cuh"
#include "kernel_fm_seg.cuh"
#include "fm_select.hpp"
#include "kernel_nodes.cuh"
#include "kernel_others.cuh"
#include "kernel_resample.cuh"
#include "graphs.hpp"
#include "dsl.hpp"
#include "codegen.hpp"
#include "jit.hpp"
#include "midi_front.hpp"
#include "midi_device.cuh"

using namespace oscen;

static_assert(sizeof(Event) == sizeof(oscen_b200_event), "event layout");
static_assert(kMaxInputs == OSCEN_B200_MAX_INPUTS, "input count");

namespace {

thread_local std::string g_create_error;
constexpr uint32_t kTimingRing = 4096;

const char *const kJitKernelName = "oscen_b200_generated_kernel";
const uint32_t kNInputs[7] = {28, 3, 2, 2, 9, 4, 8};       /* clip graphs: no value inputs */
bool is_clip(uint32_t kind) { return kind >= OSCEN_B200_GRAPH_CLIP_REF && kind < OSCEN_B200_GRAPH_CUSTOM; }
int clip_policy(uint32_t kind) { return kind == OSCEN_B200_GRAPH_CLIP_REF ? POL_LATCH : (int)(kind - OSCEN_B200_GRAPH_CLIP_OS_BASE) / 3; }
int clip_log2f(uint32_t kind) { return kind == OSCEN_B200_GRAPH_CLIP_REF ? 0 : 1 + (int)(kind - OSCEN_B200_GRAPH_CLIP_OS_BASE) % 3; }
bool has_stream_in(uint32_t kind) { return kind == OSCEN_B200_GRAPH_TWIN_PEAKS || is_clip(kind); }
uint32_t lcg_field(uint32_t kind) { return kind == OSCEN_B200_GRAPH_TWIN_PEAKS ? (uint32_t)TW_LCG : (uint32_t)CL_LCG; }
uint32_t n_state_of(uint32_t kind) {
    static const uint32_t k[7] = {FM_NSTATE, TW_NSTATE, SAT_NSTATE, SAT_NSTATE, EP_NSTATE, SY_NSTATE, SY_NSTATE};
    return is_clip(kind) ? clip_n_state(clip_policy(kind), clip_log2f(kind)) : k[kind];
}
/* defaults of the reference graphs' value inputs (fm_voice.rs:10-48; nih-twin-peaks lib.rs:21-23;
   oversampled-saturator main.rs:71; electric-piano main.rs:39-46; README.md:28-31) */
const float kDefaults[7][kMaxInputs] = {
    {440.0f, 3.0f, 0.5f, 0.0f, 0.01f, 0.1f, 0.7f, 0.3f, 2.0f, 0.5f, 0.0f, 0.01f, 0.1f, 0.7f, 0.3f,
     1.0f, 0.01f, 0.2f, 0.8f, 0.5f, 0.0f, 2000.0f, 0.707f, 0.01f, 0.2f, 0.5f, 0.3f, 0.0f},
    {1000.0f, 1900.0f, 0.54f},
    {2000.0f, 0.6f},
    {2000.0f, 0.6f},
    {440.0f, 30.0f, 50.0f, 90.0f, 70.0f, 50.0f, 40.0f, 0.3f, 5.0f},
    {5.0f, 0.2f, 440.0f, 1200.0f},
    {5.0f, 0.2f, 440.0f, 1200.0f, 0.01f, 0.2f, 0.7f, 0.3f},
};

/* ValueRampState, oscen-lib/src/graph/types.rs:300-373 */
struct Ramp {
    float current = 0, target = 0, increment = 0;
    uint32_t remaining = 0;
    void set_immediate(float v) { current = v; target = v; increment = 0; remaining = 0; }
    void set_with_ramp(float t, uint32_t frames) {
        if (frames == 0) { set_immediate(t); return; }
        target = t; increment = (t - current) / (float)frames; remaining = frames;
    }
    void tick() {
        if (remaining > 0) {
            remaining -= 1;
            if (remaining == 0) { current = target; increment = 0; }
            else current += increment;
        }
    }
};

/* NCCL through dlopen: single-GPU use needs no NCCL at all */
struct Id128 { char b[128]; };
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, /* ncclUniqueId by value: 128 bytes */ Id128, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
bool nccl_load(std::string &err) {
    if (g_nccl.lib) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) { g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
    if (!g_nccl.lib) { err = std::string("dlopen libnccl.so.2 failed: ") + dlerror(); return false; }
    g_nccl.GetUniqueId = (int (*)(void *))dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void **, int, Id128, int))dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(g_nccl.lib, "ncclAllReduce");
    g_nccl.CommDestroy = (int (*)(void *))dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce) { err = "NCCL symbols missing"; return false; }
    return true;
}

}  // namespace

/* a compiled `graph!` description (dsl.hpp): host-only, shareable between engines */
struct oscen_b200_graph {
    plan::Graph g;
    plan::Plan p;
};

struct oscen_b200_engine {
    oscen_b200_config cfg{};
    uint32_t n_inputs = 0, n_state = 0, n_inst = 0, channels = 1, n_cta = 0;
    uint32_t n_cta_mix = 0;          /* CTAs that wrote mix partials in the last block */
    cudaStream_t stream = nullptr;
    std::string err;

    /* value inputs */
    float uni[kMaxInputs] = {};
    bool uniform[kMaxInputs] = {};
    bool ramped[kMaxInputs] = {};
    Ramp ramps[kMaxInputs];
    float *d_pv = nullptr;
    float *d_ptab = nullptr;
    float *d_utab = nullptr;             /* FMVoice, segmented kernel in ramp blocks: one FmRow per frame (kernel_fm_seg.cuh) */
    FmBoundState fm_bounds;              /* fm_select.hpp: levels the previous block rendered with */
    int fm_variant = FM_GENERAL;
    /* pinned staging for per-block uploads (events, ramp tables): a ring, so that filling the next block's buffer never
       waits for the previous block's copy — each slot carries an event recorded after its copy was enqueued */
    static constexpr int kStage = 8;
    void *stage_h[kStage] = {};
    size_t stage_cap[kStage] = {};
    cudaEvent_t stage_ev[kStage] = {};
    bool stage_busy[kStage] = {};
    int stage_next = 0;

    /* state */
    float *d_state = nullptr, *d_inst_state = nullptr;

    /* events */
    std::vector<Event> pending;
    std::vector<uint32_t> seqs;
    std::vector<MidiFront> midi;         /* host front, allocated on the first push_midi */
    /* device-side MIDI front (midi_device.cuh): per-voice allocator / handler state and scratch */
    int midi_mode = 0;                   /* 0 unset, 1 host front, 2 device front (exclusive per engine) */
    std::vector<MidiMsg> midi_batch;     /* this block's messages, push order */
    MidiMsg *d_midi_msgs = nullptr;
    uint32_t *d_midi_keys_in = nullptr, *d_midi_keys = nullptr, *d_midi_idx_in = nullptr, *d_midi_idx = nullptr;
    size_t midi_cap = 0;
    void *d_cub_tmp = nullptr; size_t cub_tmp_bytes = 0;
    uint32_t *d_alloc_word = nullptr, *d_alloc_age = nullptr, *d_alloc_cur_age = nullptr;
    int *d_h_note = nullptr; float *d_h_freq = nullptr, *d_note_freq = nullptr;
    uint32_t *d_ev_count = nullptr, *d_ev_offset = nullptr, *d_midi_dropped = nullptr;
    uint32_t seq = 0;
    Event *d_events = nullptr;
    size_t events_cap = 0;
    BlockCtl *d_ctl = nullptr, *h_ctl = nullptr;
    BlockCtl ctl_on_device{};

    /* stream input */
    float *d_stream_in = nullptr;
    size_t stream_in_cap = 0;
    const float *stream_in_ptr = nullptr;
    uint32_t stream_in_frames = 0;
    bool lcg_on = false;
    unsigned long long lcg_seed = 0;
    uint32_t voice_offset = 0;

    /* outputs */
    float *d_voice_out = nullptr, *d_inst_out = nullptr, *d_mix_partial = nullptr, *d_mix = nullptr;
    float *h_mix = nullptr;
    uint32_t last_frames = 0;

    uint32_t midi_freq_input = 0;        /* value input fed by MidiVoiceHandler.frequency */
    bool midi_ok = true;                 /* the graph has a MIDI front (DSL graphs: only the polyphonic form with the MIDI chain) */
    bool has_stream_in = false;          /* the graph declares a per-voice stream input */
    uint32_t lcg_field = 0;              /* first of the two state words of the device-side input LCG */
    jit::Kernel jit;                     /* DSL-compiled graph, PLAN_FUSED: the kernel codegen.hpp generated, compiled by NVRTC */
    jit::Kernel jit_uni;                 /* oscen_b200_specialize: the variant that reads the inputs of jit_uni_mask from the constant bank */
    uint32_t jit_uni_mask = 0;
    std::string jit_uni_name;
    /* `external NAME: AudioAsset;` slots of a DSL graph (engine_types.h AssetTable) */
    uint32_t n_assets = 0;
    AssetTable h_assets{};
    AssetTable *d_assets = nullptr;
    float *d_asset_buf[kMaxAssets] = {};
    bool asset_reset_pending = false;    /* a slot was loaded since the last block: its players restart, then the flags are cleared */

    /* node plan */
    plan::Graph graph;
    plan::Plan plan;
    float *final_stream = nullptr;
    const float *captured_in = nullptr;      /* graph-input pointer / LCG mode baked into the captured CUDA graphs */
    bool captured_lcg = false;
    std::vector<float *> d_streams;
    CUtensorMap *d_tmaps = nullptr;          /* one TMA descriptor per stream buffer ([kMaxBlock][n_voices], box [kTmaF][kCta]); null: not built */
    /* captured launch sequences, one per (block carries value events?, block carries any event?, which inputs ramp): the
       kernels a wave runs depend on all three (plain / TMA kernels need event-free blocks and none of the wave's inputs
       ramping).  A handful of keys occur in practice (ramps start and stop rarely); the oldest entry goes when 12 are held. */
    struct PlanVariant { uint64_t key = 0; cudaGraphExec_t exec = nullptr; uint64_t kernels = 0; std::string names; uint64_t last_use = 0; };
    std::vector<PlanVariant> variants;
    uint64_t variant_clock = 0;
    void drop_variants() { for (PlanVariant &v : variants) if (v.exec) cudaGraphExecDestroy(v.exec); variants.clear(); }
    std::string *plan_log = nullptr;         /* where launch_plan records kernel names while capturing */

    /* multi-GPU */
    void *comm = nullptr;
    int rank = 0, n_ranks = 1;
    /* NVLink peer-memory mix bus */
    float *d_xchg = nullptr;
    PeerTable peers{};
    bool p2p = false;
    uint32_t *d_p2p_error = nullptr;

    /* stats / timing */
    oscen_b200_stats stats{};
    /* per-block CUDA events on the engine stream: a ring of {start, graph kernel done, block done}, read back in bulk by
       timing_drain() so that a timed loop does not synchronise with the device every step */
    bool timing = false;
    std::vector<cudaEvent_t> tev;                /* 3 per slot, created on demand */
    uint32_t t_count = 0, t_drained = 0;         /* blocks recorded / already handed out since enable_timing(1) */
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk = nullptr;   /* the current block's slot */
    uint32_t *d_flush = nullptr;
    bool ev_valid = false;
    const char *last_kernel = "";                /* the graph kernel (variant) the last block ran */
};

namespace {

#define CK(e, call)                                                                        \
    do {                                                                                   \
        cudaError_t _s = (call);                                                           \
        if (_s != cudaSuccess) {                                                           \
            (e)->err = std::string(#call) + ": " + cudaGetErrorString(_s);                 \
            return OSCEN_B200_E_CUDA;                                                      \
        }                                                                                  \
    } while (0)

int arg_error(oscen_b200_engine *e, const char *msg) { e->err = msg; return OSCEN_B200_E_ARG; }

/* next slot of the pinned staging ring with room for `bytes`; *slot is handed to stage_submit after the copy is enqueued */
int stage_acquire(oscen_b200_engine *e, size_t bytes, void **out, int *slot) {
    const int k = e->stage_next;
    e->stage_next = (k + 1) % oscen_b200_engine::kStage;
    if (e->stage_busy[k]) { CK(e, cudaEventSynchronize(e->stage_ev[k])); e->stage_busy[k] = false; }
    if (e->stage_cap[k] < bytes) {
        /* grow EVERY slot of the ring now: pinned allocations cost milliseconds, and a bank that sent one bank-wide burst
           (note-on) will send another (note-off) that lands on a different slot — it must not pay for it inside a block */
        const size_t cap = std::max<size_t>(bytes + bytes / 2, 64 * 1024);
        for (int j = 0; j < oscen_b200_engine::kStage; ++j) {
            if (e->stage_cap[j] >= cap) continue;
            if (e->stage_busy[j]) { CK(e, cudaEventSynchronize(e->stage_ev[j])); e->stage_busy[j] = false; }
            if (e->stage_h[j]) cudaFreeHost(e->stage_h[j]);
            e->stage_h[j] = nullptr; e->stage_cap[j] = 0;
            CK(e, cudaMallocHost(&e->stage_h[j], cap));
            e->stage_cap[j] = cap;
        }
    }
    if (!e->stage_ev[k]) CK(e, cudaEventCreateWithFlags(&e->stage_ev[k], cudaEventDisableTiming));
    *out = e->stage_h[k]; *slot = k;
    return 0;
}
int stage_submit(oscen_b200_engine *e, int slot) {
    CK(e, cudaEventRecord(e->stage_ev[slot], e->stream));
    e->stage_busy[slot] = true;
    return 0;
}

/* initial SoA state of one voice after G::new() + G::init(sr) */
std::vector<uint32_t> initial_state(uint32_t kind, float sr, const plan::Graph *custom, uint32_t n_state) {
    std::vector<float> s(n_state, 0.0f);
    auto bits = [](float f) { uint32_t u; memcpy(&u, &f, 4); return u; };
    auto adsr = [&](uint32_t base) { s[base + 0] = 0.0f; s[base + 1] = 0.0f; s[base + 2] = 0.0f; s[base + 3] = 1.0f; };
    auto tpt = [&](uint32_t base, float cutoff, float q) {
        TptState t; tpt_init(t, sr, cutoff, q);
        s[base + 0] = t.z0; s[base + 1] = t.z1; s[base + 2] = t.cur_cutoff; s[base + 3] = t.cur_q;
        s[base + 4] = t.h; s[base + 5] = t.g; s[base + 6] = t.k;
    };
    switch (kind) {
    case OSCEN_B200_GRAPH_FM_VOICE:
        adsr(FM_E3); adsr(FM_E2); adsr(FM_E1); adsr(FM_EF);
        tpt(FM_TPT, 2000.0f, 0.707f);
        break;
    case OSCEN_B200_GRAPH_TWIN_PEAKS: {
        const float cut[2] = {1000.0f, 1900.0f};
        for (int k = 0; k < 2; ++k) {
            Lp18State f; lp18_init(f, cut[k], 0.54f, sr);
            const uint32_t b = k ? TW_B : TW_A;
            s[b + 0] = f.z0; s[b + 1] = f.z1; s[b + 2] = f.z2; s[b + 3] = f.g; s[b + 4] = f.h;
            s[b + 5] = f.last_cutoff; s[b + 6] = f.last_fmod; s[b + 7] = f.last_resonance; s[b + 8] = f.resonance;
        }
    } break;
    case OSCEN_B200_GRAPH_SAT_1X: case OSCEN_B200_GRAPH_SAT_4X: break;
    case OSCEN_B200_GRAPH_EPIANO:
        for (int i = 0; i < 32; ++i) { s[EP_ORE + i] = 1.0f; s[EP_MRE + i] = 1.0f; }
        s[EP_STEP] = 64.0f;
        break;
    case OSCEN_B200_GRAPH_SYNTH: case OSCEN_B200_GRAPH_SYNTH_ADSR:
        tpt(SY_TPT, 1200.0f, 0.707f);
        adsr(SY_ENV);
        break;
    case OSCEN_B200_GRAPH_CUSTOM:
        /* G::new() builds every node from its constructor expression, G::init(sr) = set_sample_rate + prepare()
           (codegen/mod.rs:1309-1350); only filters and envelopes hold non-zero state afterwards */
        for (const plan::Node &nd : custom->nodes) {
            if (nd.type == plan::NodeType::Adsr) adsr(nd.state_base);
            else if (nd.type == plan::NodeType::Tpt) {
                /* an inner-rate (`* N`) node is initialised at sr * N (codegen/emit_struct.rs:578-581) */
                TptState t; tpt_init(t, sr * (float)(1 << nd.rate_log2), nd.init[0], nd.init[1]);
                const uint32_t b = nd.state_base;
                s[b + 0] = t.z0; s[b + 1] = t.z1; s[b + 2] = t.cur_cutoff; s[b + 3] = t.cur_q; s[b + 4] = t.h; s[b + 5] = t.g; s[b + 6] = t.k;
            }
            else if (nd.type == plan::NodeType::IirLowpass) {
                IirLpState f; iirlp_init(f, sr * (float)(1 << nd.rate_log2), nd.init[0], nd.init[1]);
                const uint32_t b = nd.state_base;
                s[b + 0] = f.b0; s[b + 1] = f.b1; s[b + 2] = f.b2; s[b + 3] = f.a1; s[b + 4] = f.a2;       /* v1 v2 frame_counter = 0 */
            }
            else if (nd.type == plan::NodeType::AmpSource) {
                s[nd.state_base + AS_STEP] = 64.0f;              /* interpolation_step: INTERPOLATION_STEPS (electric_piano_voice.rs:218) */
            }
            else if (nd.type == plan::NodeType::OscBank) {        /* oscillators / multipliers: Complex::one() (electric_piano_voice.rs:104-105) */
                for (uint32_t i = 0; i < kEpHarmonics; ++i) { s[nd.state_base + OB_ORE + i] = 1.0f; s[nd.state_base + OB_MRE + i] = 1.0f; }
            }
            else if (nd.type == plan::NodeType::Lp18) {
                Lp18State f; lp18_init(f, nd.init[0], nd.init[1], sr * (float)(1 << nd.rate_log2));
                const uint32_t b = nd.state_base;
                s[b + 0] = f.z0; s[b + 1] = f.z1; s[b + 2] = f.z2; s[b + 3] = f.g; s[b + 4] = f.h;
                s[b + 5] = f.last_cutoff; s[b + 6] = f.last_fmod; s[b + 7] = f.last_resonance; s[b + 8] = f.resonance;
            }
        }
        break;
    default: break;                                        /* clip graphs: resampler state starts at zero */
    }
    std::vector<uint32_t> out(s.size());
    for (size_t i = 0; i < s.size(); ++i) out[i] = bits(s[i]);
    /* ADSR stage/rem are integer fields: Idle(0), 0 */
    return out;
}

int fill_row(oscen_b200_engine *e, uint32_t input, uint32_t first, uint32_t count, float value) {
    uint32_t bits; memcpy(&bits, &value, 4);
    const uint32_t grid = std::min<uint32_t>((count + 255) / 256, 1184);
    fill_u32_kernel<<<grid, 256, 0, e->stream>>>((uint32_t *)(e->d_pv + (size_t)input * e->cfg.n_voices + first), count, bits);
    e->stats.kernel_launches++;
    CK(e, cudaGetLastError());
    return 0;
}

void fill_args(oscen_b200_engine *e, KernelArgs &a) {
    memset(&a, 0, sizeof a);
    a.n_voices = e->cfg.n_voices; a.vpi = e->cfg.voices_per_instance; a.n_inst = e->n_inst; a.channels = e->channels;
    a.voice_offset = e->voice_offset; a.lcg_seed = e->lcg_seed;
    a.state = e->d_state; a.pv = e->d_pv; a.events = e->d_events; a.ptab = e->d_ptab; a.ctl = e->d_ctl;
    a.stream_in = e->stream_in_ptr;
    a.voice_out = e->d_voice_out; a.inst_out = e->d_inst_out; a.inst_state = e->d_inst_state;
    a.mix_partial = e->d_mix_partial;
    a.assets = e->d_assets;
    a.utab = e->d_utab;
    {   /* diagnostics switches, read per call so that a test can flip them between engines:
           OSCEN_B200_FM_SKEW=0 (bit 0: segmented FM kernel without the skewed loop), OSCEN_B200_TMA_STATE=1 (bit 1: the
           general / legacy FM kernels hand their state over with TMA bulk copies, kernel_fm.cuh) */
        const char *x = getenv("OSCEN_B200_FM_SKEW"), *t = getenv("OSCEN_B200_TMA_STATE");
        a.pad0 = ((x && x[0] == '0') ? 1u : 0u) | ((t && t[0] == '1') ? 2u : 0u);
    }
    for (uint32_t i = 0; i < kMaxInputs; ++i) a.u.p[i] = e->uni[i];
    a.u.sr = e->cfg.sample_rate;
    if (e->cfg.graph_kind == OSCEN_B200_GRAPH_FM_VOICE) {
        const float *p = e->uni;
        adsr_derive(a.u.env[0], p[I_OP3_A], p[I_OP3_D], p[I_OP3_R], a.u.sr);
        adsr_derive(a.u.env[1], p[I_OP2_A], p[I_OP2_D], p[I_OP2_R], a.u.sr);
        adsr_derive(a.u.env[2], p[I_OP1_A], p[I_OP1_D], p[I_OP1_R], a.u.sr);
        adsr_derive(a.u.env[3], p[I_F_A], p[I_F_D], p[I_F_R], a.u.sr);
    }
}

constexpr size_t kEpDyn = kEpDynFloats * sizeof(float);

/* the cross-rate clipper family: one instantiation per (policy, log2 factor) */
template <int P, int L>
void clip_launch_one(oscen_b200_engine *e, const KernelArgs &a, uint32_t grid, bool set_attr) {
    if (set_attr) cudaFuncSetAttribute(clip_kernel<P, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)clip_dyn_bytes(P, L));
    else clip_kernel<P, L><<<grid, kCta, clip_dyn_bytes(P, L), e->stream>>>(a);
}
template <int P>
void clip_launch_policy(oscen_b200_engine *e, const KernelArgs &a, uint32_t grid, int l, bool set_attr) {
    switch (l) {
    case 1: clip_launch_one<P, 1>(e, a, grid, set_attr); break;
    case 2: clip_launch_one<P, 2>(e, a, grid, set_attr); break;
    default: clip_launch_one<P, 3>(e, a, grid, set_attr); break;
    }
}
void launch_clip(oscen_b200_engine *e, const KernelArgs &a, uint32_t grid, bool set_attr = false) {
    const uint32_t kind = e->cfg.graph_kind;
    const int l = clip_log2f(kind);
    if (l == 0) { clip_launch_one<POL_LATCH, 0>(e, a, grid, set_attr); return; }
    switch (clip_policy(kind)) {
    case POL_SINC: clip_launch_policy<POL_SINC>(e, a, grid, l, set_attr); break;
    case POL_SINC_IIR: clip_launch_policy<POL_SINC_IIR>(e, a, grid, l, set_attr); break;
    case POL_LINEAR: clip_launch_policy<POL_LINEAR>(e, a, grid, l, set_attr); break;
    default: clip_launch_policy<POL_LATCH>(e, a, grid, l, set_attr); break;
    }
}

/* ---- fused plan: one graph kernel + the mix-bus reduction */
int launch_fused(oscen_b200_engine *e, bool general) {
    KernelArgs a; fill_args(e, a);
    const uint32_t grid = e->n_cta;
    e->n_cta_mix = grid;
    switch (e->cfg.graph_kind) {
    case OSCEN_B200_GRAPH_FM_VOICE:
        (void)general;
        switch (e->fm_variant) {                              /* chosen by fm_select_variant() in process_block */
        case FM_SEG: fm_voice_seg_kernel<false, false><<<grid, kCta, 0, e->stream>>>(a); break;
        case FM_SEG_FMOD: fm_voice_seg_kernel<false, true><<<grid, kCta, 0, e->stream>>>(a); break;
        case FM_SEG_RAMP: fm_voice_seg_kernel<true, true><<<grid, kCta, 0, e->stream>>>(a); break;
        case FM_LEGACY: fm_voice_kernel<false><<<grid, kCta, 0, e->stream>>>(a); break;
        default: fm_voice_kernel<true><<<grid, kCta, 0, e->stream>>>(a); break;
        }
        e->last_kernel = fm_variant_name(e->fm_variant);
        break;
    case OSCEN_B200_GRAPH_TWIN_PEAKS: twin_peaks_kernel<<<grid, kCta, 0, e->stream>>>(a); e->last_kernel = "twin_peaks_kernel"; break;
    case OSCEN_B200_GRAPH_SAT_1X: sat_kernel<1><<<grid, kCta, 0, e->stream>>>(a); e->last_kernel = "sat_kernel<1>"; break;
    case OSCEN_B200_GRAPH_SAT_4X: sat_kernel<4><<<grid, kCta, 0, e->stream>>>(a); e->last_kernel = "sat_kernel<4>"; break;
    case OSCEN_B200_GRAPH_EPIANO: epiano_kernel<<<grid, kEpThreads, kEpDyn, e->stream>>>(a); e->last_kernel = "epiano_kernel"; break;
    case OSCEN_B200_GRAPH_SYNTH: synth_kernel<false><<<grid, kCta, 0, e->stream>>>(a); e->last_kernel = "synth_kernel<false>"; break;
    case OSCEN_B200_GRAPH_SYNTH_ADSR: synth_kernel<true><<<grid, kCta, 0, e->stream>>>(a); e->last_kernel = "synth_kernel<true>"; break;
    case OSCEN_B200_GRAPH_CUSTOM: {
        /* the specialised variant renders the blocks in which every input of its mask still holds one value for all voices
           (value events and per-voice set_value clear `uniform`) and none of them ramps; `general` frames inside it
           (gates, ramps or value events on the OTHER inputs) are handled by the kernel itself */
        bool uni = e->jit_uni.function != nullptr && (e->h_ctl->ramp_mask & e->jit_uni_mask) == 0u;
        for (uint32_t i = 0; uni && i < e->n_inputs; ++i) if (((e->jit_uni_mask >> i) & 1u) && !e->uniform[i]) uni = false;
        const int rc = (uni ? e->jit_uni : e->jit).launch(grid, kCta, (void *)e->stream, &a);
        if (rc != 0) { e->err = "cuLaunchKernel (generated kernel): " + jit::driver_error(rc); return OSCEN_B200_E_CUDA; }
        e->last_kernel = uni ? e->jit_uni_name.c_str() : kJitKernelName;
    } break;
    default: launch_clip(e, a, grid); e->last_kernel = "clip_kernel"; break;
    }
    e->stats.kernel_launches++;
    CK(e, cudaGetLastError());
    return 0;
}

/* ---- node plan: the compiled launch sequence (plan.hpp) — one launch per node type per ready wave */
template <int OP>
void launch_ew(oscen_b200_engine *e, const NodeArgs &n, bool vec) {
    if (vec) ew_node_vec_kernel<OP><<<148u * 8u, 256, 0, e->stream>>>(n);
    else { NodeArgs4 a4{}; a4.n[0] = n; ew_node_kernel<OP><<<dim3(e->n_cta, 1), kCta, 0, e->stream>>>(a4); }
    e->stats.kernel_launches++;
}

/* TMA descriptors of the node plan's stream buffers (kernel_nodes.cuh, *_node_tma_kernel): each buffer is a 2-D f32 tensor
   [kMaxBlock frames][n_voices], voice fastest, box = [kTmaF frames][kCta voices], no swizzle (a thread reads its own column),
   out-of-range voices zero-filled.  The row pitch n_voices * 4 must be a multiple of 16 bytes; otherwise, or when the driver
   has no cuTensorMapEncodeTiled, d_tmaps stays null and the register-prefetch kernels run.  OSCEN_B200_NODE_TMA=0 does the same. */
int build_stream_tmaps(oscen_b200_engine *e) {
    const char *sw = getenv("OSCEN_B200_NODE_TMA");
    const uint32_t n = e->cfg.n_voices;
    if ((sw && sw[0] == '0') || n % 4u != 0u) return 0;
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        return 0;
    }
    std::vector<CUtensorMap> maps(e->d_streams.size());
    for (size_t i = 0; i < maps.size(); ++i) {
        const cuuint64_t dims[2] = {n, kMaxBlock};
        const cuuint64_t strides[1] = {(cuuint64_t)n * sizeof(float)};
        const cuuint32_t box[2] = {kCta, kTmaF}, estr[2] = {1, 1};
        const CUresult r = ((encode_fn)fn)(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, e->d_streams[i], dims, strides, box, estr,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return 0;                       /* no descriptors: the register-prefetch kernels run */
    }
    CK(e, cudaMalloc(&e->d_tmaps, maps.size() * sizeof(CUtensorMap)));
    CK(e, cudaMemcpy(e->d_tmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    return 0;
}

/* A/B and tests: OSCEN_B200_NODE_PLAIN=0 keeps the generic node kernels on event-free blocks too */
static bool no_plain_nodes() { const char *x = getenv("OSCEN_B200_NODE_PLAIN"); return x && x[0] == '0'; }   /* read per block: a test flips it between engines */

int launch_plan(oscen_b200_engine *e, bool value_events, bool no_events, uint32_t ramp_mask) {
    const plan::Graph &g = e->graph;
    const plan::Plan &p = e->plan;
    float **S = e->d_streams.data();
    const uint32_t grid = e->n_cta, n = e->cfg.n_voices;
    const bool vec = !value_events && (n % 4u == 0u);
    /* the graph's stream input: host/device block, or generated by the LCG source node */
    const float *gin = nullptr;
    if (e->has_stream_in) {
        if (e->lcg_on) {
            lcg_source_kernel<<<grid, kCta, 0, e->stream>>>(e->d_state, S[p.n_streams], e->d_ctl, n, e->lcg_field);
            e->stats.kernel_launches++;
            gin = S[p.n_streams];
        } else gin = e->stream_in_ptr;
    }
    auto resolve = [&](const plan::StreamRef &r) -> float * {
        if (r.node == plan::kNone) return nullptr;
        if (r.node == plan::kGraphInput) return const_cast<float *>(gin);
        return S[p.first_stream[r.node] + r.port];
    };
    auto tmap_of = [&](const plan::StreamRef &r) -> const void * {     /* descriptor of an inter-node stream (not the graph input) */
        if (!e->d_tmaps || r.node < 0) return nullptr;
        return e->d_tmaps + (p.first_stream[r.node] + r.port);
    };
    /* a wave can run the TMA-tiled kernel when every connected stream of every node in it has a descriptor */
    auto wave_has_tmaps = [&](const NodeArgs4 &w, uint32_t cnt) {
        for (uint32_t k = 0; k < cnt; ++k) {
            const NodeArgs &x = w.n[k];
            if (!x.tm_out0 || (x.in0 && !x.tm_in0) || (x.in1 && !x.tm_in1)) return false;
        }
        return true;
    };
    auto args_of = [&](int idx) {
        const plan::Node &nd = g.nodes[idx];
        NodeArgs a{};
        a.n_voices = n; a.state_base = nd.state_base; a.ctl = e->d_ctl; a.events = e->d_events;
        a.ptab = e->d_ptab; a.pv = e->d_pv; a.state = e->d_state; a.sr = e->cfg.sample_rate;
        a.in0 = resolve(nd.in[0]); a.in1 = resolve(nd.in[1]); a.in2 = resolve(nd.in[2]);
        a.out0 = S[p.first_stream[idx]];
        a.out1 = plan::n_outputs(nd.type) > 1 ? S[p.first_stream[idx] + 1] : nullptr;
        a.tm_in0 = tmap_of(nd.in[0]); a.tm_in1 = tmap_of(nd.in[1]);
        a.tm_out0 = e->d_tmaps ? (const void *)(e->d_tmaps + p.first_stream[idx]) : nullptr;
        for (int k = 0; k < plan::kMaxParams; ++k) { a.pid[k] = nd.pid[k]; a.cst[k] = nd.cst[k]; }
        a.gate = nd.gate ? 1u : 0u; a.q = nd.cst[1];
        if (nd.type == plan::NodeType::Delay) { a.gate = nd.ring_cap; a.q = nd.max_delay; }
        if (nd.type == plan::NodeType::SamplePlayer) {            /* the slot table rides in in1, the slot index in gate */
            a.in1 = reinterpret_cast<const float *>(e->d_assets); a.gate = nd.asset >= 0 ? (uint32_t)nd.asset : 0xFFFFFFFFu;
        }
        return a;
    };
    for (const plan::Launch &l : p.launches) {
        NodeArgs4 a4{};
        const uint32_t cnt = (uint32_t)l.nodes.size();
        for (uint32_t k = 0; k < cnt; ++k) a4.n[k] = args_of(l.nodes[k]);
        const dim3 g2(grid, cnt);
        auto log = [&](const char *name) { if (e->plan_log) { *e->plan_log += ' '; *e->plan_log += name; } };
        /* plain wave: the block carries no event and none of the wave's value inputs ramps in it */
        bool plain = no_events;
        for (uint32_t k = 0; k < cnt && plain; ++k)
            for (int q = 0; q < plan::kMaxParams; ++q) if (a4.n[k].pid[q] != kNoInput && ((ramp_mask >> a4.n[k].pid[q]) & 1u)) plain = false;
        switch (l.type) {
        case plan::NodeType::Adsr:
            if (plain) { adsr_node_plain_kernel<<<g2, kCta, 0, e->stream>>>(a4); log("adsr_node_plain_kernel"); }
            else { adsr_node_kernel<<<g2, kCta, 0, e->stream>>>(a4); log("adsr_node_kernel"); }
            e->stats.kernel_launches++; break;
        case plan::NodeType::FmOperator:
            if (plain && wave_has_tmaps(a4, cnt)) { fmop_node_tma_kernel<<<g2, kCta, 0, e->stream>>>(a4); log("fmop_node_tma_kernel"); }
            else if (plain) { fmop_node_plain_kernel<<<g2, kCta, 0, e->stream>>>(a4); log("fmop_node_plain_kernel"); }
            else { fmop_node_kernel<<<g2, kCta, 0, e->stream>>>(a4); log("fmop_node_kernel"); }
            e->stats.kernel_launches++; break;
        case plan::NodeType::Tpt: {
            bool zero_fmod = true;                                      /* every filter of the wave: no f_mod stream, f_mod == 0 */
            for (uint32_t k = 0; k < cnt; ++k) zero_fmod = zero_fmod && a4.n[k].in2 == nullptr && a4.n[k].pid[2] == kNoInput && a4.n[k].cst[2] == 0.0f;
            if (plain && zero_fmod && wave_has_tmaps(a4, cnt)) { tpt_node_tma_kernel<<<g2, kCta, 0, e->stream>>>(a4); log("tpt_node_tma_kernel"); }
            else if (plain && zero_fmod) { tpt_node_plain_kernel<true><<<g2, kCta, 0, e->stream>>>(a4); log("tpt_node_plain_kernel<zero_fmod>"); }
            else if (plain) { tpt_node_plain_kernel<false><<<g2, kCta, 0, e->stream>>>(a4); log("tpt_node_plain_kernel<fmod>"); }
            else { tpt_node_kernel<<<g2, kCta, 0, e->stream>>>(a4); log("tpt_node_kernel"); }
            e->stats.kernel_launches++; } break;
        case plan::NodeType::PolyBlepSine: polyblep_node_kernel<WAVE_SINE><<<g2, kCta, 0, e->stream>>>(a4); e->stats.kernel_launches++; break;
        case plan::NodeType::PolyBlepSaw: polyblep_node_kernel<WAVE_SAW><<<g2, kCta, 0, e->stream>>>(a4); e->stats.kernel_launches++; break;
        case plan::NodeType::PolyBlepSquare: polyblep_node_kernel<WAVE_SQUARE><<<g2, kCta, 0, e->stream>>>(a4); e->stats.kernel_launches++; break;
        case plan::NodeType::PolyBlepTriangle: polyblep_node_kernel<WAVE_TRIANGLE><<<g2, kCta, 0, e->stream>>>(a4); e->stats.kernel_launches++; break;
        case plan::NodeType::Lp18: lp18_node_kernel<<<g2, kCta, 0, e->stream>>>(a4); e->stats.kernel_launches++; break;
        case plan::NodeType::OscSine: osc_node_kernel<OSC_SINE><<<g2, kCta, 0, e->stream>>>(a4); e->stats.kernel_launches++; break;
        case plan::NodeType::OscSquare: osc_node_kernel<OSC_SQUARE><<<g2, kCta, 0, e->stream>>>(a4); e->stats.kernel_launches++; break;
        case plan::NodeType::OscSaw: osc_node_kernel<OSC_SAW><<<g2, kCta, 0, e->stream>>>(a4); e->stats.kernel_launches++; break;
        case plan::NodeType::IirLowpass: iirlp_node_kernel<<<g2, kCta, 0, e->stream>>>(a4); e->stats.kernel_launches++; break;
        case plan::NodeType::Delay: delay_node_kernel<<<g2, kCta, 0, e->stream>>>(a4); e->stats.kernel_launches++; break;
        case plan::NodeType::SamplePlayer: sample_player_node_kernel<<<g2, kCta, 0, e->stream>>>(a4); e->stats.kernel_launches++; break;
        default:
            for (uint32_t k = 0; k < cnt; ++k) {                       /* stateless nodes: one stream kernel each */
                const NodeArgs &a = a4.n[k];
                switch (l.type) {
                case plan::NodeType::Gain:
                    if (a.in1 != nullptr) launch_ew<EW_MUL>(e, a, vec); else launch_ew<EW_GAIN>(e, a, vec);
                    break;
                case plan::NodeType::AddValue: launch_ew<EW_ADD_VALUE>(e, a, vec); break;
                case plan::NodeType::Mixer: launch_ew<EW_MIXER>(e, a, vec); break;
                case plan::NodeType::Sub: launch_ew<EW_SUB>(e, a, vec); break;
                case plan::NodeType::Crossfade: launch_ew<EW_CROSSFADE>(e, a, vec); break;
                case plan::NodeType::Expr:
                    switch (g.nodes[l.nodes[k]].op) {
                    case plan::ExprOp::AddSS: launch_ew<EW_MIXER>(e, a, vec); break;
                    case plan::ExprOp::SubSS: launch_ew<EW_SUB>(e, a, vec); break;
                    case plan::ExprOp::MulSS: launch_ew<EW_MUL>(e, a, vec); break;
                    case plan::ExprOp::DivSS: launch_ew<EW_DIV>(e, a, vec); break;
                    case plan::ExprOp::AddSV: launch_ew<EW_ADD_VALUE>(e, a, vec); break;
                    case plan::ExprOp::SubSV: launch_ew<EW_SUBV>(e, a, vec); break;
                    case plan::ExprOp::MulSV: launch_ew<EW_GAIN>(e, a, vec); break;
                    case plan::ExprOp::DivSV: launch_ew<EW_DIVV>(e, a, vec); break;
                    case plan::ExprOp::SubVS: launch_ew<EW_RSUBV>(e, a, vec); break;
                    case plan::ExprOp::DivVS: launch_ew<EW_RDIVV>(e, a, vec); break;
                    case plan::ExprOp::Value: launch_ew<EW_VALUE>(e, a, vec); break;
                    case plan::ExprOp::Tanh: launch_ew<EW_TANH>(e, a, vec); break;
                    case plan::ExprOp::Abs: launch_ew<EW_ABS>(e, a, vec); break;
                    case plan::ExprOp::Clamp: launch_ew<EW_CLAMP>(e, a, vec); break;
                    case plan::ExprOp::MinV: launch_ew<EW_MINV>(e, a, vec); break;
                    case plan::ExprOp::Sqrt: launch_ew<EW_SQRT>(e, a, vec); break;
                    case plan::ExprOp::Sin: launch_ew<EW_SIN>(e, a, vec); break;
                    case plan::ExprOp::Cos: launch_ew<EW_COS>(e, a, vec); break;
                    case plan::ExprOp::Exp: launch_ew<EW_EXP>(e, a, vec); break;
                    default: launch_ew<EW_MAXV>(e, a, vec); break;
                    }
                    break;
                default: launch_ew<EW_HARDCLIP>(e, a, vec); break;
                }
            }
            break;
        }
    }
    e->final_stream = S[p.first_stream[g.output.node] + g.output.port];
    { KernelArgs a; fill_args(e, a); a.voice_out = nullptr;           /* voices.audio_out -> audio_out */
      instance_mix_node_kernel<<<grid, kCta, 0, e->stream>>>(a, e->final_stream); e->stats.kernel_launches++; e->n_cta_mix = grid; }
    commit_value_events_kernel<<<296, 256, 0, e->stream>>>(e->d_pv, e->d_events, e->d_ctl, n, e->n_inputs);
    e->stats.kernel_launches++;
    CK(e, cudaGetLastError());
    return 0;
}

/* the mix bus of this rank's voices; with a peer-memory exchange set up (comm_p2p_import) the same launch also is the
   collective: compute (partial reduction) + exchange over NVLink peer memory in ONE kernel, either plan */
int launch_mix_reduce(oscen_b200_engine *e, uint32_t frames) {
    if (e->p2p)
        mix_exchange_kernel<<<kMaxBlock * e->channels, 256, 0, e->stream>>>(e->d_mix_partial, e->d_mix, e->n_cta_mix, e->channels, e->d_ctl,
                                                        e->peers, (uint32_t)e->rank, (uint32_t)e->n_ranks, e->d_p2p_error + 2, e->d_p2p_error + 1, e->d_p2p_error);
    else
        mix_reduce_kernel<<<kMaxBlock * e->channels, 256, 0, e->stream>>>(e->d_mix_partial, e->d_mix, e->n_cta_mix, e->channels, e->d_ctl);
    (void)frames;
    e->stats.kernel_launches++;
    CK(e, cudaGetLastError());
    return 0;
}

}  // namespace

/* ======================================================================== C ABI == */
extern "C" {

int oscen_b200_abi_version(void) { return OSCEN_B200_ABI_VERSION; }

int oscen_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

uint32_t oscen_b200_n_inputs(uint32_t kind) { return kind < 7 ? kNInputs[kind] : 0; }

int oscen_b200_kind_info(uint32_t kind, float sample_rate, oscen_b200_kind_tables *info, float *defaults, uint32_t *state_words, size_t state_cap) {
    if (kind >= OSCEN_B200_GRAPH_CUSTOM || !(sample_rate > 0.0f)) return OSCEN_B200_E_ARG;
    const uint32_t n_state = n_state_of(kind);
    if (info) {
        memset(info, 0, sizeof *info);
        info->n_inputs = kind < 7 ? kNInputs[kind] : 0u; info->n_state = n_state;
        info->channels = kind == OSCEN_B200_GRAPH_EPIANO ? 2u : 1u;
        info->dyn_smem_bytes = kind == OSCEN_B200_GRAPH_EPIANO ? (uint32_t)kEpDyn
                               : is_clip(kind) ? (uint32_t)clip_dyn_bytes(clip_policy(kind), clip_log2f(kind)) : 0u;
        info->has_stream_in = has_stream_in(kind) ? 1u : 0u;
    }
    if (defaults) for (uint32_t i = 0; i < kMaxInputs; ++i) defaults[i] = kind < 7 ? kDefaults[kind][i] : 0.0f;
    if (state_words) {
        if (n_state > state_cap) return OSCEN_B200_E_ARG;
        const std::vector<uint32_t> init = initial_state(kind, sample_rate, nullptr, n_state);
        memcpy(state_words, init.data(), init.size() * sizeof(uint32_t));
    }
    return 0;
}

const char *oscen_b200_last_error(const oscen_b200_engine *e) { return e ? e->err.c_str() : g_create_error.c_str(); }

void oscen_b200_destroy(oscen_b200_engine *e) {
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
    if (e->p2p) for (int r = 0; r < e->n_ranks; ++r) if (r != e->rank && e->peers.peer[r]) cudaIpcCloseMemHandle(e->peers.peer[r]);
    cudaFree(e->d_xchg); cudaFree(e->d_p2p_error);
    e->drop_variants();
    e->jit.unload(); e->jit_uni.unload();
    for (float *p : e->d_streams) cudaFree(p);
    cudaFree(e->d_tmaps);
    cudaFree(e->d_pv); cudaFree(e->d_ptab); cudaFree(e->d_state); cudaFree(e->d_inst_state);
    cudaFree(e->d_events); cudaFree(e->d_ctl); cudaFree(e->d_stream_in);
    cudaFree(e->d_voice_out); cudaFree(e->d_inst_out); cudaFree(e->d_mix_partial); cudaFree(e->d_mix);
    cudaFreeHost(e->h_ctl); cudaFreeHost(e->h_mix);
    for (int k = 0; k < oscen_b200_engine::kStage; ++k) { if (e->stage_h[k]) cudaFreeHost(e->stage_h[k]); if (e->stage_ev[k]) cudaEventDestroy(e->stage_ev[k]); }
    cudaFree(e->d_utab);
    for (cudaEvent_t ev : e->tev) if (ev) cudaEventDestroy(ev);
    cudaFree(e->d_flush);
    cudaFree(e->d_assets);
    for (float *b : e->d_asset_buf) cudaFree(b);
    cudaFree(e->d_midi_msgs); cudaFree(e->d_midi_keys_in); cudaFree(e->d_midi_keys); cudaFree(e->d_midi_idx_in); cudaFree(e->d_midi_idx);
    cudaFree(e->d_cub_tmp); cudaFree(e->d_alloc_word); cudaFree(e->d_alloc_age); cudaFree(e->d_alloc_cur_age);
    cudaFree(e->d_h_note); cudaFree(e->d_h_freq); cudaFree(e->d_note_freq);
    cudaFree(e->d_ev_count); cudaFree(e->d_ev_offset); cudaFree(e->d_midi_dropped);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

static int create_impl(const oscen_b200_config *cfg, const oscen_b200_graph *custom, oscen_b200_engine **out) {
    if (!cfg || !out) { g_create_error = "null argument"; return OSCEN_B200_E_ARG; }
    *out = nullptr;
    oscen_b200_config cfg_local;
    if (custom && cfg && custom->g.vpi) {
        /* a polyphonic graph brings its own voices_per_instance (the voice array size); 0 in the config means "the graph's" */
        if (cfg->voices_per_instance != 0 && cfg->voices_per_instance != custom->g.vpi) {
            g_create_error = "voices_per_instance must be " + std::to_string(custom->g.vpi) + " (the graph's voice array size) or 0"; return OSCEN_B200_E_ARG;
        }
        cfg_local = *cfg; cfg_local.voices_per_instance = custom->g.vpi; cfg = &cfg_local;
    }
    if ((custom == nullptr) != (cfg->graph_kind != OSCEN_B200_GRAPH_CUSTOM)) {
        g_create_error = "OSCEN_B200_GRAPH_CUSTOM engines are created with oscen_b200_create_from_graph (and only those)"; return OSCEN_B200_E_ARG;
    }
    if ((cfg->keep & OSCEN_B200_KEEP_VOICE_OUT) && (uint64_t)cfg->n_voices * kMaxBlock > 0xFFFFFFFFull) {
        g_create_error = "KEEP_VOICE_OUT: n_voices * 512 must fit 32 bits (<= 8 388 607 voices per engine)"; return OSCEN_B200_E_ARG;
    }
    if (cfg->plan == OSCEN_B200_PLAN_NODES && (uint64_t)cfg->n_voices * kMaxBlock > 0xFFFFFFFFull) {
        g_create_error = "PLAN_NODES: n_voices * 512 must fit 32 bits (<= 8 388 607 voices per engine; the stream buffers are indexed with 32 bits)"; return OSCEN_B200_E_ARG;
    }
    if (cfg->graph_kind >= OSCEN_B200_GRAPH_COUNT || cfg->n_voices == 0 || cfg->voices_per_instance == 0 ||
        kCta % cfg->voices_per_instance != 0 || (cfg->voices_per_instance & (cfg->voices_per_instance - 1)) != 0 || cfg->n_voices % cfg->voices_per_instance != 0 ||
        !(cfg->sample_rate > 0.0f) || cfg->plan > OSCEN_B200_PLAN_NODES) {
        g_create_error = "bad config (graph kind / voices / voices_per_instance must be a power of two <= 64 dividing n_voices / sample rate / plan)";
        return OSCEN_B200_E_ARG;
    }
    plan::Graph pg; plan::Plan pp;
    jit::Kernel jk;
    if (custom) {
        pg = custom->g; pp = custom->p;
        plan::assign_state_layout(pg, cfg->sample_rate);          /* Delay rings are sized for this engine's sample rate */
        if (cfg->plan == OSCEN_B200_PLAN_NODES)
            for (const plan::Node &nd : pg.nodes) for (int k = 0; k < plan::kMaxStreamIn; ++k) if (nd.in[k].node >= 0 && nd.fb[k]) {
                g_create_error = "PLAN_NODES: the feedback edge into `" + nd.name + "` needs the per-sample walk (node kernels render whole blocks): set plan = OSCEN_B200_PLAN_FUSED";
                return OSCEN_B200_E_ARG;
            }
        if (cfg->plan == OSCEN_B200_PLAN_FUSED) {
            /* generate the fused kernel and compile it for sm_100a before any device memory is taken */
            jk.source = codegen::fused_kernel_source(pg, pp, kJitKernelName);
            jk.dyn_smem = (unsigned)codegen::dyn_smem_bytes(pg);
            if (jk.dyn_smem > 200u * 1024u) { g_create_error = "PLAN_FUSED (generated kernel): the [sinc] cross-rate edges need more than 200 KB of shared memory per CTA"; return OSCEN_B200_E_ARG; }
            std::string jerr;
            if (!jit::compile_cubin(jk.source, jk.cubin, jk.log, jerr)) { g_create_error = "PLAN_FUSED (generated kernel): " + jerr; return OSCEN_B200_E_JIT; }
        }
        else if (pg.inner_log2 > 0) {
            g_create_error = "PLAN_NODES: a multirate graph (`* N` nodes) runs as a generated kernel: set plan = OSCEN_B200_PLAN_FUSED"; return OSCEN_B200_E_ARG;
        }
        if (cfg->plan == OSCEN_B200_PLAN_NODES) {
            bool fused_only = pg.post_tremolo;
            for (const plan::Node &nd : pg.nodes) fused_only = fused_only || plan::fused_only(nd.type);
            if (fused_only) {
                g_create_error = "PLAN_NODES: array-typed streams (`[f32; N]`) and the instance-level Tremolo -> Frame<2> tail exist in the generated kernel only (the node plan's stream buffers are one f32 per voice): set plan = OSCEN_B200_PLAN_FUSED";
                return OSCEN_B200_E_ARG;
            }
        }
    } else if (cfg->plan == OSCEN_B200_PLAN_NODES) {
        std::string perr;
        if (!plan::for_kind(cfg->graph_kind, pg)) {
            g_create_error = "PLAN_NODES: this graph (multirate / array streams) only has a fused kernel"; return OSCEN_B200_E_ARG;
        }
        if (!plan::compile(pg, pp, perr)) { g_create_error = "PLAN_NODES: " + perr; return OSCEN_B200_E_ARG; }
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        g_create_error = "no CUDA device: oscen_b200 has no CPU fallback"; return OSCEN_B200_E_NO_DEVICE;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "device ordinal out of range"; return OSCEN_B200_E_ARG; }

    oscen_b200_engine *e = new oscen_b200_engine();
    e->cfg = *cfg;
    const uint32_t kind = cfg->graph_kind, n = cfg->n_voices;
    e->has_stream_in = custom ? !pg.stream_input.empty() : has_stream_in(kind);
    if (custom) {
        e->midi_ok = pg.midi; e->midi_freq_input = pg.midi_freq_input;
        e->n_inputs = (uint32_t)pg.inputs.size();
        e->lcg_field = pg.n_state; e->n_state = pg.n_state + (e->has_stream_in ? 2u : 0u);
        e->n_assets = (uint32_t)pg.assets.size();
        if (e->n_state == 0) e->n_state = 1;                      /* stateless graph: keep the allocation non-empty */
    } else {
        e->n_inputs = kind < 7 ? kNInputs[kind] : 0u; e->n_state = n_state_of(kind);
        e->lcg_field = e->has_stream_in ? lcg_field(kind) : 0u;
    }
    e->n_inst = n / cfg->voices_per_instance;
    e->channels = kind == OSCEN_B200_GRAPH_EPIANO ? 2 : (custom ? pg.channels : 1);
    e->n_cta = (n + kCta - 1) / kCta;
    e->n_cta_mix = e->n_cta;
    int rc = [&]() -> int {
        CK(e, cudaSetDevice(cfg->device));
        CK(e, cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
        CK(e, cudaMalloc(&e->d_pv, (size_t)std::max<uint32_t>(e->n_inputs, 1u) * n * sizeof(float)));
        if (e->n_assets) {                                         /* every slot starts empty: its players are silent */
            CK(e, cudaMalloc(&e->d_assets, sizeof(AssetTable)));
            CK(e, cudaMemset(e->d_assets, 0, sizeof(AssetTable)));
        }
        CK(e, cudaMalloc(&e->d_ptab, (size_t)kMaxBlock * kMaxInputs * sizeof(float)));
        if (kind == OSCEN_B200_GRAPH_FM_VOICE) CK(e, cudaMalloc(&e->d_utab, (size_t)kMaxBlock * kFmRowFloats * sizeof(float)));
        CK(e, cudaMalloc(&e->d_state, (size_t)e->n_state * n * sizeof(float)));
        CK(e, cudaMalloc(&e->d_inst_state, (size_t)e->n_inst * sizeof(float)));
        CK(e, cudaMemsetAsync(e->d_inst_state, 0, (size_t)e->n_inst * sizeof(float), e->stream));
        CK(e, cudaMalloc(&e->d_ctl, sizeof(BlockCtl)));
        CK(e, cudaMallocHost(&e->h_ctl, sizeof(BlockCtl)));
        memset(e->h_ctl, 0, sizeof(BlockCtl));
        CK(e, cudaMemcpyAsync(e->d_ctl, e->h_ctl, sizeof(BlockCtl), cudaMemcpyHostToDevice, e->stream));
        e->events_cap = 1024;
        CK(e, cudaMalloc(&e->d_events, e->events_cap * sizeof(Event)));
        if (cfg->keep & OSCEN_B200_KEEP_VOICE_OUT) CK(e, cudaMalloc(&e->d_voice_out, (size_t)kMaxBlock * n * sizeof(float)));
        if (cfg->keep & OSCEN_B200_KEEP_INSTANCE_OUT) CK(e, cudaMalloc(&e->d_inst_out, (size_t)kMaxBlock * e->n_inst * e->channels * sizeof(float)));
        CK(e, cudaMalloc(&e->d_mix_partial, (size_t)e->n_cta * kMaxBlock * e->channels * sizeof(float)));
        CK(e, cudaMalloc(&e->d_mix, (size_t)kMaxBlock * e->channels * sizeof(float)));
        CK(e, cudaMemsetAsync(e->d_mix, 0, (size_t)kMaxBlock * e->channels * sizeof(float), e->stream));
        CK(e, cudaMallocHost(&e->h_mix, (size_t)kMaxBlock * e->channels * sizeof(float) + 16));   /* + peer-timeout flag */
        if (custom && cfg->plan == OSCEN_B200_PLAN_FUSED) {
            CK(e, cudaFree(nullptr));                              /* the runtime's primary context is current from here on */
            e->graph = pg; e->plan = pp;
            e->jit = std::move(jk);
            std::string jerr;
            if (!e->jit.load(kJitKernelName, jerr)) { e->err = "PLAN_FUSED (generated kernel): " + jerr; return OSCEN_B200_E_JIT; }
        }
        if (cfg->plan == OSCEN_B200_PLAN_NODES) {
            e->graph = pg; e->plan = pp;
            e->d_streams.resize((size_t)pp.n_streams + 1, nullptr);      /* + 1: device-generated graph input */
            for (auto &p : e->d_streams) CK(e, cudaMalloc(&p, (size_t)kMaxBlock * n * sizeof(float)));
            e->final_stream = e->d_streams[pp.first_stream[pg.output.node] + pg.output.port];
            { const int r = build_stream_tmaps(e); if (r) return r; }
        }
        CK(e, cudaFuncSetAttribute(epiano_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEpDyn));
        if (is_clip(kind)) { KernelArgs none{}; launch_clip(e, none, 0, true); CK(e, cudaGetLastError()); }
        /* new() + init(sr): default inputs, initial node state */
        for (uint32_t i = 0; i < kMaxInputs; ++i) {
            e->uni[i] = kind < 7 ? kDefaults[kind][i] : (custom && i < pg.inputs.size() ? pg.inputs[i].def : 0.0f);
            e->uniform[i] = true; e->ramped[i] = false;
            e->ramps[i].set_immediate(e->uni[i]);
        }
        for (uint32_t i = 0; i < e->n_inputs; ++i) { int r = fill_row(e, i, 0, n, e->uni[i]); if (r) return r; }
        std::vector<uint32_t> init = initial_state(kind, cfg->sample_rate, custom ? &pg : nullptr, e->n_state);
        uint32_t *d_init = nullptr;
        CK(e, cudaMalloc(&d_init, init.size() * sizeof(uint32_t)));
        cudaError_t st = cudaMemcpyAsync(d_init, init.data(), init.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream);
        if (st == cudaSuccess) {
            init_state_kernel<<<1184, 256, 0, e->stream>>>((uint32_t *)e->d_state, d_init, e->n_state, n);
            e->stats.kernel_launches++;
            st = cudaGetLastError();
        }
        if (st == cudaSuccess) st = cudaStreamSynchronize(e->stream);
        cudaFree(d_init);
        CK(e, st);
        return 0;
    }();
    if (rc) { g_create_error = e->err; oscen_b200_destroy(e); return rc; }
    *out = e;
    return 0;
}

int oscen_b200_create(const oscen_b200_config *cfg, oscen_b200_engine **out) { return create_impl(cfg, nullptr, out); }

/* ---- graph! DSL front end (dsl.hpp): text -> node plan; needs no GPU until create_from_graph */
int oscen_b200_graph_compile(const char *text, oscen_b200_graph **out, char *err, size_t err_cap) {
    return oscen_b200_graph_compile_named(text, nullptr, out, err, err_cap);
}
int oscen_b200_graph_compile_named(const char *text, const char *graph_name, oscen_b200_graph **out, char *err, size_t err_cap) {
    if (err && err_cap) err[0] = 0;
    if (!text || !out) return OSCEN_B200_E_ARG;
    *out = nullptr;
    oscen_b200_graph *g = new oscen_b200_graph();
    std::string msg;
    bool ok = dsl::compile(text, g->g, msg, graph_name);
    if (ok && !plan::compile(g->g, g->p, msg)) { ok = false; msg = "0:0: " + msg; }
    if (!ok) {
        if (err && err_cap) { strncpy(err, msg.c_str(), err_cap - 1); err[err_cap - 1] = 0; }
        delete g;
        return OSCEN_B200_E_ARG;
    }
    *out = g;
    return 0;
}
void oscen_b200_graph_free(oscen_b200_graph *g) { delete g; }
int oscen_b200_graph_describe(const oscen_b200_graph *g, char *dst, size_t cap) {
    if (!g || !dst || cap == 0) return OSCEN_B200_E_ARG;
    const std::string s = dsl::describe(g->g, g->p);
    if (s.size() + 1 > cap) return OSCEN_B200_E_ARG;
    memcpy(dst, s.c_str(), s.size() + 1);
    return (int)g->p.launches.size();
}
uint32_t oscen_b200_graph_voices_per_instance(const oscen_b200_graph *g) { return g ? (g->g.vpi ? g->g.vpi : 1u) : 0u; }
uint32_t oscen_b200_graph_channels(const oscen_b200_graph *g) { return g ? g->g.channels : 0u; }
int oscen_b200_graph_has_midi(const oscen_b200_graph *g) { return g && g->g.midi ? 1 : 0; }
uint32_t oscen_b200_graph_n_inputs(const oscen_b200_graph *g) { return g ? (uint32_t)g->g.inputs.size() : 0u; }
uint32_t oscen_b200_graph_n_assets(const oscen_b200_graph *g) { return g ? (uint32_t)g->g.assets.size() : 0u; }
int oscen_b200_graph_asset_id(const oscen_b200_graph *g, const char *name) {
    if (!g || !name) return OSCEN_B200_E_ARG;
    for (size_t i = 0; i < g->g.assets.size(); ++i) if (g->g.assets[i] == name) return (int)i;
    return OSCEN_B200_E_ARG;
}
int oscen_b200_graph_input_id(const oscen_b200_graph *g, const char *name) {
    if (!g || !name) return OSCEN_B200_E_ARG;
    for (const auto &al : g->g.aliases) if (al.first == name) return (int)al.second;      /* the outer graph's names first */
    for (size_t i = 0; i < g->g.inputs.size(); ++i) if (g->g.inputs[i].name == name) return (int)i;
    return OSCEN_B200_E_ARG;
}
int oscen_b200_graph_input_info(const oscen_b200_graph *g, uint32_t id, oscen_b200_input_info *out) {
    if (!g || !out || id >= g->g.inputs.size()) return OSCEN_B200_E_ARG;
    const plan::ValueInput &v = g->g.inputs[id];
    memset(out, 0, sizeof *out);
    strncpy(out->name, v.name.c_str(), sizeof(out->name) - 1);
    out->default_value = v.def; out->has_range = v.has_range ? 1u : 0u; out->range_min = v.lo; out->range_max = v.hi;
    out->logarithmic = v.log ? 1u : 0u; out->ramp_frames = v.ramp;
    return 0;
}
/* the CUDA source codegen.hpp generates for PLAN_FUSED (inspection / tests); returns its length */
int oscen_b200_graph_codegen(const oscen_b200_graph *g, char *dst, size_t cap) {
    if (!g || !dst || cap == 0) return OSCEN_B200_E_ARG;
    const std::string s = codegen::fused_kernel_source(g->g, g->p, kJitKernelName);
    if (s.size() + 1 > cap) return OSCEN_B200_E_ARG;
    memcpy(dst, s.c_str(), s.size() + 1);
    return (int)s.size();
}
/* compile the generated kernel for sm_100a with NVRTC without loading it (needs no GPU); returns the
   cubin size in bytes, or OSCEN_B200_E_JIT with the compiler log in `log` */
int oscen_b200_graph_jit_check(const oscen_b200_graph *g, char *log, size_t log_cap) {
    if (log && log_cap) log[0] = 0;
    if (!g) return OSCEN_B200_E_ARG;
    std::vector<char> cubin; std::string l, err;
    const bool ok = jit::compile_cubin(codegen::fused_kernel_source(g->g, g->p, kJitKernelName), cubin, l, err);
    const std::string &msg = ok ? l : err;
    if (log && log_cap) { strncpy(log, msg.c_str(), log_cap - 1); log[log_cap - 1] = 0; }
    return ok ? (int)cubin.size() : OSCEN_B200_E_JIT;
}
/* the source oscen_b200_specialize(mask) would compile for an engine at this sample rate (inspection / host-executed tests) */
int oscen_b200_graph_kernel_source(const oscen_b200_graph *g, float sample_rate, uint32_t uniform_mask, char *dst, size_t cap) {
    if (!g || !(sample_rate > 0.0f)) return OSCEN_B200_E_ARG;
    plan::Graph pg = g->g;
    plan::assign_state_layout(pg, sample_rate);
    const std::string s = codegen::fused_kernel_source(pg, g->p, kJitKernelName, uniform_mask);
    if (dst) { if (s.size() + 1 > cap) return OSCEN_B200_E_ARG; memcpy(dst, s.c_str(), s.size() + 1); }
    return (int)s.size();
}
int oscen_b200_graph_kernel_info(const oscen_b200_graph *g, float sample_rate, oscen_b200_kernel_info *info,
                                 char *source, size_t source_cap, uint32_t *state_words, size_t state_cap) {
    if (!g || !(sample_rate > 0.0f)) return OSCEN_B200_E_ARG;
    plan::Graph pg = g->g;                                   /* the same steps as create_impl, up to the device */
    plan::assign_state_layout(pg, sample_rate);
    const std::string src = codegen::fused_kernel_source(pg, g->p, kJitKernelName);
    uint32_t n_state = pg.n_state + (pg.stream_input.empty() ? 0u : 2u);
    if (n_state == 0) n_state = 1;
    const std::vector<uint32_t> init = initial_state(OSCEN_B200_GRAPH_CUSTOM, sample_rate, &pg, n_state);
    if (info) {
        info->source_len = (uint32_t)src.size(); info->n_state = n_state;
        info->dyn_smem_bytes = (uint32_t)codegen::dyn_smem_bytes(pg); info->inner_factor = 1u << pg.inner_log2;
    }
    int rc = 0;
    if (source) { if (src.size() + 1 > source_cap) rc = OSCEN_B200_E_ARG; else memcpy(source, src.c_str(), src.size() + 1); }
    if (state_words) { if (init.size() > state_cap) rc = OSCEN_B200_E_ARG; else memcpy(state_words, init.data(), init.size() * sizeof(uint32_t)); }
    return rc;
}
/* compile and load the variant of the generated kernel for the value inputs that hold one value for every voice now */
int oscen_b200_specialize(oscen_b200_engine *e, uint32_t *uniform_mask) {
    if (uniform_mask) *uniform_mask = 0;
    if (!e) return OSCEN_B200_E_ARG;
    if (e->cfg.graph_kind != OSCEN_B200_GRAPH_CUSTOM || e->cfg.plan != OSCEN_B200_PLAN_FUSED) {
        e->err = "specialize: only a PLAN_FUSED engine of a compiled graph has a generated kernel (the built-in FMVoice picks its variant itself)";
        return OSCEN_B200_E_STATE;
    }
    CK(e, cudaSetDevice(e->cfg.device));
    uint32_t mask = 0;
    for (uint32_t i = 0; i < e->n_inputs && i < 32u; ++i)
        if (e->uniform[i] && !(e->midi_ok && i == e->midi_freq_input)) mask |= 1u << i;
    for (const Event &x : e->pending) if (x.input_id != kEventGate && x.input_id < 32u) mask &= ~(1u << x.input_id);
    if (mask == e->jit_uni_mask && (mask == 0u || e->jit_uni.function)) { if (uniform_mask) *uniform_mask = mask; return 0; }
    CK(e, cudaStreamSynchronize(e->stream));                   /* the old variant may still be running */
    e->jit_uni.unload(); e->jit_uni = jit::Kernel(); e->jit_uni_mask = 0;
    if (mask != 0u) {
        jit::Kernel jk;
        jk.source = codegen::fused_kernel_source(e->graph, e->plan, kJitKernelName, mask);
        jk.dyn_smem = (unsigned)codegen::dyn_smem_bytes(e->graph);      /* decided with the source, like create */
        std::string jerr;
        if (!jit::compile_cubin(jk.source, jk.cubin, jk.log, jerr)) { e->err = "specialize (generated kernel): " + jerr; return OSCEN_B200_E_JIT; }
        if (!jk.load(kJitKernelName, jerr)) { e->err = "specialize (generated kernel): " + jerr; return OSCEN_B200_E_JIT; }
        e->jit_uni = std::move(jk);
        e->jit_uni_mask = mask;
        char nm[96]; snprintf(nm, sizeof nm, "%s<uniform=0x%x>", kJitKernelName, mask);
        e->jit_uni_name = nm;
    }
    if (uniform_mask) *uniform_mask = mask;
    return 0;
}
int oscen_b200_create_from_graph(const oscen_b200_graph *g, const oscen_b200_config *cfg, oscen_b200_engine **out) {
    if (!g) { g_create_error = "null graph"; return OSCEN_B200_E_ARG; }
    return create_impl(cfg, g, out);
}

int oscen_b200_set_value(oscen_b200_engine *e, uint32_t id, uint32_t first, uint32_t count, const float *v, int broadcast) {
    if (!e || !v) return OSCEN_B200_E_ARG;
    if (id >= e->n_inputs || (uint64_t)first + count > e->cfg.n_voices || count == 0) return arg_error(e, "set_value: id/range");
    CK(e, cudaSetDevice(e->cfg.device));
    const bool all = broadcast && first == 0 && count == e->cfg.n_voices;
    if (e->ramped[id] && !all) { e->err = "set_value: a ramped input is one ValueRampState (uniform)"; return OSCEN_B200_E_STATE; }
    if (broadcast) {
        int r = fill_row(e, id, first, count, v[0]); if (r) return r;
    } else {
        /* the copy must complete before the caller's buffer goes away */
        CK(e, cudaMemcpyAsync(e->d_pv + (size_t)id * e->cfg.n_voices + first, v, (size_t)count * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        CK(e, cudaStreamSynchronize(e->stream));
        e->stats.h2d_bytes += (uint64_t)count * sizeof(float);
    }
    if (all) { e->uni[id] = v[0]; e->uniform[id] = true; e->ramps[id].set_immediate(v[0]); }
    else e->uniform[id] = false;
    return 0;
}

/* set_<v>_with_ramp: no-op when the target is unchanged (codegen/mod.rs:948-957) */
int oscen_b200_set_value_ramp(oscen_b200_engine *e, uint32_t id, float target, uint32_t frames) {
    if (!e) return OSCEN_B200_E_ARG;
    if (id >= e->n_inputs) return arg_error(e, "set_value_ramp: id");
    if (!e->uniform[id]) { e->err = "set_value_ramp: input is not uniform"; return OSCEN_B200_E_STATE; }
    for (const Event &p : e->pending) if (p.input_id == id) {     /* the mirror image of push_events' refusal */
        e->err = "set_value_ramp: value events are pending on this input: a ramped input is one ValueRampState (uniform)"; return OSCEN_B200_E_STATE;
    }
    e->ramped[id] = true;
    if (target != e->ramps[id].target) {
        e->ramps[id].set_with_ramp(target, frames);
        if (frames == 0) {                      /* set_immediate */
            CK(e, cudaSetDevice(e->cfg.device));
            e->uni[id] = target;
            return fill_row(e, id, 0, e->cfg.n_voices, target);
        }
    }
    return 0;
}

int oscen_b200_push_events(oscen_b200_engine *e, const oscen_b200_event *ev, size_t n) {
    if (!e || (!ev && n)) return OSCEN_B200_E_ARG;
    for (size_t i = 0; i < n; ++i) {
        if (ev[i].voice >= e->cfg.n_voices) return arg_error(e, "push_events: voice out of range");
        if (ev[i].input_id != kEventGate && ev[i].input_id >= e->n_inputs) return arg_error(e, "push_events: input id");
        /* a ramped input is ONE ValueRampState for the whole bank, read through .current every frame (emit_node.rs:317-321):
           a per-voice write to it is as undefined as a per-voice set_value on it (refused there too) */
        /* ElectricPiano: vibrato_intensity / vibrato_speed feed the INSTANCE's Tremolo (main.rs:88-90), not a voice: a per-voice
           event has no voice to land on (the kernel reads them once per output tile, from the instance's first voice) */
        if (e->cfg.graph_kind == OSCEN_B200_GRAPH_EPIANO && (ev[i].input_id == 7u || ev[i].input_id == 8u))
            return arg_error(e, "push_events: vibrato_intensity / vibrato_speed are instance-level inputs (the graph's Tremolo): use set_value / set_value_ramp");
        if (ev[i].input_id != kEventGate && e->ramped[ev[i].input_id]) {
            e->err = "push_events: a value event on a ramped input (set_value_ramp was used on it): a ramped input is one ValueRampState (uniform)";
            return OSCEN_B200_E_STATE;
        }
    }
    for (size_t i = 0; i < n; ++i) {
        e->pending.push_back(Event{ev[i].voice, ev[i].frame_offset, ev[i].input_id, ev[i].scalar});
        e->seqs.push_back(e->seq++);
    }
    return 0;
}

int oscen_b200_push_midi(oscen_b200_engine *e, uint32_t instance, uint32_t frame_offset, const uint8_t *bytes, uint32_t len) {
    if (!e || !bytes) return OSCEN_B200_E_ARG;
    if (instance >= e->n_inst) return arg_error(e, "push_midi: instance out of range");
    if (!e->midi_ok) return arg_error(e, "push_midi: this graph declares no MIDI front (midi_in -> MidiParser -> VoiceAllocator -> [MidiVoiceHandler; N])");
    if (e->cfg.voices_per_instance > (uint32_t)kMaxPolyVoices) return arg_error(e, "push_midi: VoiceAllocator holds at most 24 voices (voice_allocator.rs:5)");
    if (e->midi_mode == 2) { e->err = "push_midi: this engine runs the device-side MIDI front (push_midi_batch)"; return OSCEN_B200_E_STATE; }
    if (e->midi_mode == 0) { e->midi.assign(e->n_inst, MidiFront((int)e->cfg.voices_per_instance)); e->midi_mode = 1; }
    if (!e->midi[instance].push(frame_offset, bytes, len, e->seq++)) e->stats.events_dropped++;
    return 0;
}

namespace {
/* device front: state after VoiceAllocator::new / MidiVoiceHandler::new (voice_allocator.rs:44-55, midi.rs:52-66) */
int midi_device_init(oscen_b200_engine *e) {
    const uint32_t n = e->cfg.n_voices;
    CK(e, cudaMalloc(&e->d_alloc_word, (size_t)n * 4)); CK(e, cudaMalloc(&e->d_alloc_age, (size_t)n * 4));
    CK(e, cudaMalloc(&e->d_alloc_cur_age, (size_t)e->n_inst * 4));
    CK(e, cudaMalloc(&e->d_h_note, (size_t)n * 4)); CK(e, cudaMalloc(&e->d_h_freq, (size_t)n * 4));
    CK(e, cudaMalloc(&e->d_ev_count, (size_t)n * 4)); CK(e, cudaMalloc(&e->d_ev_offset, (size_t)n * 4));
    CK(e, cudaMalloc(&e->d_note_freq, 256 * sizeof(float))); CK(e, cudaMalloc(&e->d_midi_dropped, 4));
    CK(e, cudaMemsetAsync(e->d_alloc_word, 0, (size_t)n * 4, e->stream));
    CK(e, cudaMemsetAsync(e->d_alloc_age, 0, (size_t)n * 4, e->stream));
    CK(e, cudaMemsetAsync(e->d_alloc_cur_age, 0, (size_t)e->n_inst * 4, e->stream));
    CK(e, cudaMemsetAsync(e->d_h_note, 0xFF, (size_t)n * 4, e->stream));               /* current_note = -1 */
    CK(e, cudaMemsetAsync(e->d_midi_dropped, 0, 4, e->stream));
    uint32_t bits440; const float f440 = 440.0f; memcpy(&bits440, &f440, 4);
    fill_u32_kernel<<<std::min<uint32_t>((n + 255) / 256, 1184), 256, 0, e->stream>>>((uint32_t *)e->d_h_freq, n, bits440);
    e->stats.kernel_launches++;
    float tab[256];
    for (int i = 0; i < 256; ++i) tab[i] = midi_note_to_freq((uint8_t)i);               /* host libm powf, as the host front */
    CK(e, cudaMemcpyAsync(e->d_note_freq, tab, sizeof tab, cudaMemcpyHostToDevice, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    return 0;
}

/* run the device front for this block: fills d_events (sorted by voice, frame, order) and BlockCtl.n_events */
int midi_device_run(oscen_b200_engine *e, uint32_t frames) {
    const uint32_t n = e->cfg.n_voices, nm = (uint32_t)e->midi_batch.size();
    if (nm > e->midi_cap) {
        CK(e, cudaStreamSynchronize(e->stream));
        cudaFree(e->d_midi_msgs); cudaFree(e->d_midi_keys_in); cudaFree(e->d_midi_keys); cudaFree(e->d_midi_idx_in); cudaFree(e->d_midi_idx);
        cudaFree(e->d_cub_tmp); e->d_cub_tmp = nullptr;
        e->midi_cap = (size_t)nm * 2;
        CK(e, cudaMalloc(&e->d_midi_msgs, e->midi_cap * sizeof(MidiMsg)));
        CK(e, cudaMalloc(&e->d_midi_keys_in, e->midi_cap * 4)); CK(e, cudaMalloc(&e->d_midi_keys, e->midi_cap * 4));
        CK(e, cudaMalloc(&e->d_midi_idx_in, e->midi_cap * 4)); CK(e, cudaMalloc(&e->d_midi_idx, e->midi_cap * 4));
        size_t t_sort = 0, t_scan = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, t_sort, e->d_midi_keys_in, e->d_midi_keys, e->d_midi_idx_in, e->d_midi_idx, (int)e->midi_cap, 0, 32, e->stream);
        cub::DeviceScan::ExclusiveSum(nullptr, t_scan, e->d_ev_count, e->d_ev_offset, (int)n, e->stream);
        e->cub_tmp_bytes = std::max(t_sort, t_scan);
        CK(e, cudaMalloc(&e->d_cub_tmp, e->cub_tmp_bytes));
    }
    /* every message yields at most a gate and a frequency event */
    if ((size_t)nm * 2 > e->events_cap) {
        CK(e, cudaStreamSynchronize(e->stream));
        cudaFree(e->d_events);
        e->events_cap = (size_t)nm * 4;
        CK(e, cudaMalloc(&e->d_events, e->events_cap * sizeof(Event)));
        e->drop_variants();
    }
    /* pageable source: the copy is staged by the runtime before the call returns */
    CK(e, cudaMemcpyAsync(e->d_midi_msgs, e->midi_batch.data(), (size_t)nm * sizeof(MidiMsg), cudaMemcpyHostToDevice, e->stream));
    e->stats.h2d_bytes += (uint64_t)nm * sizeof(MidiMsg);
    midi_keys_kernel<<<(nm + 255) / 256, 256, 0, e->stream>>>(e->d_midi_msgs, nm, e->d_midi_keys_in, e->d_midi_idx_in);
    int bits = 1; while ((1ull << bits) < (unsigned long long)e->n_inst) ++bits;
    size_t tmp = e->cub_tmp_bytes;
    CK(e, cub::DeviceRadixSort::SortPairs(e->d_cub_tmp, tmp, e->d_midi_keys_in, e->d_midi_keys, e->d_midi_idx_in, e->d_midi_idx, (int)nm, 0, bits, e->stream));
    CK(e, cudaMemsetAsync(e->d_ev_count, 0, (size_t)n * 4, e->stream));
    MidiFrontArgs a{};
    a.msgs = e->d_midi_msgs; a.keys = e->d_midi_keys; a.idx = e->d_midi_idx;
    a.n_msgs = nm; a.n_inst = e->n_inst; a.vpi = e->cfg.voices_per_instance; a.frames = frames;
    a.alloc_word = e->d_alloc_word; a.alloc_age = e->d_alloc_age; a.alloc_cur_age = e->d_alloc_cur_age;
    a.h_note = e->d_h_note; a.h_freq = e->d_h_freq; a.note_freq = e->d_note_freq;
    a.ev_count = e->d_ev_count; a.ev_offset = e->d_ev_offset; a.events = e->d_events; a.dropped = e->d_midi_dropped;
    a.freq_input = e->midi_freq_input;
    const uint32_t grid = (e->n_inst + 127) / 128;
    midi_front_kernel<false><<<grid, 128, 0, e->stream>>>(a);
    tmp = e->cub_tmp_bytes;
    CK(e, cub::DeviceScan::ExclusiveSum(e->d_cub_tmp, tmp, e->d_ev_count, e->d_ev_offset, (int)n, e->stream));
    midi_front_kernel<true><<<grid, 128, 0, e->stream>>>(a);
    midi_finish_kernel<<<1, 1, 0, e->stream>>>(e->d_ev_count, e->d_ev_offset, n, e->d_ctl);
    e->stats.kernel_launches += 4;
    CK(e, cudaGetLastError());
    e->midi_batch.clear();
    return 0;
}
}  // namespace

int oscen_b200_push_midi_batch(oscen_b200_engine *e, const oscen_b200_midi_msg *msgs, size_t n) {
    static_assert(sizeof(MidiMsg) == sizeof(oscen_b200_midi_msg), "midi message layout");
    if (!e || (!msgs && n)) return OSCEN_B200_E_ARG;
    if (!e->midi_ok) return arg_error(e, "push_midi_batch: this graph declares no MIDI front (midi_in -> MidiParser -> VoiceAllocator -> [MidiVoiceHandler; N])");
    if (e->cfg.voices_per_instance > (uint32_t)kDevMaxPoly) return arg_error(e, "push_midi_batch: voices_per_instance must be <= 16");
    if (e->midi_mode == 1) { e->err = "push_midi_batch: this engine runs the host-side MIDI front (push_midi)"; return OSCEN_B200_E_STATE; }
    for (size_t i = 0; i < n; ++i) if (msgs[i].instance >= e->n_inst) return arg_error(e, "push_midi_batch: instance out of range");
    CK(e, cudaSetDevice(e->cfg.device));
    if (e->midi_mode == 0) { const int r = midi_device_init(e); if (r) return r; e->midi_mode = 2; }
    const size_t at = e->midi_batch.size();
    e->midi_batch.resize(at + n);
    for (size_t i = 0; i < n; ++i) {
        MidiMsg m{};
        m.instance = msgs[i].instance; m.frame = msgs[i].frame_offset;
        m.len = msgs[i].len > 3 ? 3 : msgs[i].len;                                     /* RawMidiMessage::new, midi.rs:15-22 */
        for (uint32_t k = 0; k < m.len; ++k) m.bytes[k] = msgs[i].bytes[k];
        e->midi_batch[at + i] = m;
    }
    return 0;
}

int oscen_b200_stream_in(oscen_b200_engine *e, uint32_t id, const float *src, uint32_t frames, int is_device) {
    if (!e || !src) return OSCEN_B200_E_ARG;
    if (!e->has_stream_in || id != 0 || frames == 0 || frames > kMaxBlock)
        return arg_error(e, "stream_in: only graphs with a stream input (TWIN_PEAKS audio_in, CLIP_* input; id 0), frames <= 512");
    CK(e, cudaSetDevice(e->cfg.device));
    const size_t nfl = (size_t)frames * e->cfg.n_voices;
    if (is_device) {
        e->stream_in_ptr = src;
    } else {
        if (e->stream_in_cap < nfl) {
            CK(e, cudaStreamSynchronize(e->stream));
            cudaFree(e->d_stream_in); e->d_stream_in = nullptr;
            CK(e, cudaMalloc(&e->d_stream_in, nfl * sizeof(float)));
            e->stream_in_cap = nfl;
        }
        CK(e, cudaMemcpyAsync(e->d_stream_in, src, nfl * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        CK(e, cudaStreamSynchronize(e->stream));
        e->stats.h2d_bytes += nfl * sizeof(float);
        e->stream_in_ptr = e->d_stream_in;
    }
    e->stream_in_frames = frames;
    e->lcg_on = false;
    return 0;
}

__global__ void lcg_seed_kernel(float *state, uint32_t n, unsigned long long seed, uint32_t voice_offset, uint32_t field) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const unsigned long long s = lcg_seed_state(seed + (unsigned long long)(voice_offset + v));
    state[(size_t)(field + 0) * n + v] = __uint_as_float((uint32_t)(s & 0xffffffffull));
    state[(size_t)(field + 1) * n + v] = __uint_as_float((uint32_t)(s >> 32));
}

int oscen_b200_stream_in_lcg(oscen_b200_engine *e, uint32_t id, uint64_t seed, uint32_t voice_offset) {
    if (!e) return OSCEN_B200_E_ARG;
    if (!e->has_stream_in || id != 0) return arg_error(e, "stream_in_lcg: this graph has no stream input");
    CK(e, cudaSetDevice(e->cfg.device));
    e->lcg_seed = seed; e->voice_offset = voice_offset; e->lcg_on = true;
    lcg_seed_kernel<<<(e->cfg.n_voices + 255) / 256, 256, 0, e->stream>>>(e->d_state, e->cfg.n_voices, seed, voice_offset, e->lcg_field);
    e->stats.kernel_launches++;
    CK(e, cudaGetLastError());
    return 0;
}

int oscen_b200_process_block(oscen_b200_engine *e, uint32_t frames) {
    if (!e) return OSCEN_B200_E_ARG;
    if (frames == 0 || frames > kMaxBlock) return arg_error(e, "process_block: frames must be 1..512");
    const uint32_t kind = e->cfg.graph_kind, n = e->cfg.n_voices, vpi = e->cfg.voices_per_instance;
    if (e->has_stream_in && !e->lcg_on && e->stream_in_ptr && e->stream_in_frames < frames)
        return arg_error(e, "process_block: stream_in block shorter than frames");
    CK(e, cudaSetDevice(e->cfg.device));

    const bool device_midi = e->midi_mode == 2 && !e->midi_batch.empty();
    if (device_midi && !e->pending.empty()) {
        e->err = "process_block: push_events and push_midi_batch in the same block are not supported";
        return OSCEN_B200_E_STATE;
    }
    /* ---- MIDI front -> per-voice events */
    std::vector<VoiceEvent> ve;
    for (uint32_t i = 0; i < (e->midi_mode == 1 ? e->n_inst : 0u); ++i) {
        if (e->midi[i].pending.empty()) continue;
        ve.clear();
        e->midi[i].run(frames, ve);
        for (const VoiceEvent &x : ve) {
            e->pending.push_back(Event{i * vpi + x.voice_in_instance, x.frame, x.is_gate ? kEventGate : e->midi_freq_input, x.value});
            e->seqs.push_back(e->seq++);
        }
    }
    /* ---- stage events.  Reference semantics: every pushed gate event takes a slot of the voice's
       32-slot queue in PUSH order (try_push on a full ArrayVec drops, graph/types.rs:183-188), then
       process_block stages the queue sorted by frame_offset and never reaches offsets >= frames
       (codegen/mod.rs:836-872).  Value events are field assignments: no cap. */
    const size_t n_pend = e->pending.size();
    /* already in (voice, frame) order, as a host that walks its voices produces it?  then arrival order is the tie-break
       and the 65 536-event burst of a bank-wide note-on needs no sort */
    bool sorted = true;
    for (size_t i = 1; i < n_pend && sorted; ++i) {
        const Event &a = e->pending[i - 1], &b = e->pending[i];
        sorted = a.voice < b.voice || (a.voice == b.voice && a.frame <= b.frame);
    }
    std::vector<uint32_t> order;
    if (!sorted) {
        order.resize(n_pend);
        for (uint32_t i = 0; i < order.size(); ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
            const Event &a = e->pending[x], &b = e->pending[y];
            if (a.voice != b.voice) return a.voice < b.voice;
            if (a.frame != b.frame) return a.frame < b.frame;
            return e->seqs[x] < e->seqs[y];
        });
    }
    auto at = [&](size_t k) -> uint32_t { return sorted ? (uint32_t)k : order[k]; };
    uint32_t n_ev = 0;
    bool general = false, has_value_events = false;
    Event *h_events = nullptr;
    int ev_slot = -1;
    if (n_pend) {
        if (n_pend > e->events_cap) {
            CK(e, cudaStreamSynchronize(e->stream));
            cudaFree(e->d_events); e->d_events = nullptr;
            e->events_cap = n_pend * 2;
            CK(e, cudaMalloc(&e->d_events, e->events_cap * sizeof(Event)));
            e->drop_variants();
        }
        { void *hp = nullptr; const int r = stage_acquire(e, n_pend * sizeof(Event), &hp, &ev_slot); if (r) return r; h_events = (Event *)hp; }
        std::vector<uint32_t> gate_seqs;
        size_t g0 = 0;
        while (g0 < n_pend) {
            size_t g1 = g0;
            const uint32_t voice = e->pending[at(g0)].voice;
            uint32_t n_gates = 0;
            while (g1 < n_pend && e->pending[at(g1)].voice == voice) { if (e->pending[at(g1)].input_id == kEventGate) ++n_gates; ++g1; }
            uint32_t seq_limit = 0xFFFFFFFFu;                  /* gates pushed at or after this seq found the queue full */
            if (n_gates > OSCEN_B200_MAX_EVENTS_PER_ENDPOINT) {
                gate_seqs.clear();
                for (size_t k = g0; k < g1; ++k) if (e->pending[at(k)].input_id == kEventGate) gate_seqs.push_back(e->seqs[at(k)]);
                std::nth_element(gate_seqs.begin(), gate_seqs.begin() + OSCEN_B200_MAX_EVENTS_PER_ENDPOINT, gate_seqs.end());
                seq_limit = gate_seqs[OSCEN_B200_MAX_EVENTS_PER_ENDPOINT];
            }
            for (size_t k = g0; k < g1; ++k) {
                const uint32_t idx = at(k);
                const Event &x = e->pending[idx];
                if (x.input_id == kEventGate && e->seqs[idx] >= seq_limit) { e->stats.events_dropped++; continue; }
                if (x.frame >= frames) { e->stats.events_dropped++; continue; }
                if (x.input_id != kEventGate) {
                    e->uniform[x.input_id] = false; has_value_events = true;
                    if (x.input_id != 0u || kind != OSCEN_B200_GRAPH_FM_VOICE) general = true;
                }
                h_events[n_ev++] = x;
            }
            g0 = g1;
        }
    }
    e->pending.clear(); e->seqs.clear();

    /* ---- ramps: tick_ramps() once per frame (codegen/mod.rs:878-914) */
    uint32_t ramp_mask = 0;
    for (uint32_t i = 0; i < e->n_inputs; ++i)
        if (e->ramped[i] && e->ramps[i].remaining > 0) ramp_mask |= 1u << i;
    float *h_ptab = nullptr;
    int ptab_slot = -1;
    if (ramp_mask) {
        { void *hp = nullptr; const int r = stage_acquire(e, (size_t)frames * (kMaxInputs + kFmRowFloats) * sizeof(float), &hp, &ptab_slot); if (r) return r; h_ptab = (float *)hp; }
        for (uint32_t f = 0; f < frames; ++f)
            for (uint32_t i = 0; i < e->n_inputs; ++i) if ((ramp_mask >> i) & 1u) {
                e->ramps[i].tick();
                h_ptab[(size_t)f * kMaxInputs + i] = e->ramps[i].current;
            }
    }
    if (device_midi) { e->uniform[e->midi_freq_input] = false; has_value_events = true; }               /* per-voice `frequency` value events */
    for (uint32_t i = 1; i < e->n_inputs; ++i) if (!e->uniform[i]) general = true;
    if (kind != OSCEN_B200_GRAPH_FM_VOICE) general = true;
    /* FMVoice: which kernel renders this block (fm_select.hpp), then the ramp table it reads */
    bool upload_ptab = ramp_mask != 0;
    if (kind == OSCEN_B200_GRAPH_FM_VOICE && e->cfg.plan == OSCEN_B200_PLAN_FUSED) {
        e->fm_variant = fm_select_variant(general, ramp_mask, e->uni, h_ptab, frames, e->fm_bounds);
        if (e->fm_variant == FM_SEG_RAMP) {
            float *rows = h_ptab + (size_t)frames * kMaxInputs;
            fm_fill_rows(ramp_mask, e->uni, h_ptab, frames, rows);
            CK(e, cudaMemcpyAsync(e->d_utab, rows, (size_t)frames * kFmRowFloats * sizeof(float), cudaMemcpyHostToDevice, e->stream));
            e->stats.h2d_bytes += (uint64_t)frames * kFmRowFloats * sizeof(float);
            upload_ptab = false;
        }
    }
    if (upload_ptab) {
        CK(e, cudaMemcpyAsync(e->d_ptab, h_ptab, (size_t)frames * kMaxInputs * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        e->stats.h2d_bytes += (uint64_t)frames * kMaxInputs * sizeof(float);
    }
    if (ptab_slot >= 0) { const int r = stage_submit(e, ptab_slot); if (r) return r; }
    if (ramp_mask) general = true;

    /* ---- per-block control + events to the device */
    e->h_ctl->frames = frames; e->h_ctl->n_events = n_ev; e->h_ctl->ramp_mask = ramp_mask; e->h_ctl->lcg_on = e->lcg_on ? 1u : 0u;
    if (n_ev) {
        CK(e, cudaMemcpyAsync(e->d_events, h_events, (size_t)n_ev * sizeof(Event), cudaMemcpyHostToDevice, e->stream));
        e->stats.h2d_bytes += (uint64_t)n_ev * sizeof(Event);
    }
    if (ev_slot >= 0) { const int r = stage_submit(e, ev_slot); if (r) return r; }
    if (memcmp(&e->ctl_on_device, e->h_ctl, sizeof(BlockCtl)) != 0) {
        /* BlockCtl only travels when it changed (steady-state blocks have no events / ramps) */
        e->ctl_on_device = *e->h_ctl;
        CK(e, cudaMemcpyAsync(e->d_ctl, &e->ctl_on_device, sizeof(BlockCtl), cudaMemcpyHostToDevice, e->stream));
        e->stats.h2d_bytes += sizeof(BlockCtl);
    }

    if (device_midi) {
        /* the device front writes the event list and BlockCtl.n_events itself; the host's copy of
           BlockCtl is stale afterwards, so the next block uploads it again */
        const int r = midi_device_run(e, frames); if (r) return r;
        e->ctl_on_device.n_events = 0xFFFFFFFFu;
    }
    if (e->timing) {
        const uint32_t slot = e->t_count % kTimingRing;
        if (e->tev.size() < (size_t)(slot + 1) * 3) {
            const size_t at = e->tev.size();
            e->tev.resize((size_t)(slot + 1) * 3, nullptr);
            for (size_t k = at; k < e->tev.size(); ++k) CK(e, cudaEventCreate(&e->tev[k]));
        }
        e->ev0 = e->tev[slot * 3 + 0]; e->evk = e->tev[slot * 3 + 1]; e->ev1 = e->tev[slot * 3 + 2];
        CK(e, cudaEventRecord(e->ev0, e->stream));
    }
    int rc = 0;
    if (e->cfg.plan == OSCEN_B200_PLAN_FUSED) {
        rc = launch_fused(e, general); if (rc) return rc;
        if (e->timing) CK(e, cudaEventRecord(e->evk, e->stream));
        rc = launch_mix_reduce(e, frames); if (rc) return rc;
    } else {
        /* the launch sequence is captured once (per variant) and replayed: per-block words live in BlockCtl */
        const bool no_events = n_ev == 0 && !device_midi && !no_plain_nodes();
        const uint64_t key = (uint64_t)ramp_mask | ((uint64_t)(has_value_events ? 1u : 0u) << 32) | ((uint64_t)(no_events ? 1u : 0u) << 33);
        if (e->captured_in != e->stream_in_ptr || e->captured_lcg != e->lcg_on) {
            e->drop_variants();
            e->captured_in = e->stream_in_ptr; e->captured_lcg = e->lcg_on;
        }
        oscen_b200_engine::PlanVariant *var = nullptr;
        for (auto &v : e->variants) if (v.key == key && v.exec) var = &v;
        if (!var) {
            if (e->variants.size() >= 12) {                        /* evict the least recently launched variant */
                size_t old = 0;
                for (size_t k = 1; k < e->variants.size(); ++k) if (e->variants[k].last_use < e->variants[old].last_use) old = k;
                CK(e, cudaStreamSynchronize(e->stream));           /* its last launch may still be running */
                if (e->variants[old].exec) cudaGraphExecDestroy(e->variants[old].exec);
                e->variants.erase(e->variants.begin() + (long)old);
            }
            e->variants.emplace_back();
            var = &e->variants.back();
            var->key = key;
            cudaGraph_t g = nullptr;
            { const cudaError_t be = cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal); if (be != cudaSuccess) { e->variants.pop_back(); CK(e, be); } }
            const uint64_t before = e->stats.kernel_launches;
            var->names = "node plan (CUDA graph):";
            e->plan_log = &var->names;
            rc = launch_plan(e, has_value_events, no_events, ramp_mask);
            e->plan_log = nullptr;
            if (!rc) rc = launch_mix_reduce(e, frames);
            cudaError_t s = cudaStreamEndCapture(e->stream, &g);
            var->kernels = e->stats.kernel_launches - before;
            e->stats.kernel_launches = before;
            if (rc || s != cudaSuccess) { e->variants.pop_back(); if (g) cudaGraphDestroy(g); }
            if (rc) return rc;
            CK(e, s);
            const cudaError_t ie = cudaGraphInstantiate(&var->exec, g, 0);
            cudaGraphDestroy(g);
            if (ie != cudaSuccess) { e->variants.pop_back(); CK(e, ie); }
        }
        var->last_use = ++e->variant_clock;
        CK(e, cudaGraphLaunch(var->exec, e->stream));
        e->last_kernel = var->names.c_str();                       /* valid until the variant is evicted */
        if (e->timing) CK(e, cudaEventRecord(e->evk, e->stream));
        e->stats.graph_launches++;
        e->stats.kernel_launches += var->kernels;
    }
    if (e->comm && !e->p2p) {
        const int r = g_nccl.AllReduce(e->d_mix, e->d_mix, (size_t)frames * e->channels, /*ncclFloat32*/ 7, /*ncclSum*/ 0, e->comm, e->stream);
        if (r != 0) { e->err = std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"); return OSCEN_B200_E_NCCL; }
    }
    if (e->timing) { CK(e, cudaEventRecord(e->ev1, e->stream)); e->ev_valid = true; e->t_count++; }
    /* ramp-driven inputs end the block at their last per-frame value; the per-voice row is only
       read again once the input has stopped ramping, so it is refilled when the ramp completes */
    if (ramp_mask) {
        for (uint32_t i = 0; i < e->n_inputs; ++i) if ((ramp_mask >> i) & 1u) {
            e->uni[i] = e->ramps[i].current;
            if (e->ramps[i].remaining == 0) { rc = fill_row(e, i, 0, n, e->uni[i]); if (rc) return rc; }
        }
    }
    if (e->asset_reset_pending) {
        /* the block just enqueued restarted the players of the freshly loaded slots; later blocks carry on (stream order) */
        for (uint32_t k = 0; k < kMaxAssets; ++k) e->h_assets.reset[k] = 0u;
        CK(e, cudaMemcpyAsync(e->d_assets, &e->h_assets, sizeof(AssetTable), cudaMemcpyHostToDevice, e->stream));
        e->asset_reset_pending = false;
    }
    e->last_frames = frames;
    e->stats.blocks++;
    return 0;
}

/* graph.<slot>.load(..) (oscen-lib/src/asset/mod.rs; sample_player/mod.rs:36-49,104-112): publish a new buffer into an
   `external NAME: AudioAsset;` slot.  `interleaved` holds frames x channels f32 at the engine's sample rate (decoding and
   rate conversion stay with the caller); a mono SamplePlayer takes channel 0.  Every player bound to the slot restarts at
   the next block, for every voice of the engine.  frames == 0 empties the slot (silence).  Synchronises the stream. */
int oscen_b200_load_asset(oscen_b200_engine *e, uint32_t asset_id, const float *interleaved, uint32_t frames, uint32_t channels) {
    if (!e) return OSCEN_B200_E_ARG;
    if (asset_id >= e->n_assets) return arg_error(e, "load_asset: the graph declares no such external slot");
    if (frames && (!interleaved || channels == 0)) return arg_error(e, "load_asset: null samples / zero channels");
    CK(e, cudaSetDevice(e->cfg.device));
    CK(e, cudaStreamSynchronize(e->stream));                       /* no block in flight still reads the old buffer */
    float *fresh = nullptr;
    if (frames) {
        std::vector<float> mono(frames);
        for (uint32_t t = 0; t < frames; ++t) mono[t] = interleaved[(size_t)t * channels];
        CK(e, cudaMalloc(&fresh, (size_t)frames * sizeof(float)));
        const cudaError_t c = cudaMemcpy(fresh, mono.data(), (size_t)frames * sizeof(float), cudaMemcpyHostToDevice);
        if (c != cudaSuccess) { cudaFree(fresh); CK(e, c); }
        e->stats.h2d_bytes += (uint64_t)frames * sizeof(float);
    }
    cudaFree(e->d_asset_buf[asset_id]);
    e->d_asset_buf[asset_id] = fresh;
    e->h_assets.data[asset_id] = fresh; e->h_assets.frames[asset_id] = frames; e->h_assets.reset[asset_id] = 1u;
    CK(e, cudaMemcpy(e->d_assets, &e->h_assets, sizeof(AssetTable), cudaMemcpyHostToDevice));
    e->asset_reset_pending = true;
    return 0;
}

int oscen_b200_synchronize(oscen_b200_engine *e) {
    if (!e) return OSCEN_B200_E_ARG;
    CK(e, cudaSetDevice(e->cfg.device));
    CK(e, cudaStreamSynchronize(e->stream));
    return 0;
}

int oscen_b200_read_voice_out(oscen_b200_engine *e, float *dst, uint32_t first, uint32_t count) {
    if (!e || !dst) return OSCEN_B200_E_ARG;
    const float *src = e->cfg.plan == OSCEN_B200_PLAN_NODES ? e->final_stream : e->d_voice_out;
    if (!src) { e->err = "read_voice_out: engine created without KEEP_VOICE_OUT"; return OSCEN_B200_E_STATE; }
    if ((uint64_t)first + count > e->cfg.n_voices || count == 0) return arg_error(e, "read_voice_out: range");
    CK(e, cudaSetDevice(e->cfg.device));
    CK(e, cudaMemcpy2DAsync(dst, (size_t)count * sizeof(float), src + first, (size_t)e->cfg.n_voices * sizeof(float),
                            (size_t)count * sizeof(float), e->last_frames, cudaMemcpyDeviceToHost, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    e->stats.d2h_bytes += (uint64_t)count * e->last_frames * sizeof(float);
    return 0;
}

int oscen_b200_read_instance_out(oscen_b200_engine *e, float *dst, uint32_t first, uint32_t count) {
    if (!e || !dst) return OSCEN_B200_E_ARG;
    if (!e->d_inst_out) { e->err = "read_instance_out: engine created without KEEP_INSTANCE_OUT"; return OSCEN_B200_E_STATE; }
    if ((uint64_t)first + count > e->n_inst || count == 0) return arg_error(e, "read_instance_out: range");
    CK(e, cudaSetDevice(e->cfg.device));
    const size_t ch = e->channels;
    CK(e, cudaMemcpy2DAsync(dst, (size_t)count * ch * sizeof(float), e->d_inst_out + (size_t)first * ch,
                            (size_t)e->n_inst * ch * sizeof(float), (size_t)count * ch * sizeof(float), e->last_frames,
                            cudaMemcpyDeviceToHost, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    e->stats.d2h_bytes += (uint64_t)count * ch * e->last_frames * sizeof(float);
    return 0;
}

int oscen_b200_read_mix(oscen_b200_engine *e, float *dst) {
    if (!e || !dst) return OSCEN_B200_E_ARG;
    CK(e, cudaSetDevice(e->cfg.device));
    const size_t bytes = (size_t)e->last_frames * e->channels * sizeof(float);
    CK(e, cudaMemcpyAsync(e->h_mix, e->d_mix, bytes, cudaMemcpyDeviceToHost, e->stream));
    uint32_t *h_flag = reinterpret_cast<uint32_t *>(e->h_mix + (size_t)kMaxBlock * e->channels);
    if (e->p2p) CK(e, cudaMemcpyAsync(h_flag, e->d_p2p_error, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    memcpy(dst, e->h_mix, bytes);
    e->stats.d2h_bytes += bytes;
    if (e->p2p && *h_flag != 0u) { e->err = "mix bus: a peer rank did not publish its block (bounded wait expired)"; return OSCEN_B200_E_NCCL; }
    return 0;
}

uint32_t oscen_b200_channels(const oscen_b200_engine *e) { return e ? e->channels : 0; }

/* latency_samples(): sum over Down edges of kernel latency / factor
   (codegen/emit_struct.rs:533-566); SincDownFir<4> reports 33 high-rate samples */
uint32_t oscen_b200_latency_samples(const oscen_b200_engine *e) {
    if (!e) return 0u;
    const uint32_t kind = e->cfg.graph_kind;
    if (kind == OSCEN_B200_GRAPH_SAT_4X) return 33u / 4u;
    if (kind == OSCEN_B200_GRAPH_CUSTOM) {
        /* sum over Down edges of the kernel's latency in high-rate samples / factor (codegen/emit_struct.rs:533-566) */
        uint32_t total = 0;
        for (const plan::Node &nd : e->graph.nodes) if (nd.type == plan::NodeType::Down) {
            const uint32_t factor = 1u << nd.rate_log2;
            const uint32_t lat = nd.policy == plan::kPolicySinc ? 11u * (factor - 1u) : nd.policy == plan::kPolicySincIir ? 2u * (factor - 1u)
                               : nd.policy == plan::kPolicyLinear ? (factor - 1u) / 2u : 0u;
            total += lat / factor;
        }
        return total;
    }
    if (is_clip(kind) && clip_log2f(kind) > 0) {
        const uint32_t l = (uint32_t)clip_log2f(kind), factor = 1u << l;
        switch (clip_policy(kind)) {
        case POL_SINC: return 11u * (factor - 1u) / factor;          /* sinc_fir.rs:250-260 */
        case POL_SINC_IIR: return 2u * (factor - 1u) / factor;       /* halfband_iir.rs:253-262 */
        case POL_LINEAR: return ((factor - 1u) / 2u) / factor;       /* linear.rs:67-70 */
        default: return 0u;                                          /* latch.rs:50-53 */
        }
    }
    return 0u;
}

size_t oscen_b200_state_bytes_per_voice(const oscen_b200_engine *e) { return e ? (size_t)e->n_state * sizeof(float) : 0; }

int oscen_b200_state_download(oscen_b200_engine *e, void *dst, size_t cap) {
    if (!e || !dst) return OSCEN_B200_E_ARG;
    const size_t bytes = (size_t)e->n_state * e->cfg.n_voices * sizeof(float);
    if (cap < bytes) return arg_error(e, "state_download: buffer too small");
    CK(e, cudaSetDevice(e->cfg.device));
    CK(e, cudaMemcpyAsync(dst, e->d_state, bytes, cudaMemcpyDeviceToHost, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    e->stats.d2h_bytes += bytes;
    return 0;
}

int oscen_b200_get_stats(const oscen_b200_engine *e, oscen_b200_stats *out) {
    if (!e || !out) return OSCEN_B200_E_ARG;
    *out = e->stats;
    if (e->d_midi_dropped) {                               /* messages the device front found a full queue for */
        uint32_t d = 0;
        cudaSetDevice(e->cfg.device);
        if (cudaStreamSynchronize(e->stream) == cudaSuccess &&
            cudaMemcpy(&d, e->d_midi_dropped, 4, cudaMemcpyDeviceToHost) == cudaSuccess) out->events_dropped += d;
    }
    return 0;
}

int oscen_b200_enable_timing(oscen_b200_engine *e, int on) {
    if (!e) return OSCEN_B200_E_ARG;
    e->timing = on != 0; e->ev_valid = false; e->t_count = 0; e->t_drained = 0;
    return 0;
}

/* device times of the blocks rendered since enable_timing(1) / the previous drain, oldest first: ONE synchronisation for
   the whole batch.  Returns the number of blocks written (<= cap); blocks older than the ring (4096) are lost. */
int oscen_b200_timing_drain(oscen_b200_engine *e, float *block_ms, float *kernel_ms, uint32_t cap) {
    if (!e || (!block_ms && !kernel_ms)) return OSCEN_B200_E_ARG;
    if (!e->timing || e->t_count == 0) return 0;
    CK(e, cudaSetDevice(e->cfg.device));
    CK(e, cudaEventSynchronize(e->ev1));
    if (e->t_count - e->t_drained > kTimingRing) e->t_drained = e->t_count - kTimingRing;
    uint32_t n = 0;
    while (e->t_drained < e->t_count && n < cap) {
        const uint32_t slot = e->t_drained % kTimingRing;
        float ms = 0.0f;
        if (block_ms) { CK(e, cudaEventElapsedTime(&ms, e->tev[slot * 3 + 0], e->tev[slot * 3 + 2])); block_ms[n] = ms; }
        if (kernel_ms) { CK(e, cudaEventElapsedTime(&ms, e->tev[slot * 3 + 0], e->tev[slot * 3 + 1])); kernel_ms[n] = ms; }
        ++n; e->t_drained++;
    }
    return (int)n;
}

/* name of the graph kernel (variant) the last process_block launched: static string, never NULL */
const char *oscen_b200_last_kernel_name(const oscen_b200_engine *e) { return e ? e->last_kernel : ""; }

int oscen_b200_last_block_ms(oscen_b200_engine *e, float *ms) {
    if (!e || !ms) return OSCEN_B200_E_ARG;
    if (!e->ev_valid) { e->err = "last_block_ms: timing not enabled"; return OSCEN_B200_E_STATE; }
    CK(e, cudaSetDevice(e->cfg.device));
    CK(e, cudaEventSynchronize(e->ev1));
    CK(e, cudaEventElapsedTime(ms, e->ev0, e->ev1));
    return 0;
}

int oscen_b200_last_kernel_ms(oscen_b200_engine *e, float *ms) {
    if (!e || !ms) return OSCEN_B200_E_ARG;
    if (!e->ev_valid) { e->err = "last_kernel_ms: timing not enabled"; return OSCEN_B200_E_STATE; }
    CK(e, cudaSetDevice(e->cfg.device));
    CK(e, cudaEventSynchronize(e->evk));
    CK(e, cudaEventElapsedTime(ms, e->ev0, e->evk));
    return 0;
}

int oscen_b200_l2_flush(oscen_b200_engine *e) {
    if (!e) return OSCEN_B200_E_ARG;
    CK(e, cudaSetDevice(e->cfg.device));
    const size_t words = (size_t)64 << 20;                 /* 256 MiB */
    if (!e->d_flush) CK(e, cudaMalloc(&e->d_flush, words * sizeof(uint32_t)));
    fill_u32_kernel<<<1184, 256, 0, e->stream>>>(e->d_flush, words, (uint32_t)e->stats.blocks);
    CK(e, cudaGetLastError());
    return 0;
}

int oscen_b200_plan_describe(uint32_t graph_kind, char *dst, size_t cap) {
    if (!dst || cap == 0) return OSCEN_B200_E_ARG;
    plan::Graph g; plan::Plan p; std::string err;
    if (!plan::for_kind(graph_kind, g) || !plan::compile(g, p, err)) { dst[0] = 0; return OSCEN_B200_E_ARG; }
    const std::string s = plan::describe(g, p);
    if (s.size() + 1 > cap) return OSCEN_B200_E_ARG;
    memcpy(dst, s.c_str(), s.size() + 1);
    return (int)p.launches.size();
}

int oscen_b200_comm_p2p_export(oscen_b200_engine *e, int n_ranks, void *handle64) {
    if (!e || !handle64 || n_ranks < 1 || n_ranks > (int)kMaxRanks) return OSCEN_B200_E_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == OSCEN_B200_IPC_HANDLE_BYTES, "ipc handle size");
    CK(e, cudaSetDevice(e->cfg.device));
    if (!e->d_xchg) {
        CK(e, cudaMalloc(&e->d_xchg, kXchgBytes));
        CK(e, cudaMemset(e->d_xchg, 0, kXchgBytes));
        CK(e, cudaMalloc(&e->d_p2p_error, 4 * sizeof(uint32_t)));      /* [0] error flag, [1] CTA arrival counter, [2] block sequence number */
        CK(e, cudaMemset(e->d_p2p_error, 0, 4 * sizeof(uint32_t)));
    }
    cudaIpcMemHandle_t h;
    CK(e, cudaIpcGetMemHandle(&h, e->d_xchg));
    memcpy(handle64, &h, sizeof h);
    return 0;
}

int oscen_b200_comm_p2p_import(oscen_b200_engine *e, int rank, int n_ranks, const void *handles) {
    if (!e || !handles || n_ranks < 1 || n_ranks > (int)kMaxRanks || rank < 0 || rank >= n_ranks) return OSCEN_B200_E_ARG;
    if (!e->d_xchg) { e->err = "comm_p2p_import: call comm_p2p_export first"; return OSCEN_B200_E_STATE; }
    CK(e, cudaSetDevice(e->cfg.device));
    CK(e, cudaStreamSynchronize(e->stream));
    /* a fresh exchange: sequence numbers restart at 1, so no flag or datum of an earlier import may survive, and a
       timeout recorded against an earlier peer set is forgotten.  The caller must put a barrier between the ranks'
       imports and their first process_block (a peer's first flag store must not race this memset). */
    CK(e, cudaMemset(e->d_xchg, 0, kXchgBytes));
    CK(e, cudaMemset(e->d_p2p_error, 0, 4 * sizeof(uint32_t)));
    for (int r = 0; r < n_ranks; ++r) {
        if (r == rank) { e->peers.peer[r] = e->d_xchg; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + (size_t)r * sizeof h, sizeof h);
        void *p = nullptr;
        CK(e, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        e->peers.peer[r] = (float *)p;
    }
    e->rank = rank; e->n_ranks = n_ranks; e->p2p = true;
    e->drop_variants();
    return 0;
}

int oscen_b200_comm_p2p_close(oscen_b200_engine *e) {
    if (!e) return OSCEN_B200_E_ARG;
    CK(e, cudaSetDevice(e->cfg.device));
    CK(e, cudaStreamSynchronize(e->stream));
    for (int r = 0; r < (int)kMaxRanks; ++r) {
        if (e->p2p && r != e->rank && e->peers.peer[r]) cudaIpcCloseMemHandle(e->peers.peer[r]);
        e->peers.peer[r] = nullptr;
    }
    e->p2p = false;
    if (e->d_p2p_error) CK(e, cudaMemset(e->d_p2p_error, 0, 4 * sizeof(uint32_t)));
    e->drop_variants();
    return 0;
}

int oscen_b200_comm_unique_id(void *id128) {
    std::string err;
    if (!id128) return OSCEN_B200_E_ARG;
    if (!nccl_load(err)) { g_create_error = err; return OSCEN_B200_E_NCCL; }
    const int r = g_nccl.GetUniqueId(id128);
    if (r != 0) { g_create_error = "ncclGetUniqueId failed"; return OSCEN_B200_E_NCCL; }
    return 0;
}

int oscen_b200_comm_init(oscen_b200_engine *e, int rank, int n_ranks, const void *id128) {
    if (!e || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return OSCEN_B200_E_ARG;
    if (!nccl_load(e->err)) return OSCEN_B200_E_NCCL;
    CK(e, cudaSetDevice(e->cfg.device));
    Id128 id; memcpy(&id, id128, sizeof id);
    const int r = g_nccl.CommInitRank(&e->comm, n_ranks, id, rank);
    if (r != 0) { e->err = std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"); e->comm = nullptr; return OSCEN_B200_E_NCCL; }
    e->rank = rank; e->n_ranks = n_ranks;
    e->drop_variants();
    return 0;
}

}  /* extern "C" */
