/*
 * engine_types.h — host/device shared plain structs of the engine (kernel argument blocks,
 * state-field maps).  See DESIGN.md "Data layout in HBM".
 */
#pragma once
#if !defined(__CUDACC_RTC__)
#include <stdint.h>
#endif

#include "nodes.cuh"

namespace oscen {

constexpr uint32_t kMaxInputs = 28;
constexpr uint32_t kMaxBlock = 512;
constexpr uint32_t kCta = 64;            /* voices per CTA (one thread = one voice): small CTAs balance
                                            148 SMs even at 65 536 voices (1024 CTAs, 6.9 per SM) */
constexpr uint32_t kTileF = 32;          /* frames per shared-memory output tile */
constexpr uint32_t kTileW = kCta + kCta / 2;   /* padded tile row: kCta + kCta/vpi (vpi >= 2) */
constexpr uint32_t kEventGate = 0xFFFFFFFFu;

struct Event {                           /* == oscen_b200_event */
    uint32_t voice, frame, input_id;
    float scalar;
};

/* per-block control words that change from block to block; lives in device memory so a
   captured CUDA graph can be replayed unchanged (kernels read it through a pointer) */
struct BlockCtl {
    uint32_t frames;
    uint32_t n_events;
    uint32_t ramp_mask;                  /* bit i: value input i follows ptab[frame][i] this block */
    uint32_t lcg_on;                     /* TWIN_PEAKS: generate audio_in from the LCG */
};

/* ---- FM voice: state fields [field][voice] -------------------------------------------- */
enum : uint32_t {
    FM_E3 = 0, FM_E2 = 4, FM_E1 = 8, FM_EF = 12,      /* 4 x {stage, rem, level, vel} */
    FM_OP3 = 16, FM_OP2 = 18, FM_OP1 = 20,           /* 3 x {phase, prev} */
    FM_TPT = 22,                                     /* z0 z1 cur_cutoff cur_q h g k */
    FM_NSTATE = 29
};
/* TwinPeaks: 2 x {z0 z1 z2 g h last_cutoff last_fmod last_resonance resonance} + lcg(2 words) */
enum : uint32_t { TW_A = 0, TW_B = 9, TW_LCG = 18, TW_NSTATE = 20 };
/* Sat: phase + 2 stages x 24 taps (chronological, newest last) */
enum : uint32_t { SAT_PHASE = 0, SAT_H1 = 1, SAT_H2 = 25, SAT_NSTATE = 49 };
/* Synth (+ADSR): modulator phase, carrier phase, tpt(7), adsr(4) */
enum : uint32_t { SY_MOD = 0, SY_CAR = 1, SY_TPT = 2, SY_ENV = 9, SY_NSTATE = 13 };
/* EPiano voice: amp cur[32] target[32] decay[32] release[32], osc re[32] im[32], mul re[32] im[32],
   released, interp_step, last_frequency  (instance: tremolo phase in a separate array) */
enum : uint32_t { EP_CUR = 0, EP_TGT = 32, EP_DEC = 64, EP_REL = 96, EP_ORE = 128, EP_OIM = 160,
                  EP_MRE = 192, EP_MIM = 224, EP_RELEASED = 256, EP_STEP = 257, EP_LASTF = 258,
                  EP_NSTATE = 259 };

/* `external NAME: AudioAsset;` slots of a DSL graph: one mono buffer per slot, shared by every voice of the engine (the
   reference's AssetSlot holds one Arc<Vec<f32>> per graph instance; a bank plays one sample set).  Lives in device memory
   so that a load between blocks changes no kernel argument.  `reset[k]` is set for the first block after a load: the
   next process() picks the buffer up and restarts its playhead (sample_player/mod.rs:107-112). */
constexpr uint32_t kMaxAssets = 4;
struct AssetTable {
    const float *data[kMaxAssets];
    uint32_t frames[kMaxAssets];
    uint32_t reset[kMaxAssets];
};

struct Uniform {
    float p[kMaxInputs];                 /* uniform value of every value input */
    AdsrDerived env[4];                  /* host-derived ADSR step data for uniform a/d/r */
    float sr;
};

struct KernelArgs {
    uint32_t n_voices, vpi, n_inst, channels;
    uint32_t voice_offset;               /* global index of voice 0 of this shard (LCG seeding) */
    uint32_t pad0;
    unsigned long long lcg_seed;
    float *state;                        /* [n_state][n_voices] */
    float *pv;                           /* [n_inputs][n_voices] per-voice value inputs */
    const Event *events;                 /* sorted by (voice, frame, seq) */
    const float *ptab;                   /* [512][kMaxInputs] per-frame uniform values (ramps) */
    const BlockCtl *ctl;
    const float *stream_in;              /* [frames][n_voices] or null */
    float *voice_out;                    /* [512][n_voices] or null */
    float *inst_out;                     /* [512][n_inst][channels] or null */
    float *inst_state;                   /* [n_inst] instance-level state (EPIANO tremolo phase) */
    float *mix_partial;                  /* [n_cta][512][channels] */
    const AssetTable *assets;            /* DSL graphs with `external` slots, else null */
    const float *utab;                   /* [512][12] per-frame uniform rows of fm_voice_seg_kernel<RAMP> (kernel_fm_seg.cuh FmRow) */
    Uniform u;
};

}  // namespace oscen
