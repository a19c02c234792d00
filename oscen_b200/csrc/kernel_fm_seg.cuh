/*
 * kernel_fm_seg.cuh — FMVoice (examples/fm-synth/src/fm_voice.rs:6-156) for banks whose value inputs are uniform
 * (everything but `frequency`): the frame loop is SEGMENTED the way the reference's block driver splits a block at its
 * event frames (oscen-graph-compiler/src/codegen/mod.rs:836-872), but per warp and also at envelope stage boundaries:
 *
 *   per warp:  L = min over lanes of { frames to the lane's next event, frames to the next stage completion of any of
 *              its four envelopes }                                                  (one REDUX.MIN per segment)
 *              L "plain" frames run a branch-free body — no event test, no stage machine: an envelope in Attack / Decay
 *              is `level = clamp(level + (target - level) * k)` with per-lane (target, k) fixed for the segment, an
 *              envelope in Sustain / Idle is a constant and costs nothing, Release lanes keep their division behind a
 *              warp-uniform test; the three operators call om_sinf_r (both polynomials, no range test: the host
 *              guarantees |phase| < 120 from the uniform levels / feedbacks, fm_select.hpp); the filter's coefficient
 *              refresh is absent when the cutoff cannot move (FMOD = false) and warp-uniform otherwise;
 *              then ONE "general" frame — the full per-frame walk of kernel_fm.cuh (events, gate handlers, stage
 *              completion, coefficient refresh) — handles the boundary frame.  Frame 0 of a block is always general.
 *
 * Every plain frame computes exactly the operations the general frame would execute for that lane state, in the same
 * order, so the output is bit-identical to fm_voice_kernel<false> / <true> and to the oracle (host-executed in the CPU
 * suite, tests/test_host_kernel_cpu.py).
 *
 * RAMP = true: the eight inputs a uniform ramp may drive without touching envelope timing or operator pitch
 * (op3/op2 level and feedback, route, filter cutoff / resonance / env amount) are read per frame from `a.utab`
 * (FmRow per frame, host-filled with the exact ValueRampState::tick() sequence and the values derived from it:
 * clamp(route), 1 - clamp(route), clamp(q), 1 / clamp(q)) — one broadcast row per frame instead of 28 per-thread
 * registers and per-frame `adsr_derived_stale` checks.
 */
#pragma once
#include "kernel_fm.cuh"

namespace oscen {

/* per-frame uniform values (and values derived from them) of the inputs a ramp may drive in this kernel */
struct FmRow {
    float lvl3, fb3, lvl2, fb2;
    float mixc, omix, cutoff, qc;
    float inv_q, amt, pad0, pad1;
};
constexpr uint32_t kFmRowFloats = 12;
static_assert(sizeof(FmRow) == kFmRowFloats * sizeof(float), "FmRow layout");
/* inputs whose ramps this kernel follows (bit i = value input i) */
constexpr uint32_t kFmSegRampMask = (1u << I_OP3_LEVEL) | (1u << I_OP3_FB) | (1u << I_OP2_LEVEL) | (1u << I_OP2_FB) |
                                    (1u << I_ROUTE) | (1u << I_CUTOFF) | (1u << I_RES) | (1u << I_ENV_AMT);

OSC_HD FmRow fm_row_from(const float *p) {
    FmRow r;
    r.lvl3 = p[I_OP3_LEVEL]; r.fb3 = p[I_OP3_FB]; r.lvl2 = p[I_OP2_LEVEL]; r.fb2 = p[I_OP2_FB];
    r.mixc = clampf(p[I_ROUTE], 0.0f, 1.0f);            /* Crossfade: clamp(mix) (crossfade.rs:39-44) */
    r.omix = 1.0f - r.mixc;
    r.cutoff = p[I_CUTOFF];
    r.qc = clampf(p[I_RES], 0.1f, 10.0f);               /* TptFilter::apply_parameter_updates (filters/tpt/mod.rs:84-99) */
    r.inv_q = 1.0f / r.qc;
    r.amt = p[I_ENV_AMT];
    r.pad0 = 0.0f; r.pad1 = 0.0f;
    return r;
}

/* FmOperator::process with the bounded-argument sine and `fract` as x - trunc(x) (one FRND + one FADD) */
OSC_D float fmop_process_r(FmOpState &s, float phase_mod, float feedback, float envelope, float level, const OmSinRegs &sk) {
    const float feedback_mod = s.prev * feedback;
    const float total_phase_mod = phase_mod + feedback_mod;
    const float phase_rad = (s.phase + total_phase_mod) * kTau;
    const float out = om_sinf_rk(phase_rad, sk) * envelope * level;
    s.prev = out;
    const float ph = s.phase + s.inc;
    s.phase = ph - truncf(ph);
    return out;
}

/* plain frames of one envelope.  `ad` / `rl` = this lane is in Attack|Decay / Release with rem >= 2 for the whole segment,
   any_ad = some lane of the warp is in Attack|Decay (warp-uniform).  A static lane (Sustain, Idle) keeps `level`.
   RELEASE = false: no lane of the warp is in Release in any envelope — the body holds no division (whose slow path is a
   call: a call inside the frame loop would force every loop-invariant constant to be rebuilt after it). */
template <bool RELEASE>
OSC_D float adsr_plain(AdsrState &s, bool ad, bool rl, float tgt, float k, bool any_ad) {
    if (RELEASE) {
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
    } else if (any_ad) {
        const float nl = clampf(s.level + (tgt - s.level) * k, 0.0f, 1.0f);
        s.level = ad ? nl : s.level;
    }
    return s.level;
}

/* TptFilter::process with cutoff fixed for the segment (coefficients current): the five-line core (mod.rs:101-123) */
OSC_D float tpt_core(TptState &s, float input) {
    const float high = (input - s.z0 * s.k - s.z1) * s.h;
    const float band = high * s.g + s.z0;
    const float low = band * s.g + s.z1;
    s.z0 = high * s.g + band;
    s.z1 = band * s.g + low;
    return low;
}
/* TptFilter::process with a cutoff that may move every frame (f_mod = 0): the refresh test is per lane, its body runs
   for the whole warp when any lane needs it (in a modulated block every lane does), results kept per lane */
OSC_D float tpt_process_mod(TptState &s, float input, float cutoff_in, float qc, float inv_q, float max_cutoff, float sr,
                            float two_sr, float period) {
    const float cutoff = clampf(cutoff_in, 20.0f, max_cutoff);
    const bool need = fabsf(cutoff - s.cur_cutoff) > FLT_EPSILON || fabsf(qc - s.cur_q) > FLT_EPSILON;
    if (__any_sync(0xffffffffu, need)) {
        const float arg = 2.0f * kPi * cutoff * period;
        if (__all_sync(0xffffffffu, om_tanf_mid_ok(arg))) {
            const float f = two_sr * om_k_tanf_mid(arg) * period;
            const float h = 1.0f / (1.0f + inv_q * f + f * f);
            const float k = f + inv_q;
            s.h = need ? h : s.h; s.g = need ? f : s.g; s.k = need ? k : s.k;
            s.cur_cutoff = need ? cutoff : s.cur_cutoff; s.cur_q = need ? qc : s.cur_q;
        } else if (need) {
            tpt_update_coefficients(s, sr, cutoff, qc);
        }
    }
    return tpt_core(s, input);
}

template <bool RAMP, bool FMOD>
__global__ void __launch_bounds__(kCta, 8) fm_voice_seg_kernel(const __grid_constant__ KernelArgs a) {
    __shared__ __align__(128) float tile[kTileF * kTileW];
    __shared__ float isum[kTileF * kCta];

    const uint32_t n = a.n_voices;
    const uint32_t v = blockIdx.x * kCta + threadIdx.x;
    const bool active = v < n;
    const uint32_t vs = active ? v : 0u;
    const uint32_t frames = a.ctl->frames;
    const float sr = a.u.sr;
    const uint32_t col = tile_col(threadIdx.x, a.vpi);

    /* ---- uniform value inputs: the constant bank (and, RAMP, one row of a.utab per frame) */
    const float *P = a.u.p;
    float freq = a.pv[vs];
    FmRow U = fm_row_from(P);
    const FmRow *rows = reinterpret_cast<const FmRow *>(a.utab);
    const float sus3 = P[I_OP3_S], sus2 = P[I_OP2_S], sus1 = P[I_OP1_S], susf = P[I_F_S];
    const float nyquist = sr * 0.5f - FLT_EPSILON;
    const float max_cutoff = fminf(nyquist, 20000.0f);
    const float period = 0.5f / sr, two_sr = 2.0f * sr;
    OmSinRegs sk = om_sin_regs();                    /* the sine's addend constants stay in registers for the whole block */
    sk.s2 = pin(sk.s2); sk.c3 = pin(sk.c3); sk.c1 = pin(sk.c1);

    /* ---- state */
    const float *sp = a.state;
    AdsrState e3, e2, e1, ef;
    adsr_load(e3, sp, FM_E3, n, vs); adsr_load(e2, sp, FM_E2, n, vs);
    adsr_load(e1, sp, FM_E1, n, vs); adsr_load(ef, sp, FM_EF, n, vs);
    FmOpState op3, op2, op1;
    op3.phase = sp[(size_t)(FM_OP3 + 0) * n + vs]; op3.prev = sp[(size_t)(FM_OP3 + 1) * n + vs];
    op2.phase = sp[(size_t)(FM_OP2 + 0) * n + vs]; op2.prev = sp[(size_t)(FM_OP2 + 1) * n + vs];
    op1.phase = sp[(size_t)(FM_OP1 + 0) * n + vs]; op1.prev = sp[(size_t)(FM_OP1 + 1) * n + vs];
#define OSC_SEG_INCS()                                                        \
    do {                                                                      \
        op3.inc = fmop_phase_inc(freq, P[I_OP3_RATIO], sr);                   \
        op2.inc = fmop_phase_inc(freq, P[I_OP2_RATIO], sr);                   \
        op1.inc = fmop_phase_inc(freq, P[I_OP1_RATIO], sr);                   \
    } while (0)
    OSC_SEG_INCS();
    TptState tp;
    tpt_load(tp, sp, FM_TPT, n, vs);

    const AdsrDerived d3 = a.u.env[0], d2 = a.u.env[1], d1 = a.u.env[2], df = a.u.env[3];
    AdsrHot h3, h2, h1, hf;
#define OSC_SEG_REFRESH(COMMIT)                                               \
    do {                                                                      \
        adsr_refresh<COMMIT>(e3, h3, sus3, d3); adsr_refresh<COMMIT>(e2, h2, sus2, d2); \
        adsr_refresh<COMMIT>(e1, h1, sus1, d1); adsr_refresh<COMMIT>(ef, hf, susf, df); \
    } while (0)
    EventCursor ec;
    ec.init(a.events, active ? a.ctl->n_events : 0u, vs);
    {
        const bool ev0 = ec.due(0u);
        adsr_begin_block(e3, d3, ev0); adsr_begin_block(e2, d2, ev0); adsr_begin_block(e1, d1, ev0); adsr_begin_block(ef, df, ev0);
    }
    OSC_SEG_REFRESH(false);

    const uint32_t kBig = 0x7FFFFFFFu;
    for (uint32_t f0 = 0; f0 < frames; f0 += kTileF) {
        const uint32_t fe = min(kTileF, frames - f0);
        uint32_t ff = 0;
        while (ff < fe) {
            const uint32_t f = f0 + ff;
            /* ---- how many plain frames can this warp run from here? */
            uint32_t np = active ? ec.next_frame - f : kBig;            /* 0 when an event is due now */
#define OSC_SEG_ENV_NP(e)                                                                                        \
            { const bool timed = (e.stage == ST_ATTACK) | (e.stage == ST_DECAY) | (e.stage == ST_RELEASE);       \
              const uint32_t c = timed ? (e.rem >= 2u ? e.rem - 1u : 0u) : kBig; np = umin(np, c); }
            OSC_SEG_ENV_NP(e3) OSC_SEG_ENV_NP(e2) OSC_SEG_ENV_NP(e1) OSC_SEG_ENV_NP(ef)
#undef OSC_SEG_ENV_NP
            if (f == 0u) np = 0u;                                       /* the block's first frame refreshes the filter / caches */
            const uint32_t L = umin(__reduce_min_sync(0xffffffffu, np), fe - ff);

            if (L > 0u) {
                /* ---- L plain frames: per-lane envelope modes are fixed for the segment */
#define OSC_SEG_MODE(e, h, d, X)                                                                                 \
                const bool ad##X = (e.stage == ST_ATTACK) | (e.stage == ST_DECAY), rl##X = e.stage == ST_RELEASE; \
                const float tg##X = e.stage == ST_ATTACK ? 1.0f : h.sus, kk##X = e.stage == ST_ATTACK ? d.a_k : d.d_k; \
                const bool aad##X = __any_sync(0xffffffffu, ad##X);
                OSC_SEG_MODE(e3, h3, d3, 3) OSC_SEG_MODE(e2, h2, d2, 2) OSC_SEG_MODE(e1, h1, d1, 1) OSC_SEG_MODE(ef, hf, df, F)
#undef OSC_SEG_MODE
                const bool any_rl = __any_sync(0xffffffffu, rl3 | rl2 | rl1 | rlF);
                const bool keepv = a.voice_out != nullptr && active;
                uint32_t vidx = f * n + v;                               /* n_voices * 512 < 2^32 (checked at create) */
                float *trow = tile + ff * kTileW + col;
#define OSC_SEG_PLAIN_FRAME(REL)                                                                          \
                {                                                                                         \
                    if (RAMP) U = rows[f + i];                                                            \
                    const float env3 = adsr_plain<REL>(e3, ad3, rl3, tg3, kk3, aad3);                     \
                    const float env2 = adsr_plain<REL>(e2, ad2, rl2, tg2, kk2, aad2);                     \
                    const float env1 = adsr_plain<REL>(e1, ad1, rl1, tg1, kk1, aad1);                     \
                    const float envf = adsr_plain<REL>(ef, adF, rlF, tgF, kkF, aadF);                     \
                    const float o3 = fmop_process_r(op3, 0.0f, U.fb3, env3, U.lvl3, sk);                      \
                    const float xa = o3 * U.omix;                                                         \
                    const float xb = o3 * U.mixc;                                                         \
                    const float o2 = fmop_process_r(op2, xa, U.fb2, env2, U.lvl2, sk);                        \
                    const float pm1 = o2 + xb;                                                            \
                    const float o1 = fmop_process_r(op1, pm1, 0.0f, env1, 1.0f, sk);                          \
                    float lp;                                                                             \
                    if (FMOD) {                                                                           \
                        const float feg = envf * U.amt;                                                   \
                        const float cutoff = feg + U.cutoff;                                              \
                        lp = tpt_process_mod(tp, o1, cutoff, U.qc, U.inv_q, max_cutoff, sr, two_sr, period); \
                    } else {                                                                              \
                        lp = tpt_core(tp, o1);                                                            \
                    }                                                                                     \
                    const float y = lp * 0.3f;                                                            \
                    if (keepv) a.voice_out[vidx] = y;                                                     \
                    vidx += n;                                                                            \
                    *trow = active ? y : 0.0f; trow += kTileW;                                            \
                }
                if (!any_rl) {
                    for (uint32_t i = 0; i < L; ++i) OSC_SEG_PLAIN_FRAME(false)
                    /* countdowns of the lanes that ran the one-pole form */
                    if (ad3) e3.rem -= L;
                    if (ad2) e2.rem -= L;
                    if (ad1) e1.rem -= L;
                    if (adF) ef.rem -= L;
                } else {
                    for (uint32_t i = 0; i < L; ++i) OSC_SEG_PLAIN_FRAME(true)      /* counts down per frame */
                }
#undef OSC_SEG_PLAIN_FRAME
                ff += L;
                continue;
            }

            /* ---- one general frame: the per-frame walk of kernel_fm.cuh */
            float y = 0.0f;
            if (RAMP) U = rows[f];
            if (active) {
                if (ec.due(f)) {
                    const uint32_t j = ec.end_of_frame(f);
                    for (uint32_t k = ec.i; k < j; ++k) {                /* value edges are assigned before process_event_inputs() */
                        const Event e = ec.ev[k];
                        if (e.input_id != kEventGate) { freq = e.scalar; a.pv[(size_t)e.input_id * n + v] = e.scalar; }   /* host guarantees id == frequency */
                    }
                    for (uint32_t k = ec.i; k < j; ++k) {
                        const Event e = ec.ev[k];
                        if (e.input_id == kEventGate) {                   /* gate -> env3, env2, env1, env_filter */
                            adsr_on_gate(e3, e.scalar, P[I_OP3_A], P[I_OP3_R], d3);
                            adsr_on_gate(e2, e.scalar, P[I_OP2_A], P[I_OP2_R], d2);
                            adsr_on_gate(e1, e.scalar, P[I_OP1_A], P[I_OP1_R], d1);
                            adsr_on_gate(ef, e.scalar, P[I_F_A], P[I_F_R], df);
                        }
                    }
                    OSC_SEG_INCS();
                    OSC_SEG_REFRESH(true);
                    ec.advance(j);
                }
                const float env3 = adsr_process_t<false>(e3, h3, sus3, d3);
                const float env2 = adsr_process_t<false>(e2, h2, sus2, d2);
                const float env1 = adsr_process_t<false>(e1, h1, sus1, d1);
                const float envf = adsr_process_t<false>(ef, hf, susf, df);
                const float feg = envf * U.amt;                          /* filter_env_gain: Gain */
                const float cutoff = feg + U.cutoff;                     /* cutoff_mod: AddValue  */
                const float o3 = fmop_process_r(op3, 0.0f, U.fb3, env3, U.lvl3, sk);
                const float xa = o3 * U.omix;                            /* op3_route: Crossfade  */
                const float xb = o3 * U.mixc;
                const float o2 = fmop_process_r(op2, xa, U.fb2, env2, U.lvl2, sk);
                const float pm1 = o2 + xb;                               /* op1_mod_mixer: Mixer  */
                const float o1 = fmop_process_r(op1, pm1, 0.0f, env1, 1.0f, sk);
                /* TptFilter::process (f_mod = 0): clamp, refresh test, core */
                const float cb = clampf(cutoff, 20.0f, max_cutoff);
                if (fabsf(cb - tp.cur_cutoff) > FLT_EPSILON || fabsf(U.qc - tp.cur_q) > FLT_EPSILON)
                    tpt_update_coefficients(tp, sr, cb, U.qc);
                const float lp = tpt_core(tp, o1);
                y = lp * 0.3f;                                           /* output_gain: Gain(0.3) */
                if (a.voice_out != nullptr) a.voice_out[(size_t)f * n + v] = y;
            }
            tile[ff * kTileW + col] = y;
            ff += 1u;
        }
        tile_flush_mono(a, tile, isum, f0, fe);
    }
#undef OSC_SEG_INCS
#undef OSC_SEG_REFRESH

    if (active) {
        float *dp = a.state;
        adsr_store(e3, dp, FM_E3, n, v); adsr_store(e2, dp, FM_E2, n, v);
        adsr_store(e1, dp, FM_E1, n, v); adsr_store(ef, dp, FM_EF, n, v);
        dp[(size_t)(FM_OP3 + 0) * n + v] = op3.phase; dp[(size_t)(FM_OP3 + 1) * n + v] = op3.prev;
        dp[(size_t)(FM_OP2 + 0) * n + v] = op2.phase; dp[(size_t)(FM_OP2 + 1) * n + v] = op2.prev;
        dp[(size_t)(FM_OP1 + 0) * n + v] = op1.phase; dp[(size_t)(FM_OP1 + 1) * n + v] = op1.prev;
        tpt_store(tp, dp, FM_TPT, n, v);
    }
}

}  // namespace oscen
