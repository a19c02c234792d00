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

/* per-frame uniform values (and values derived from them) of the inputs a ramp may drive in this kernel, grouped by the
   pipeline stage that reads them (stage A: op3; stage B: route + op2; stage C: op1's filter) */
struct FmRow {
    float lvl3, fb3, pad0, pad1;
    float lvl2, fb2, mixc, omix;
    float cutoff, qc, inv_q, amt;
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

/* ---- the plain frame as three pipeline stages.  Stage A (the four envelopes and op3) of frame i needs nothing from
   stages B (route, op2, mixer) and C (op1, filter, output) of earlier frames, B(i) needs A(i), C(i) needs B(i): a segment of
   L plain frames can run them SKEWED — one loop iteration computes C(i), B(i+1), A(i+2) — which puts three independent
   sine chains into one basic block instead of one chain of three.  At 65 536 voices an SM sub-partition holds 3.5 warps:
   the frame loop is bound by each warp's own dependency chain, not by issue slots, and the skew cuts that chain to a third.
   Same operations per voice in the same order per node as the unskewed frame, so the same bits. */
struct FmRegs {
    AdsrState e3, e2, e1, ef;
    FmOpState op3, op2, op1;
    TptState tp;
};
struct FmEnvMode {                       /* per lane, fixed for a segment: in Attack|Decay?, its target and coefficient */
    bool ad3, ad2, ad1, adF;
    float tg3, tg2, tg1, tgF, kk3, kk2, kk1, kkF;
};
struct FmStageA { float o3, env2, env1, envf; };
struct FmStageB { float pm1, env1, envf; };
struct FmConsts { float max_cutoff, sr, two_sr, period; };

OSC_D FmStageA fm_stage_a(FmRegs &r, const FmEnvMode &m, float lvl3, float fb3, const OmSinRegs &sk) {
    FmStageA o;
    const float env3 = adsr_ad(r.e3, m.ad3, m.tg3, m.kk3);
    o.env2 = adsr_ad(r.e2, m.ad2, m.tg2, m.kk2);
    o.env1 = adsr_ad(r.e1, m.ad1, m.tg1, m.kk1);
    o.envf = adsr_ad(r.ef, m.adF, m.tgF, m.kkF);
    o.o3 = fmop_process_r(r.op3, 0.0f, fb3, env3, lvl3, sk);
    return o;
}
OSC_D FmStageB fm_stage_b(FmRegs &r, const FmStageA &a, float lvl2, float fb2, float mixc, float omix, const OmSinRegs &sk) {
    FmStageB o;
    const float xa = a.o3 * omix;                            /* op3_route: Crossfade */
    const float xb = a.o3 * mixc;
    const float o2 = fmop_process_r(r.op2, xa, fb2, a.env2, lvl2, sk);
    o.pm1 = o2 + xb;                                         /* op1_mod_mixer: Mixer */
    o.env1 = a.env1; o.envf = a.envf;
    return o;
}
template <bool FMOD>
OSC_D float fm_stage_c(FmRegs &r, const FmStageB &b, float cutoff_p, float qc, float inv_q, float amt, const FmConsts &c, const OmSinRegs &sk) {
    const float o1 = fmop_process_r(r.op1, b.pm1, 0.0f, b.env1, 1.0f, sk);
    float lp;
    if (FMOD) {
        const float feg = b.envf * amt;                      /* filter_env_gain: Gain */
        const float cutoff = feg + cutoff_p;                 /* cutoff_mod: AddValue  */
        lp = tpt_process_mod(r.tp, o1, cutoff, qc, inv_q, c.max_cutoff, c.sr, c.two_sr, c.period);
    } else {
        lp = tpt_core(r.tp, o1);
    }
    return lp * 0.3f;                                        /* output_gain: Gain(0.3) */
}

/* FmOperator::process split around its sine, so that a skewed iteration can evaluate its three sines together:
   fmop_arg = the phase argument (state untouched), fmop_finish = output, prev_output and phase advance.
   `+ 0.0` terms of unconnected endpoints (op3.phase_mod, op1.feedback * prev) are left out where they cannot change a bit:
   an operator phase is x - trunc(x), never -0, so adding +0 or -0 to it gives the same sum. */
OSC_D float fmop_arg(const FmOpState &s, float total_phase_mod) { return (s.phase + total_phase_mod) * kTau; }
OSC_D float fmop_finish(FmOpState &s, float sine, float envelope, float level) {
    const float out = sine * envelope * level;
    s.prev = out;
    const float ph = s.phase + s.inc;
    s.phase = ph - truncf(ph);
    return out;
}

/* L >= 3 plain frames from frame f, no lane in Release: the skewed schedule.  Iteration j: stage C of frame j (op1, filter,
   output), stage B of frame j+1 (route, op2, mixer), stage A of frame j+2 (envelopes, op3) — their three sines in one call */
template <bool RAMP, bool FMOD, bool KEEPV>
OSC_D void fm_segment_skewed(FmRegs &r, const FmEnvMode &m, const FmRow &U0, const FmRow *rows, uint32_t f, uint32_t L,
                             const FmConsts &c, const OmSinRegs &sk, float *trow, float *voice_out, uint32_t vidx, uint32_t n,
                             bool keepv, bool active) {
    FmRow U = U0;
#define OSC_ROW(i, field) (RAMP ? rows[i].field : U.field)
    auto emit = [&](float y) {
        if (KEEPV) { if (keepv) voice_out[vidx] = y; vidx += n; }
        *trow = active ? y : 0.0f; trow += kTileW;
    };
    /* prologue: A(f), A(f+1), B(f) */
    const FmStageA a0 = fm_stage_a(r, m, OSC_ROW(f, lvl3), OSC_ROW(f, fb3), sk);
    FmStageA a = fm_stage_a(r, m, OSC_ROW(f + 1u, lvl3), OSC_ROW(f + 1u, fb3), sk);
    FmStageB b = fm_stage_b(r, a0, OSC_ROW(f, lvl2), OSC_ROW(f, fb2), OSC_ROW(f, mixc), OSC_ROW(f, omix), sk);
    for (uint32_t j = 0; j + 2u < L; ++j) {
        const uint32_t fc = f + j, fb = fc + 1u, fa = fc + 2u;
        /* arguments of the three sines: op1 of frame fc, op2 of frame fb, op3 of frame fa */
        const float env3 = adsr_ad(r.e3, m.ad3, m.tg3, m.kk3);
        FmStageA an;
        an.env2 = adsr_ad(r.e2, m.ad2, m.tg2, m.kk2);
        an.env1 = adsr_ad(r.e1, m.ad1, m.tg1, m.kk1);
        an.envf = adsr_ad(r.ef, m.adF, m.tgF, m.kkF);
        const float arg1 = fmop_arg(r.op1, b.pm1);
        const float xa = a.o3 * OSC_ROW(fb, omix), xb = a.o3 * OSC_ROW(fb, mixc);
        const float arg2 = fmop_arg(r.op2, xa + r.op2.prev * OSC_ROW(fb, fb2));
        const float arg3 = fmop_arg(r.op3, r.op3.prev * OSC_ROW(fa, fb3));
        float s1, s2, s3;
        om_sinf_rk3(arg1, arg2, arg3, sk, s1, s2, s3);
        /* stage C(fc) */
        const float o1 = fmop_finish(r.op1, s1, b.env1, 1.0f);
        float lp;
        if (FMOD) {
            const float feg = b.envf * OSC_ROW(fc, amt);
            const float cutoff = feg + OSC_ROW(fc, cutoff);
            lp = tpt_process_mod(r.tp, o1, cutoff, OSC_ROW(fc, qc), OSC_ROW(fc, inv_q), c.max_cutoff, c.sr, c.two_sr, c.period);
        } else {
            lp = tpt_core(r.tp, o1);
        }
        emit(lp * 0.3f);
        /* stage B(fb) */
        FmStageB bn;
        const float o2 = fmop_finish(r.op2, s2, a.env2, OSC_ROW(fb, lvl2));
        bn.pm1 = o2 + xb; bn.env1 = a.env1; bn.envf = a.envf;
        /* stage A(fa) */
        an.o3 = fmop_finish(r.op3, s3, env3, OSC_ROW(fa, lvl3));
        b = bn; a = an;
    }
    {
        const uint32_t fc = f + L - 2u, fl = f + L - 1u;
        const float y0 = fm_stage_c<FMOD>(r, b, OSC_ROW(fc, cutoff), OSC_ROW(fc, qc), OSC_ROW(fc, inv_q), OSC_ROW(fc, amt), c, sk);
        const FmStageB bl = fm_stage_b(r, a, OSC_ROW(fl, lvl2), OSC_ROW(fl, fb2), OSC_ROW(fl, mixc), OSC_ROW(fl, omix), sk);
        emit(y0);
        const float y1 = fm_stage_c<FMOD>(r, bl, OSC_ROW(fl, cutoff), OSC_ROW(fl, qc), OSC_ROW(fl, inv_q), OSC_ROW(fl, amt), c, sk);
        emit(y1);
    }
#undef OSC_ROW
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
    FmConsts kc;
    kc.sr = sr;
    kc.max_cutoff = fminf(sr * 0.5f - FLT_EPSILON, 20000.0f);
    kc.period = 0.5f / sr; kc.two_sr = 2.0f * sr;
    const OmSinRegs sk = om_sin_regs();              /* the sine's addend constants: loaded once, live for the whole block */

    /* ---- state */
    const float *sp = a.state;
    FmRegs r;
    adsr_load(r.e3, sp, FM_E3, n, vs); adsr_load(r.e2, sp, FM_E2, n, vs);
    adsr_load(r.e1, sp, FM_E1, n, vs); adsr_load(r.ef, sp, FM_EF, n, vs);
    r.op3.phase = sp[(size_t)(FM_OP3 + 0) * n + vs]; r.op3.prev = sp[(size_t)(FM_OP3 + 1) * n + vs];
    r.op2.phase = sp[(size_t)(FM_OP2 + 0) * n + vs]; r.op2.prev = sp[(size_t)(FM_OP2 + 1) * n + vs];
    r.op1.phase = sp[(size_t)(FM_OP1 + 0) * n + vs]; r.op1.prev = sp[(size_t)(FM_OP1 + 1) * n + vs];
#define OSC_SEG_INCS()                                                        \
    do {                                                                      \
        r.op3.inc = fmop_phase_inc(freq, P[I_OP3_RATIO], sr);                 \
        r.op2.inc = fmop_phase_inc(freq, P[I_OP2_RATIO], sr);                 \
        r.op1.inc = fmop_phase_inc(freq, P[I_OP1_RATIO], sr);                 \
    } while (0)
    OSC_SEG_INCS();
    tpt_load(r.tp, sp, FM_TPT, n, vs);

    const AdsrDerived d3 = a.u.env[0], d2 = a.u.env[1], d1 = a.u.env[2], df = a.u.env[3];
    AdsrHot h3, h2, h1, hf;
#define OSC_SEG_REFRESH(COMMIT)                                               \
    do {                                                                      \
        adsr_refresh<COMMIT>(r.e3, h3, sus3, d3); adsr_refresh<COMMIT>(r.e2, h2, sus2, d2); \
        adsr_refresh<COMMIT>(r.e1, h1, sus1, d1); adsr_refresh<COMMIT>(r.ef, hf, susf, df); \
    } while (0)
    EventCursor ec;
    ec.init(a.events, active ? a.ctl->n_events : 0u, vs);
    {
        const bool ev0 = ec.due(0u);
        adsr_begin_block(r.e3, d3, ev0); adsr_begin_block(r.e2, d2, ev0); adsr_begin_block(r.e1, d1, ev0); adsr_begin_block(r.ef, df, ev0);
    }
    OSC_SEG_REFRESH(false);

    const uint32_t kBig = 0x7FFFFFFFu;
    const bool keepv = a.voice_out != nullptr && active;
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
            OSC_SEG_ENV_NP(r.e3) OSC_SEG_ENV_NP(r.e2) OSC_SEG_ENV_NP(r.e1) OSC_SEG_ENV_NP(r.ef)
#undef OSC_SEG_ENV_NP
            if (f == 0u) np = 0u;                                       /* the block's first frame refreshes the filter / caches */
            const uint32_t L = umin(__reduce_min_sync(0xffffffffu, np), fe - ff);

            if (L > 0u) {
                /* ---- L plain frames: per-lane envelope modes are fixed for the segment */
                FmEnvMode m;
#define OSC_SEG_MODE(e, h, d, X)                                                                                 \
                m.ad##X = (e.stage == ST_ATTACK) | (e.stage == ST_DECAY);                                        \
                m.tg##X = e.stage == ST_ATTACK ? 1.0f : h.sus; m.kk##X = e.stage == ST_ATTACK ? d.a_k : d.d_k;
                OSC_SEG_MODE(r.e3, h3, d3, 3) OSC_SEG_MODE(r.e2, h2, d2, 2) OSC_SEG_MODE(r.e1, h1, d1, 1) OSC_SEG_MODE(r.ef, hf, df, F)
#undef OSC_SEG_MODE
                const bool rl3 = r.e3.stage == ST_RELEASE, rl2 = r.e2.stage == ST_RELEASE, rl1 = r.e1.stage == ST_RELEASE, rlF = r.ef.stage == ST_RELEASE;
                const bool any_rl = __any_sync(0xffffffffu, rl3 | rl2 | rl1 | rlF);
                const uint32_t vidx = f * n + v;                         /* n_voices * 512 < 2^32 (checked at create) */
                float *trow = tile + ff * kTileW + col;
                if (!any_rl) {
                    if (L >= 3u && !(a.pad0 & 1u)) {                     /* pad0 bit 0: diagnostics switch OSCEN_B200_FM_SKEW=0 */
                        if (a.voice_out != nullptr) fm_segment_skewed<RAMP, FMOD, true>(r, m, U, rows, f, L, kc, sk, trow, a.voice_out, vidx, n, keepv, active);
                        else fm_segment_skewed<RAMP, FMOD, false>(r, m, U, rows, f, L, kc, sk, trow, a.voice_out, vidx, n, keepv, active);
                    } else {
                        uint32_t vi = vidx;
                        for (uint32_t i = 0; i < L; ++i) {
                            if (RAMP) U = rows[f + i];
                            const FmStageA sa = fm_stage_a(r, m, U.lvl3, U.fb3, sk);
                            const FmStageB sb = fm_stage_b(r, sa, U.lvl2, U.fb2, U.mixc, U.omix, sk);
                            const float y = fm_stage_c<FMOD>(r, sb, U.cutoff, U.qc, U.inv_q, U.amt, kc, sk);
                            if (keepv) a.voice_out[vi] = y;
                            vi += n;
                            *trow = active ? y : 0.0f; trow += kTileW;
                        }
                    }
                    /* countdowns of the lanes that ran the one-pole form */
                    if (m.ad3) r.e3.rem -= L;
                    if (m.ad2) r.e2.rem -= L;
                    if (m.ad1) r.e1.rem -= L;
                    if (m.adF) r.ef.rem -= L;
                } else {
                    /* some lane is in Release: the generic envelope frame (its division counts down per frame) */
                    uint32_t vi = vidx;
                    for (uint32_t i = 0; i < L; ++i) {
                        if (RAMP) U = rows[f + i];
                        FmStageA sa;
                        const float env3 = adsr_plain_rel(r.e3, m.ad3, rl3, m.tg3, m.kk3);
                        sa.env2 = adsr_plain_rel(r.e2, m.ad2, rl2, m.tg2, m.kk2);
                        sa.env1 = adsr_plain_rel(r.e1, m.ad1, rl1, m.tg1, m.kk1);
                        sa.envf = adsr_plain_rel(r.ef, m.adF, rlF, m.tgF, m.kkF);
                        sa.o3 = fmop_process_r(r.op3, 0.0f, U.fb3, env3, U.lvl3, sk);
                        const FmStageB sb = fm_stage_b(r, sa, U.lvl2, U.fb2, U.mixc, U.omix, sk);
                        const float y = fm_stage_c<FMOD>(r, sb, U.cutoff, U.qc, U.inv_q, U.amt, kc, sk);
                        if (keepv) a.voice_out[vi] = y;
                        vi += n;
                        *trow = active ? y : 0.0f; trow += kTileW;
                    }
                }
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
                            adsr_on_gate(r.e3, e.scalar, P[I_OP3_A], P[I_OP3_R], d3);
                            adsr_on_gate(r.e2, e.scalar, P[I_OP2_A], P[I_OP2_R], d2);
                            adsr_on_gate(r.e1, e.scalar, P[I_OP1_A], P[I_OP1_R], d1);
                            adsr_on_gate(r.ef, e.scalar, P[I_F_A], P[I_F_R], df);
                        }
                    }
                    OSC_SEG_INCS();
                    OSC_SEG_REFRESH(true);
                    ec.advance(j);
                }
                FmStageA sa;
                const float env3 = adsr_process_t<false>(r.e3, h3, sus3, d3);
                sa.env2 = adsr_process_t<false>(r.e2, h2, sus2, d2);
                sa.env1 = adsr_process_t<false>(r.e1, h1, sus1, d1);
                sa.envf = adsr_process_t<false>(r.ef, hf, susf, df);
                sa.o3 = fmop_process_r(r.op3, 0.0f, U.fb3, env3, U.lvl3, sk);
                const FmStageB sb = fm_stage_b(r, sa, U.lvl2, U.fb2, U.mixc, U.omix, sk);
                const float o1 = fmop_process_r(r.op1, sb.pm1, 0.0f, sb.env1, 1.0f, sk);
                /* TptFilter::process (f_mod = 0): clamp, refresh test, core */
                const float feg = sb.envf * U.amt;                       /* filter_env_gain: Gain */
                const float cutoff = feg + U.cutoff;                     /* cutoff_mod: AddValue  */
                const float cb = clampf(cutoff, 20.0f, kc.max_cutoff);
                if (fabsf(cb - r.tp.cur_cutoff) > FLT_EPSILON || fabsf(U.qc - r.tp.cur_q) > FLT_EPSILON)
                    tpt_update_coefficients(r.tp, sr, cb, U.qc);
                const float lp = tpt_core(r.tp, o1);
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
        adsr_store(r.e3, dp, FM_E3, n, v); adsr_store(r.e2, dp, FM_E2, n, v);
        adsr_store(r.e1, dp, FM_E1, n, v); adsr_store(r.ef, dp, FM_EF, n, v);
        dp[(size_t)(FM_OP3 + 0) * n + v] = r.op3.phase; dp[(size_t)(FM_OP3 + 1) * n + v] = r.op3.prev;
        dp[(size_t)(FM_OP2 + 0) * n + v] = r.op2.phase; dp[(size_t)(FM_OP2 + 1) * n + v] = r.op2.prev;
        dp[(size_t)(FM_OP1 + 0) * n + v] = r.op1.phase; dp[(size_t)(FM_OP1 + 1) * n + v] = r.op1.prev;
        tpt_store(r.tp, dp, FM_TPT, n, v);
    }
}

}  // namespace oscen
