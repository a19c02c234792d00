/*
 * kernel_fm.cuh — FMVoice (examples/fm-synth/src/fm_voice.rs:6-156) as ONE kernel:
 * one thread renders one voice for the whole block with all 14 nodes' state in registers,
 * walking the nodes in the generated topological order:
 *   env3, env2, env1, env_filter, filter_env_gain, cutoff_mod, op3_osc, op3_route, op2_osc,
 *   op1_mod_mixer, op1_osc, filter, output_gain.
 * HBM traffic per block: state (29 words) in + out, the per-voice frequency, and whatever
 * outputs the engine was asked to keep.
 *
 * GENERAL=false: every value input except `frequency` is uniform and constant over the block
 *   (uniform values and the ADSR step data come from the kernel-argument constant bank).
 * GENERAL=true : any input may be per-voice, ramped per frame (ptab) or set by a value event.
 */
#pragma once
#include "kernels_common.cuh"
#include "state_io.cuh"

/* KernelArgs.pad0 bit 1 (engine: OSCEN_B200_TMA_STATE=1): move the per-voice state rows HBM <-> shared memory with TMA
   bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP / SYNCS) instead of per-thread coalesced loads / stores — north_star's
   "TMA for block-to-block state handoff".  Bit-exact (tests/test_parity_gpu.py runs both), measured SLOWER on B200 (0.2736
   vs 0.2652 ms per block at 65 536 voices; 3.76 vs 3.67 ms at 1 M): the state is touched once per 256-frame block
   (1.4 B/voice-sample), so the extra barrier, the mbarrier wait before the first instruction of the recurrence and the
   shared-memory round trip cost more than the handful of LDG / STG they replace.  Off by default; a run-time switch so
   that the one library the tests and the bench load carries, and keeps testing, both paths. */

namespace oscen {

enum : uint32_t {   /* == OSCEN_B200_FM_* */
    I_FREQ = 0, I_OP3_RATIO, I_OP3_LEVEL, I_OP3_FB, I_OP3_A, I_OP3_D, I_OP3_S, I_OP3_R,
    I_OP2_RATIO, I_OP2_LEVEL, I_OP2_FB, I_OP2_A, I_OP2_D, I_OP2_S, I_OP2_R,
    I_OP1_RATIO, I_OP1_A, I_OP1_D, I_OP1_S, I_OP1_R, I_ROUTE,
    I_CUTOFF, I_RES, I_F_A, I_F_D, I_F_S, I_F_R, I_ENV_AMT, I_FM_N
};

template <bool GENERAL>
__global__ void __launch_bounds__(kCta, 8) fm_voice_kernel(const __grid_constant__ KernelArgs a) {
    __shared__ __align__(128) float tile[kTileF * kTileW];    /* also the TMA staging area for the state rows */
    __shared__ float isum[kTileF * kCta];
    __shared__ __align__(8) unsigned long long state_bar;
    static_assert(FM_NSTATE * kCta <= kTileF * kTileW, "state staging fits in the output tile");

    const uint32_t n = a.n_voices;
    const uint32_t v = blockIdx.x * kCta + threadIdx.x;
    const bool active = v < n;
    const uint32_t vs = active ? v : 0u;
    const uint32_t frames = a.ctl->frames;
    const uint32_t ramp_mask = GENERAL ? a.ctl->ramp_mask : 0u;
    const float sr = a.u.sr;
    const uint32_t col = tile_col(threadIdx.x, a.vpi);

    /* ---- value inputs */
    float P[I_FM_N];
#pragma unroll
    for (uint32_t i = 0; i < I_FM_N; ++i) P[i] = GENERAL ? a.pv[(size_t)i * n + vs] : a.u.p[i];
    if (!GENERAL) P[I_FREQ] = a.pv[vs];

    /* ---- state: block-to-block handoff through TMA bulk copies (HBM rows -> smem -> registers) */
    const float *sp = a.state; uint32_t sn = n, sv = vs;
    bool staged = false;                 /* default: plain coalesced loads (measured 2-3 % faster) */
    if (a.pad0 & 2u) staged = state_stage_in(tile, &state_bar, a.state, FM_NSTATE, n, sp, sn, sv, vs);
    AdsrState e3, e2, e1, ef;
    adsr_load(e3, sp, FM_E3, sn, sv); adsr_load(e2, sp, FM_E2, sn, sv);
    adsr_load(e1, sp, FM_E1, sn, sv); adsr_load(ef, sp, FM_EF, sn, sv);
    FmOpState op3, op2, op1;
    op3.phase = sp[(size_t)(FM_OP3 + 0) * sn + sv]; op3.prev = sp[(size_t)(FM_OP3 + 1) * sn + sv];
    op2.phase = sp[(size_t)(FM_OP2 + 0) * sn + sv]; op2.prev = sp[(size_t)(FM_OP2 + 1) * sn + sv];
    op1.phase = sp[(size_t)(FM_OP1 + 0) * sn + sv]; op1.prev = sp[(size_t)(FM_OP1 + 1) * sn + sv];
#define OSC_FM_INCS()                                                        \
    do {                                                                     \
        op3.inc = fmop_phase_inc(P[I_FREQ], P[I_OP3_RATIO], sr);             \
        op2.inc = fmop_phase_inc(P[I_FREQ], P[I_OP2_RATIO], sr);             \
        op1.inc = fmop_phase_inc(P[I_FREQ], P[I_OP1_RATIO], sr);             \
    } while (0)
    OSC_FM_INCS();
    TptState tp;
    tpt_load(tp, sp, FM_TPT, sn, sv);
    __syncthreads();                                         /* staging area becomes the output tile */

    AdsrDerived d3, d2, d1, df;
    if (GENERAL) {
        adsr_derive_r(d3, P[I_OP3_A], P[I_OP3_D], P[I_OP3_R], sr);
        adsr_derive_r(d2, P[I_OP2_A], P[I_OP2_D], P[I_OP2_R], sr);
        adsr_derive_r(d1, P[I_OP1_A], P[I_OP1_D], P[I_OP1_R], sr);
        adsr_derive_r(df, P[I_F_A], P[I_F_D], P[I_F_R], sr);
    } else {
        d3 = a.u.env[0]; d2 = a.u.env[1]; d1 = a.u.env[2]; df = a.u.env[3];
        adsr_pin(d3); adsr_pin(d2); adsr_pin(d1); adsr_pin(df);
    }

    /* FAST variant: sustain and the stage lengths cannot change inside the block, so the
       sustain-level / countdown-clamp part of update_sustain_level() runs at block start and after
       gate events only (see adsr_refresh) */
    AdsrHot h3, h2, h1, hf;
#define OSC_FM_REFRESH(COMMIT)                                              \
    do {                                                                    \
        adsr_refresh<COMMIT>(e3, h3, P[I_OP3_S], d3); adsr_refresh<COMMIT>(e2, h2, P[I_OP2_S], d2); \
        adsr_refresh<COMMIT>(e1, h1, P[I_OP1_S], d1); adsr_refresh<COMMIT>(ef, hf, P[I_F_S], df);   \
    } while (0)
    const OmSinRegs sk = om_sin_regs();
    EventCursor ec;
    ec.init(a.events, active ? a.ctl->n_events : 0u, vs);
    {   /* the release length a gate-off at frame 0 starts from is the previous block's (AdsrState::r_gate) */
        const bool ev0 = ec.due(0u);
        adsr_begin_block(e3, d3, ev0); adsr_begin_block(e2, d2, ev0); adsr_begin_block(e1, d1, ev0); adsr_begin_block(ef, df, ev0);
    }
    OSC_FM_REFRESH(false);

    for (uint32_t f0 = 0; f0 < frames; f0 += kTileF) {
        const uint32_t fe = min(kTileF, frames - f0);
        uint32_t ff = 0;
        while (ff < fe) {
            const uint32_t f = f0 + ff;
            /* ---- plain segment (round 2, the scheme of the generated kernels): per warp, L frames until the first lane's next
               event or envelope stage completion run without ramp / event / staleness code, the envelopes in their segment
               form; every per-voice parameter is loop-invariant there.  Frame 0 and ramp blocks stay per-frame. */
            {
                uint32_t np = (ramp_mask != 0u || f == 0u) ? 0u : (active ? ec.next_frame - f : 0x7FFFFFFFu);
                np = umin(umin(np, adsr_plain_frames(e3)), umin(adsr_plain_frames(e2), umin(adsr_plain_frames(e1), adsr_plain_frames(ef))));
                const uint32_t L = umin(__reduce_min_sync(0xffffffffu, np), fe - ff);
                if (L > 0u) {
#define OSC_FM_MODE(e, h, d, X)                                                                                      \
                    const bool ad##X = (e.stage == ST_ATTACK) | (e.stage == ST_DECAY), rl##X = e.stage == ST_RELEASE;  \
                    const float tg##X = e.stage == ST_ATTACK ? 1.0f : h.sus, kk##X = e.stage == ST_ATTACK ? d.a_k : d.d_k;
                    OSC_FM_MODE(e3, h3, d3, 3) OSC_FM_MODE(e2, h2, d2, 2) OSC_FM_MODE(e1, h1, d1, 1) OSC_FM_MODE(ef, hf, df, F)
#undef OSC_FM_MODE
                    const bool any_rl = __any_sync(0xffffffffu, rl3 | rl2 | rl1 | rlF);
                    for (uint32_t i = 0; i < L; ++i) {
                        float y = 0.0f;
                        if (active) {
                            const float env3 = any_rl ? adsr_plain_rel(e3, ad3, rl3, tg3, kk3) : adsr_ad(e3, ad3, tg3, kk3);
                            const float env2 = any_rl ? adsr_plain_rel(e2, ad2, rl2, tg2, kk2) : adsr_ad(e2, ad2, tg2, kk2);
                            const float env1 = any_rl ? adsr_plain_rel(e1, ad1, rl1, tg1, kk1) : adsr_ad(e1, ad1, tg1, kk1);
                            const float envf = any_rl ? adsr_plain_rel(ef, adF, rlF, tgF, kkF) : adsr_ad(ef, adF, tgF, kkF);
                            const float feg = envf * P[I_ENV_AMT];
                            const float cutoff = feg + P[I_CUTOFF];
                            const float o3 = fmop_process_k(op3, 0.0f, P[I_OP3_FB], env3, P[I_OP3_LEVEL], sk);
                            const float mix = clampf(P[I_ROUTE], 0.0f, 1.0f);
                            const float xa = o3 * (1.0f - mix);
                            const float xb = o3 * mix;
                            const float o2 = fmop_process_k(op2, xa, P[I_OP2_FB], env2, P[I_OP2_LEVEL], sk);
                            const float pm1 = o2 + xb;
                            const float o1 = fmop_process_k(op1, pm1, 0.0f, env1, 1.0f, sk);
                            const float lp = tpt_process(tp, o1, cutoff, P[I_RES], 0.0f, sr);
                            y = lp * 0.3f;
                            if (a.voice_out != nullptr) a.voice_out[(size_t)(f + i) * n + v] = y;
                        }
                        tile[(ff + i) * kTileW + col] = y;
                    }
                    if (!any_rl) {                                      /* the Release form counts down per frame itself */
                        if (ad3) e3.rem -= L;
                        if (ad2) e2.rem -= L;
                        if (ad1) e1.rem -= L;
                        if (adF) ef.rem -= L;
                    }
                    ff += L;
                    continue;
                }
            }
            float y = 0.0f;
            if (active) {
                if (GENERAL && ramp_mask) {               /* ramped inputs read `.current` */
                    const float *row = a.ptab + (size_t)f * kMaxInputs;
#pragma unroll
                    for (uint32_t i = 0; i < I_FM_N; ++i)
                        if ((ramp_mask >> i) & 1u) P[i] = row[i];
                    if (ramp_mask & ((1u << I_FREQ) | (1u << I_OP3_RATIO) | (1u << I_OP2_RATIO) | (1u << I_OP1_RATIO))) OSC_FM_INCS();
                }
                if (ec.due(f)) {
                    const uint32_t j = ec.end_of_frame(f);
                    /* value edges are assigned before process_event_inputs() */
                    for (uint32_t k = ec.i; k < j; ++k) {
                        const Event e = ec.ev[k];
                        if (e.input_id != kEventGate) {
                            if (GENERAL) {
#pragma unroll
                                for (uint32_t i = 0; i < I_FM_N; ++i) P[i] = (e.input_id == i) ? e.scalar : P[i];   /* select form: keeps P in registers */
                            } else {
                                P[I_FREQ] = e.scalar;       /* host guarantees id == frequency */
                            }
                            a.pv[(size_t)e.input_id * n + v] = e.scalar;
                        }
                    }
                    if (GENERAL) {
                        if (adsr_derived_stale(d3, P[I_OP3_A], P[I_OP3_D], P[I_OP3_R])) adsr_derive_r(d3, P[I_OP3_A], P[I_OP3_D], P[I_OP3_R], sr);
                        if (adsr_derived_stale(d2, P[I_OP2_A], P[I_OP2_D], P[I_OP2_R])) adsr_derive_r(d2, P[I_OP2_A], P[I_OP2_D], P[I_OP2_R], sr);
                        if (adsr_derived_stale(d1, P[I_OP1_A], P[I_OP1_D], P[I_OP1_R])) adsr_derive_r(d1, P[I_OP1_A], P[I_OP1_D], P[I_OP1_R], sr);
                        if (adsr_derived_stale(df, P[I_F_A], P[I_F_D], P[I_F_R])) adsr_derive_r(df, P[I_F_A], P[I_F_D], P[I_F_R], sr);
                    }
                    for (uint32_t k = ec.i; k < j; ++k) {
                        const Event e = ec.ev[k];
                        if (e.input_id == kEventGate) {   /* gate -> env3, env2, env1, env_filter */
                            adsr_on_gate(e3, e.scalar, P[I_OP3_A], P[I_OP3_R], d3);
                            adsr_on_gate(e2, e.scalar, P[I_OP2_A], P[I_OP2_R], d2);
                            adsr_on_gate(e1, e.scalar, P[I_OP1_A], P[I_OP1_R], d1);
                            adsr_on_gate(ef, e.scalar, P[I_F_A], P[I_F_R], df);
                        }
                    }
                    OSC_FM_INCS();
                    if (!GENERAL) OSC_FM_REFRESH(true);
                    ec.advance(j);
                } else if (GENERAL && ramp_mask) {
                    if (adsr_derived_stale(d3, P[I_OP3_A], P[I_OP3_D], P[I_OP3_R])) adsr_derive_r(d3, P[I_OP3_A], P[I_OP3_D], P[I_OP3_R], sr);
                    if (adsr_derived_stale(d2, P[I_OP2_A], P[I_OP2_D], P[I_OP2_R])) adsr_derive_r(d2, P[I_OP2_A], P[I_OP2_D], P[I_OP2_R], sr);
                    if (adsr_derived_stale(d1, P[I_OP1_A], P[I_OP1_D], P[I_OP1_R])) adsr_derive_r(d1, P[I_OP1_A], P[I_OP1_D], P[I_OP1_R], sr);
                    if (adsr_derived_stale(df, P[I_F_A], P[I_F_D], P[I_F_R])) adsr_derive_r(df, P[I_F_A], P[I_F_D], P[I_F_R], sr);
                }

                /* ---- the walk (one frame) */
                const float env3 = adsr_process_t<GENERAL>(e3, h3, P[I_OP3_S], d3);
                const float env2 = adsr_process_t<GENERAL>(e2, h2, P[I_OP2_S], d2);
                const float env1 = adsr_process_t<GENERAL>(e1, h1, P[I_OP1_S], d1);
                const float envf = adsr_process_t<GENERAL>(ef, hf, P[I_F_S], df);
                const float feg = envf * P[I_ENV_AMT];                 /* filter_env_gain: Gain */
                const float cutoff = feg + P[I_CUTOFF];                /* cutoff_mod: AddValue  */
                const float o3 = fmop_process_k(op3, 0.0f, P[I_OP3_FB], env3, P[I_OP3_LEVEL], sk);
                const float mix = clampf(P[I_ROUTE], 0.0f, 1.0f);      /* op3_route: Crossfade  */
                const float xa = o3 * (1.0f - mix);
                const float xb = o3 * mix;
                const float o2 = fmop_process_k(op2, xa, P[I_OP2_FB], env2, P[I_OP2_LEVEL], sk);
                const float pm1 = o2 + xb;                             /* op1_mod_mixer: Mixer  */
                const float o1 = fmop_process_k(op1, pm1, 0.0f, env1, 1.0f, sk);
                const float lp = tpt_process(tp, o1, cutoff, P[I_RES], 0.0f, sr);
                y = lp * 0.3f;                                         /* output_gain: Gain(0.3) */
                if (a.voice_out != nullptr) a.voice_out[(size_t)f * n + v] = y;
            }
            tile[ff * kTileW + col] = y;
            ff += 1u;
        }
        tile_flush_mono(a, tile, isum, f0, fe);
    }

    {
        float *dp = staged ? tile : a.state;
        const uint32_t dn = staged ? kCta : n, dv = staged ? threadIdx.x : v;
        if (active) {
            adsr_store(e3, dp, FM_E3, dn, dv); adsr_store(e2, dp, FM_E2, dn, dv);
            adsr_store(e1, dp, FM_E1, dn, dv); adsr_store(ef, dp, FM_EF, dn, dv);
            dp[(size_t)(FM_OP3 + 0) * dn + dv] = op3.phase; dp[(size_t)(FM_OP3 + 1) * dn + dv] = op3.prev;
            dp[(size_t)(FM_OP2 + 0) * dn + dv] = op2.phase; dp[(size_t)(FM_OP2 + 1) * dn + dv] = op2.prev;
            dp[(size_t)(FM_OP1 + 0) * dn + dv] = op1.phase; dp[(size_t)(FM_OP1 + 1) * dn + dv] = op1.prev;
            tpt_store(tp, dp, FM_TPT, dn, dv);
        }
        if (staged) state_stage_out(tile, a.state, FM_NSTATE, n);
    }
}

}  // namespace oscen
