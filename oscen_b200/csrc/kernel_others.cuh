/*
 * kernel_others.cuh — fused one-thread-per-voice kernels for the other reference graphs on
 * the hot path: TwinPeaksGraph, SatGraph_{1x,4x}, README Synth (+ADSR), ElectricPiano.
 * Same structure as kernel_fm.cuh; value inputs are always read per voice (few inputs).
 */
#pragma once
#include "kernels_common.cuh"
#include "kernel_fm.cuh"

namespace oscen {

/* shared prologue for the small graphs: load NP per-voice inputs, apply ramps / value events */
template <uint32_t NP>
struct SmallParams {
    float P[NP];
    OSC_D void load(const KernelArgs &a, uint32_t vs) {
#pragma unroll
        for (uint32_t i = 0; i < NP; ++i) P[i] = a.pv[(size_t)i * a.n_voices + vs];
    }
    OSC_D void ramps(const KernelArgs &a, uint32_t ramp_mask, uint32_t f) {
        if (ramp_mask) {
            const float *row = a.ptab + (size_t)f * kMaxInputs;
#pragma unroll
            for (uint32_t i = 0; i < NP; ++i) if ((ramp_mask >> i) & 1u) P[i] = row[i];
        }
    }
    OSC_D void value_event(const KernelArgs &a, const Event &e, uint32_t v) {
#pragma unroll
        for (uint32_t i = 0; i < NP; ++i) if (e.input_id == i) P[i] = e.scalar;
        if (e.input_id < NP) a.pv[(size_t)e.input_id * a.n_voices + v] = e.scalar;
    }
};

/* ============================================================== TwinPeaksGraph ==== */
/* examples/nih-twin-peaks/src/lib.rs:15-48: audio_in -> filter_a, filter_b; out = a - b */
__global__ void __launch_bounds__(kCta) twin_peaks_kernel(const __grid_constant__ KernelArgs a) {
    __shared__ float tile[kTileF * kTileW];
    __shared__ float isum[kTileF * kCta];
    const uint32_t n = a.n_voices;
    const uint32_t v = blockIdx.x * kCta + threadIdx.x;
    const bool active = v < n;
    const uint32_t vs = active ? v : 0u;
    const uint32_t frames = a.ctl->frames, ramp_mask = a.ctl->ramp_mask;
    const bool lcg_on = a.ctl->lcg_on != 0u;
    const float sr = a.u.sr;
    const uint32_t col = tile_col(threadIdx.x, a.vpi);

    SmallParams<3> prm; prm.load(a, vs);
    Lp18State fa, fb;
    lp18_load(fa, a.state, TW_A, n, vs); lp18_load(fb, a.state, TW_B, n, vs);
    unsigned long long lcg = (unsigned long long)__float_as_uint(a.state[(size_t)(TW_LCG + 0) * n + vs]) |
                             ((unsigned long long)__float_as_uint(a.state[(size_t)(TW_LCG + 1) * n + vs]) << 32);
    EventCursor ec; ec.init(a.events, active ? a.ctl->n_events : 0u, vs);

    for (uint32_t f0 = 0; f0 < frames; f0 += kTileF) {
        const uint32_t fe = min(kTileF, frames - f0);
        /* stage this tile's input column in shared memory: 32 independent coalesced loads in flight,
           then a compact (not unrolled) compute loop — the two LP18 bodies are ~400 instructions */
        if (active && !lcg_on && a.stream_in != nullptr) {
#pragma unroll 8
            for (uint32_t ff = 0; ff < fe; ++ff) isum[ff * kCta + threadIdx.x] = __ldcs(a.stream_in + (size_t)(f0 + ff) * n + v);
        }
#pragma unroll 1
        for (uint32_t ff = 0; ff < fe; ++ff) {
            const uint32_t f = f0 + ff;
            float y = 0.0f;
            if (active) {
                prm.ramps(a, ramp_mask, f);
                if (ec.due(f)) {
                    const uint32_t j = ec.end_of_frame(f);
                    for (uint32_t q = ec.i; q < j; ++q) prm.value_event(a, ec.ev[q], v);
                    ec.advance(j);
                }
                const float x = lcg_on ? lcg_next(lcg) : ((a.stream_in != nullptr) ? isum[ff * kCta + threadIdx.x] : 0.0f);
                const float ya = lp18_process(fa, x, prm.P[0], 0.0f, prm.P[2], sr);
                const float yb = lp18_process(fb, x, prm.P[1], 0.0f, prm.P[2], sr);
                y = ya - yb;
                if (a.voice_out != nullptr) a.voice_out[(size_t)f * n + v] = y;
            }
            tile[ff * kTileW + col] = y;
        }
        tile_flush_mono(a, tile, isum, f0, fe);
    }
    if (active) {
        lp18_store(fa, a.state, TW_A, n, v); lp18_store(fb, a.state, TW_B, n, v);
        a.state[(size_t)(TW_LCG + 0) * n + v] = __uint_as_float((uint32_t)(lcg & 0xffffffffull));
        a.state[(size_t)(TW_LCG + 1) * n + v] = __uint_as_float((uint32_t)(lcg >> 32));
    }
}

/* ============================================================== SatGraph_{1x,4x} === */
/* examples/oversampled-saturator/src/main.rs:64-80; multirate body codegen/emit_frame.rs:114-176 */
template <int FACTOR>
__global__ void __launch_bounds__(kCta) sat_kernel(const __grid_constant__ KernelArgs a) {
    __shared__ float tile[kTileF * kTileW];
    extern __shared__ float dyn[];                         /* isum [32][kCta] | h1 [48][kCta] | h2 [48][kCta] */
    float *isum = dyn;
    float *h1 = dyn + kTileF * kCta + threadIdx.x;
    float *h2 = h1 + 48 * kCta;
    const uint32_t n = a.n_voices;
    const uint32_t v = blockIdx.x * kCta + threadIdx.x;
    const bool active = v < n;
    const uint32_t vs = active ? v : 0u;
    const uint32_t frames = a.ctl->frames, ramp_mask = a.ctl->ramp_mask;
    const float sr_inner = a.u.sr * (float)FACTOR;         /* `* N` nodes: sample_rate * N */
    const uint32_t col = tile_col(threadIdx.x, a.vpi);

    SmallParams<2> prm; prm.load(a, vs);
    float phase = a.state[(size_t)SAT_PHASE * n + vs];
    uint32_t head1 = 23u, head2 = 23u;                     /* chronological order, newest last */
    if (FACTOR == 4) {
        for (uint32_t i = 0; i < 24; ++i) {
            const float t1 = a.state[(size_t)(SAT_H1 + i) * n + vs], t2 = a.state[(size_t)(SAT_H2 + i) * n + vs];
            h1[i * kCta] = t1; h1[(i + 24) * kCta] = t1;
            h2[i * kCta] = t2; h2[(i + 24) * kCta] = t2;
        }
    }
    EventCursor ec; ec.init(a.events, active ? a.ctl->n_events : 0u, vs);

    for (uint32_t f0 = 0; f0 < frames; f0 += kTileF) {
        const uint32_t fe = min(kTileF, frames - f0);
        for (uint32_t ff = 0; ff < fe; ++ff) {
            const uint32_t f = f0 + ff;
            float y = 0.0f;
            if (active) {
                prm.ramps(a, ramp_mask, f);
                if (ec.due(f)) {
                    const uint32_t j = ec.end_of_frame(f);
                    for (uint32_t q = ec.i; q < j; ++q) prm.value_event(a, ec.ev[q], v);
                    ec.advance(j);
                }
                if (FACTOR == 1) {
                    const float o = polyblep_process<WAVE_SAW>(phase, 0.0f, prm.P[0], 0.0f, prm.P[1], 0.5f, sr_inner);
                    y = hardclip_process(o);
                } else {
                    float buf[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float o = polyblep_process<WAVE_SAW>(phase, 0.0f, prm.P[0], 0.0f, prm.P[1], 0.5f, sr_inner);
                        buf[i] = hardclip_process(o);
                    }
                    /* SincDownFir<4>::downsample: stage 0 twice, stage 1 once (sinc_fir.rs:233-249) */
                    const float s0 = hb_down_step(h1, kCta, head1, buf[0], buf[1]);
                    const float s1 = hb_down_step(h1, kCta, head1, buf[2], buf[3]);
                    y = hb_down_step(h2, kCta, head2, s0, s1);
                }
                if (a.voice_out != nullptr) a.voice_out[(size_t)f * n + v] = y;
            }
            tile[ff * kTileW + col] = y;
        }
        tile_flush_mono(a, tile, isum, f0, fe);
    }
    if (active) {
        a.state[(size_t)SAT_PHASE * n + v] = phase;
        if (FACTOR == 4) {
            for (uint32_t i = 0; i < 24; ++i) {            /* store oldest -> newest */
                a.state[(size_t)(SAT_H1 + i) * n + v] = h1[((head1 + 1u + i) % 24u) * kCta];
                a.state[(size_t)(SAT_H2 + i) * n + v] = h2[((head2 + 1u + i) % 24u) * kCta];
            }
        }
    }
}

/* ===================================================== README Synth (+ ADSR vca) ==== */
/* README.md:24-53; ADSR -> Gain wiring as oscen-lib/benches/static_vs_runtime.rs:57-59 */
template <bool WITH_ADSR>
__global__ void __launch_bounds__(kCta) synth_kernel(const __grid_constant__ KernelArgs a) {
    __shared__ float tile[kTileF * kTileW];
    __shared__ float isum[kTileF * kCta];
    const uint32_t n = a.n_voices;
    const uint32_t v = blockIdx.x * kCta + threadIdx.x;
    const bool active = v < n;
    const uint32_t vs = active ? v : 0u;
    const uint32_t frames = a.ctl->frames, ramp_mask = a.ctl->ramp_mask;
    const float sr = a.u.sr;
    const uint32_t col = tile_col(threadIdx.x, a.vpi);

    SmallParams<8> prm; prm.load(a, vs);
    float ph_mod = a.state[(size_t)SY_MOD * n + vs], ph_car = a.state[(size_t)SY_CAR * n + vs];
    TptState tp; tpt_load(tp, a.state, SY_TPT, n, vs);
    AdsrState env; adsr_load(env, a.state, SY_ENV, n, vs);
    AdsrDerived d;
    adsr_derive_r(d, prm.P[4], prm.P[5], prm.P[7], sr);
    EventCursor ec; ec.init(a.events, active ? a.ctl->n_events : 0u, vs);
    adsr_begin_block(env, d, ec.due(0u));

    for (uint32_t f0 = 0; f0 < frames; f0 += kTileF) {
        const uint32_t fe = min(kTileF, frames - f0);
        for (uint32_t ff = 0; ff < fe; ++ff) {
            const uint32_t f = f0 + ff;
            float y = 0.0f;
            if (active) {
                prm.ramps(a, ramp_mask, f);
                uint32_t j = ec.i;
                const bool due = ec.due(f);
                if (due) {
                    j = ec.end_of_frame(f);
                    for (uint32_t q = ec.i; q < j; ++q) if (ec.ev[q].input_id != kEventGate) prm.value_event(a, ec.ev[q], v);
                }
                if (WITH_ADSR && adsr_derived_stale(d, prm.P[4], prm.P[5], prm.P[7])) adsr_derive_r(d, prm.P[4], prm.P[5], prm.P[7], sr);
                if (due) {
                    if (WITH_ADSR)
                        for (uint32_t q = ec.i; q < j; ++q) if (ec.ev[q].input_id == kEventGate)
                            adsr_on_gate(env, ec.ev[q].scalar, prm.P[4], prm.P[7], d);
                    ec.advance(j);
                }
                const float m = polyblep_process<WAVE_SINE>(ph_mod, 0.0f, prm.P[0], 0.0f, prm.P[1], 0.5f, sr);
                const float c = polyblep_process<WAVE_SAW>(ph_car, 0.0f, prm.P[2], m, 0.5f, 0.5f, sr);
                const float lp = tpt_process(tp, c, prm.P[3], 0.707f, 0.0f, sr);
                if (WITH_ADSR) {
                    const float g = adsr_process(env, prm.P[6], d);
                    y = lp * g;                                  /* vca: Gain */
                } else {
                    y = lp;
                }
                if (a.voice_out != nullptr) a.voice_out[(size_t)f * n + v] = y;
            }
            tile[ff * kTileW + col] = y;
        }
        tile_flush_mono(a, tile, isum, f0, fe);
    }
    if (active) {
        a.state[(size_t)SY_MOD * n + v] = ph_mod; a.state[(size_t)SY_CAR * n + v] = ph_car;
        tpt_store(tp, a.state, SY_TPT, n, v);
        adsr_store(env, a.state, SY_ENV, n, v);
    }
}

/* ================================================================ ElectricPiano ==== */
/* examples/electric-piano/src/electric_piano_voice.rs (AmplitudeSource, OscillatorBank,
   ElectricPianoVoiceNode), tremolo.rs, main.rs:33-97 (16 voices -> sum -> Tremolo -> Frame<2>).
   Per-voice state (259 words) lives in a [field][thread] shared-memory tile for the block. */
__constant__ float kEpVel0[32] = { 0.02f, 0.05f };
__constant__ float kEpVel127[32] = {
    0.150869f, 0.385766f, 0.215543f, 0.117811f, 0.100411f, 0.0128637f, 0.0288844f, 0.00243388f,
    0.00963092f, 0.0035634f, 0.00256945f, 0.00184799f, 0.000399878f, 0.000660576f, 3.00995e-05f,
    0.00021866f, 9.33705e-05f, 0.000177973f, 0.0002545f, 0.000323602f, 0.000779045f, 0.000116569f,
    0.000772873f, 0.000364486f, 0.000248027f, 0.00018236f, 3.27292e-05f, 6.64988e-05f,
    0.0f, 0.0f, 0.0f, 0.0f };

#define EPS(field, i) S[(size_t)((field) + (i)) * n_stride]
/* AmplitudeSource::on_gate for the harmonics [h0, h0+8) of one voice (registers) */
OSC_D void ep_gate_chunk(float *S, size_t n_stride, uint32_t h0, float scalar, const float *P, float (&cur)[8], float (&ore)[8],
                         float (&oim)[8], bool &released, uint32_t &step) {
    if (scalar > 0.0f) {
        const float velocity = scalar;
        const float note = 60.0f;                                   /* note_pitch is never updated */
        const float base_decay_rate = (100.0f - P[3]) / 40000.0f;
        const float harmonic_scaling = 1.0f - ((100.0f - P[4]) / 200000.0f);
        const float scaling_multiplier = (48.0f - note) / 12.0f;
        const float key_scaling_factor = scaling_multiplier * (P[5] * 0.02f);
        const float adjusted = (key_scaling_factor > 0.0f)
            ? 1.0f - (base_decay_rate / (1.0f + key_scaling_factor))
            : 1.0f - (base_decay_rate * (1.0f - key_scaling_factor));
        float scaling = 1.0f;                                       /* scaling *= harmonic_scaling, i times */
        for (uint32_t i = 0; i < h0; ++i) scaling *= harmonic_scaling;
        const float rv = 0.999f - ((100.0f - P[6]) / 1000.0f);
        float bs = -0.2f + (0.8f * (P[1] * 0.01f));
        bs += velocity * P[2] * 0.01f * 0.5f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            EPS(EP_DEC, h0 + i) = adjusted * scaling; scaling *= harmonic_scaling;
            EPS(EP_REL, h0 + i) = rv;
            float amp = (kEpVel127[h0 + i] * velocity) + (kEpVel0[h0 + i] * (1.0f - velocity));
            amp *= 1.0f + bs * (float)(h0 + i);
            cur[i] = amp;
            ore[i] = 1.0f; oim[i] = 0.0f;                           /* OscillatorBank::on_gate */
        }
        released = false; step = 0u;
    } else {
        released = true; step = 0u;
    }
}

/*
 * One thread = one voice; the 259-word state sits in a [field][thread] shared-memory tile for the
 * block.  The 32 harmonics are processed in 4 chunks of 8 held in REGISTERS across a 32-frame
 * tile (chunk-major inside the tile): per frame and chunk, 8 lerps + 8 complex rotations run out
 * of registers, and the running output sum is carried through a per-thread accumulator row so the
 * f32 additions happen in exactly the reference order sum += osc[i].imag * amp[i], i = 0..31.
 * The per-frame control flow (interpolation step, release flag, gate / value events, multiplier
 * refresh) is replayed identically by every chunk pass.
 */
__global__ void __launch_bounds__(kCta, 6) epiano_kernel(const __grid_constant__ KernelArgs a) {
    extern __shared__ float dyn[];          /* acc [kTileF][kCta] | tile [kTileF][kTileW] | isum [kTileF*kCta*2] */
    float *acc = dyn + threadIdx.x;
    float *tile = dyn + kTileF * kCta;
    float *isum = tile + kTileF * kTileW;
    const uint32_t n = a.n_voices;
    const uint32_t v = blockIdx.x * kCta + threadIdx.x;
    const bool active = v < n;
    const uint32_t vs = active ? v : 0u;
    const uint32_t frames = a.ctl->frames, ramp_mask = a.ctl->ramp_mask;
    const float sr = a.u.sr;
    const uint32_t vpi = a.vpi;
    const uint32_t col = tile_col(threadIdx.x, vpi);

    SmallParams<9> prm; prm.load(a, vs);
    /* the 259-word voice state stays in HBM/L2 as SoA rows; each chunk pass pulls its 48 words into
       registers per 32-frame tile (coalesced: a warp reads 128 contiguous bytes per field) */
    float *S = a.state + vs;
    const size_t n_stride = n;
    EventCursor ec; ec.init(a.events, active ? a.ctl->n_events : 0u, vs);
    const uint32_t ipc = kCta / vpi;
    const uint32_t my_inst = blockIdx.x * ipc + threadIdx.x;
    float trem_phase = (threadIdx.x < ipc && my_inst < a.n_inst) ? a.inst_state[my_inst] : 0.0f;

    for (uint32_t f0 = 0; f0 < frames; f0 += kTileF) {
        const uint32_t fe = min(kTileF, frames - f0);
        if (active) {
            const uint32_t step0 = (uint32_t)EPS(EP_STEP, 0);
            const bool rel0 = EPS(EP_RELEASED, 0) != 0.0f;
            const float lastf0 = EPS(EP_LASTF, 0);
            const SmallParams<9> prm0 = prm;
            const EventCursor ec0 = ec;
            uint32_t step = step0; bool released = rel0; float lastf = lastf0;
            for (uint32_t h0 = 0; h0 < 32u; h0 += 8u) {
                float cur[8], tgt[8], ore[8], oim[8], mre[8], mim[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    cur[i] = EPS(EP_CUR, h0 + i); tgt[i] = EPS(EP_TGT, h0 + i);
                    ore[i] = EPS(EP_ORE, h0 + i); oim[i] = EPS(EP_OIM, h0 + i);
                    mre[i] = EPS(EP_MRE, h0 + i); mim[i] = EPS(EP_MIM, h0 + i);
                }
                step = step0; released = rel0; lastf = lastf0; prm = prm0; ec = ec0;
                for (uint32_t ff = 0; ff < fe; ++ff) {
                    const uint32_t f = f0 + ff;
                    prm.ramps(a, ramp_mask, f);
                    if (ec.due(f)) {
                        const uint32_t j = ec.end_of_frame(f);
                        for (uint32_t q = ec.i; q < j; ++q) if (ec.ev[q].input_id != kEventGate) prm.value_event(a, ec.ev[q], v);
                        for (uint32_t q = ec.i; q < j; ++q) if (ec.ev[q].input_id == kEventGate)
                            ep_gate_chunk(S, n_stride, h0, ec.ev[q].scalar, prm.P, cur, ore, oim, released, step);
                        ec.advance(j);
                    }
                    /* AmplitudeSource::process, electric_piano_voice.rs:322-351 */
                    if (step == 0u) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) tgt[i] = cur[i] * (released ? EPS(EP_REL, h0 + i) : EPS(EP_DEC, h0 + i));
                    }
                    if (step < 64u) {
                        const float t = (float)(step + 1u) / 64.0f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) cur[i] = cur[i] * (1.0f - t) + tgt[i] * t;
                        step += 1u;
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) cur[i] = tgt[i];
                        step = 0u;
                    }
                    /* OscillatorBank::process, electric_piano_voice.rs:127-170 */
                    const float freq = prm.P[0];
                    if (freq > 0.0f && !(fabsf(lastf - freq) < 0.01f)) {
                        lastf = freq;
                        const float nyquist = sr * 0.5f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float hf = freq * (float)(h0 + i + 1);
                            if (hf < nyquist) {
                                const float angle = 2.0f * kPi * hf / sr;
                                mre[i] = om_cosf_u(angle); mim[i] = om_sinf_u(angle);
                            } else { mre[i] = 1.0f; mim[i] = 0.0f; }
                            ore[i] = 1.0f; oim[i] = 0.0f;
                        }
                    }
                    float sum = (h0 == 0u) ? 0.0f : acc[ff * kCta];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float nr = ore[i] * mre[i] - oim[i] * mim[i];
                        const float ni = ore[i] * mim[i] + oim[i] * mre[i];
                        ore[i] = nr; oim[i] = ni;
                        sum += ni * cur[i];
                    }
                    acc[ff * kCta] = sum;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    EPS(EP_CUR, h0 + i) = cur[i]; EPS(EP_TGT, h0 + i) = tgt[i];
                    EPS(EP_ORE, h0 + i) = ore[i]; EPS(EP_OIM, h0 + i) = oim[i];
                    EPS(EP_MRE, h0 + i) = mre[i]; EPS(EP_MIM, h0 + i) = mim[i];
                }
            }
            EPS(EP_STEP, 0) = (float)step; EPS(EP_RELEASED, 0) = released ? 1.0f : 0.0f; EPS(EP_LASTF, 0) = lastf;
        }
        for (uint32_t ff = 0; ff < fe; ++ff) {
            const float y = active ? acc[ff * kCta] * 3.0f : 0.0f;       /* output scaling */
            if (active && a.voice_out != nullptr) a.voice_out[(size_t)(f0 + ff) * n + v] = y;
            tile[ff * kTileW + col] = y;
        }
        /* ---- flush: instance sum (index order) -> Tremolo (sequential in time) -> Frame<2> */
        __syncthreads();
        if (threadIdx.x < ipc) {
            const uint32_t i = threadIdx.x;
            const uint32_t c0 = tile_col(i * vpi, vpi);
            const uint32_t v_first = min((blockIdx.x * ipc + i) * vpi, n - 1u);   /* instance-level inputs: first voice */
            const float depth0 = a.pv[(size_t)7 * n + v_first];
            const float rate0 = a.pv[(size_t)8 * n + v_first];
            for (uint32_t ff = 0; ff < fe; ++ff) {
                const float depth = ((ramp_mask >> 7) & 1u) ? a.ptab[(size_t)(f0 + ff) * kMaxInputs + 7] : depth0;
                const float rate = ((ramp_mask >> 8) & 1u) ? a.ptab[(size_t)(f0 + ff) * kMaxInputs + 8] : rate0;
                const float *row = tile + ff * kTileW;
                float s = 0.0f;
                for (uint32_t j = 0; j < vpi; ++j) s = s + row[c0 + j];
                float l, r;
                tremolo_process(trem_phase, s, rate, depth, sr, l, r);
                isum[(ff * kCta + i) * 2 + 0] = l; isum[(ff * kCta + i) * 2 + 1] = r;
                if (a.inst_out != nullptr && my_inst < a.n_inst) {
                    a.inst_out[((size_t)(f0 + ff) * a.n_inst + my_inst) * 2 + 0] = l;
                    a.inst_out[((size_t)(f0 + ff) * a.n_inst + my_inst) * 2 + 1] = r;
                }
            }
        }
        __syncthreads();
        for (uint32_t t = threadIdx.x; t < fe * 2u; t += kCta) {   /* mix partial: in-order over the CTA's instances */
            const uint32_t ff = t >> 1, c = t & 1u;
            float s = 0.0f;
            for (uint32_t i = 0; i < ipc; ++i) s = s + isum[(ff * kCta + i) * 2 + c];
            a.mix_partial[((size_t)blockIdx.x * kMaxBlock + f0 + ff) * 2 + c] = s;
        }
        __syncthreads();
    }
    if (threadIdx.x < ipc && my_inst < a.n_inst) a.inst_state[my_inst] = trem_phase;
#undef EPS
}

}  // namespace oscen
