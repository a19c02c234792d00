/*
 * kernels_common.cuh — pieces shared by every graph kernel: event lookup, per-voice value
 * inputs, the shared-memory output tile with in-order instance sums and the deterministic
 * mix-bus partials.
 */
#pragma once
#include "engine_types.h"

namespace oscen {

/* first event of `voice` in the (voice, frame, seq)-sorted list */
OSC_D uint32_t event_lower_bound(const Event *ev, uint32_t n, uint32_t voice) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (ev[mid].voice < voice) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct EventCursor {
    const Event *ev;
    uint32_t n, i, voice, next_frame;
    OSC_D void init(const Event *e, uint32_t n_events, uint32_t v) {
        ev = e; n = n_events; voice = v;
        i = n_events ? event_lower_bound(e, n_events, v) : 0u;
        next_frame = (i < n && ev[i].voice == v) ? ev[i].frame : 0xFFFFFFFFu;
    }
    OSC_D bool due(uint32_t f) const { return f == next_frame; }
    /* events of this voice at frame f: [i, end) ; advance() moves past them */
    OSC_D uint32_t end_of_frame(uint32_t f) const {
        uint32_t j = i;
        while (j < n && ev[j].voice == voice && ev[j].frame == f) ++j;
        return j;
    }
    OSC_D void advance(uint32_t j) {
        i = j;
        next_frame = (i < n && ev[i].voice == voice) ? ev[i].frame : 0xFFFFFFFFu;
    }
};

/* ---- TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) for the block-to-block state handoff:
   a CTA's slice of every SoA state row is 64 voices x 4 B = 256 contiguous bytes; one elected
   thread moves all rows HBM -> shared memory (completion on an mbarrier) at block start and
   shared memory -> HBM (bulk async-group) at block end.  Needs 16-byte aligned rows: n_voices % 4 == 0. */
OSC_D uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
/* inline PTX exists for the device compile only; the host-executed test harness (tests/csrc/host_cuda) never takes a TMA
   path (no descriptor, switch off), and must fault loudly if it ever did */
#if defined(__CUDACC__)
#define OSC_PTX(...) asm volatile(__VA_ARGS__)
#else
#define OSC_PTX(...) __trap()
#endif
OSC_D void mbar_init(unsigned long long *bar, uint32_t count) {
    OSC_PTX("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    OSC_PTX("fence.mbarrier_init.release.cluster;" ::: "memory");
}
OSC_D void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    OSC_PTX("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
OSC_D void mbar_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    do {
        OSC_PTX("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();          /* a lost completion must fault, not hang the GPU */
    } while (!ok);
}
OSC_D void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, unsigned long long *bar) {
    OSC_PTX("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
OSC_D void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes) {
    OSC_PTX("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
OSC_D void bulk_commit_wait() {
    OSC_PTX("cp.async.bulk.commit_group;" ::: "memory");
    OSC_PTX("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
OSC_D void fence_async_smem() { OSC_PTX("fence.proxy.async.shared::cta;" ::: "memory"); }

/* state handoff, load side: after the call `*sp / *sn / *sv` address this thread's state as
   sp[field * sn + sv] (shared-memory staging when the TMA path is usable, HBM otherwise) */
OSC_D bool state_stage_in(float *stage /* [n_fields][kCta], 128-B aligned */, unsigned long long *bar, const float *gstate,
                          uint32_t n_fields, uint32_t n, const float *&sp, uint32_t &sn, uint32_t &sv, uint32_t vs) {
    const uint32_t v0 = blockIdx.x * kCta;
    const bool tma_ok = (n % 4u == 0u);
    if (!tma_ok) { sp = gstate; sn = n; sv = vs; return false; }
    const uint32_t valid = min(kCta, n - v0);
    if (threadIdx.x == 0) mbar_init(bar, 1u);
    __syncthreads();
    /* one row per thread, all rows in flight at once; thread 0 posts the expected byte count (a
       complete_tx that lands first just drives the transaction count negative until then) */
    if (threadIdx.x == 0) mbar_expect_tx(bar, n_fields * valid * 4u);
    for (uint32_t f = threadIdx.x; f < n_fields; f += kCta) bulk_g2s(stage + f * kCta, gstate + (size_t)f * n + v0, valid * 4u, bar);
    mbar_wait(bar, 0u);
    sp = stage; sn = kCta; sv = min(threadIdx.x, valid - 1u);
    return true;
}
/* store side: every thread has written its column of `stage`; one thread streams the rows out */
OSC_D void state_stage_out(const float *stage, float *gstate, uint32_t n_fields, uint32_t n) {
    const uint32_t v0 = blockIdx.x * kCta;
    const uint32_t valid = min(kCta, n - v0);
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x < n_fields) {
        for (uint32_t f = threadIdx.x; f < n_fields; f += kCta) bulk_s2g(gstate + (size_t)f * n + v0, stage + f * kCta, valid * 4u);
        bulk_commit_wait();                                /* shared memory must outlive the reads */
    }
}

/* column of voice-in-CTA `t` in the padded tile row: one pad word per instance so that the
   in-order instance sum (stride vpi) is bank-conflict free */
OSC_D uint32_t tile_col(uint32_t t, uint32_t vpi) { return vpi > 1u ? t + t / vpi : t; }

/*
 * Flush one tile of per-voice outputs of a polyphonic graph whose voice sum runs through ONE instance-level Tremolo into a
 * Frame<2> output (examples/electric-piano/src/main.rs:88-96, tremolo.rs:41-61): per instance, the in-order voice sum, then
 * the tremolo frame by frame (its LFO phase is the instance's one state word), then
 *   inst_out[f][instance][2], mix_partial[cta][f][2] (instances folded in index order).
 * rate / depth: a graph value input (read at the instance's first voice, or from the ramp table) or the constructor's
 * literal.  tile: [kTileF][kTileW], isum: [kTileF][kCta][2].  All threads call.
 */
constexpr uint32_t kNoPostInput = 0xFFFFFFFFu;
OSC_D void tile_flush_tremolo(const KernelArgs &a, const float *tile, float *isum, uint32_t f0, uint32_t fe, float &trem_phase,
                              uint32_t rate_pid, float rate_cst, uint32_t depth_pid, float depth_cst) {
    const uint32_t vpi = a.vpi, n = a.n_voices;
    const uint32_t ipc = kCta / vpi;
    const uint32_t ramp_mask = a.ctl->ramp_mask;
    __syncthreads();
    if (threadIdx.x < ipc) {
        const uint32_t i = threadIdx.x;
        const uint32_t inst = blockIdx.x * ipc + i;
        const uint32_t c0 = tile_col(i * vpi, vpi);
        const uint32_t v_first = min(inst * vpi, n - 1u);           /* instance-level inputs: the instance's first voice */
        const float rate0 = rate_pid == kNoPostInput ? rate_cst : a.pv[(size_t)rate_pid * n + v_first];
        const float depth0 = depth_pid == kNoPostInput ? depth_cst : a.pv[(size_t)depth_pid * n + v_first];
        const bool rate_ramped = rate_pid != kNoPostInput && ((ramp_mask >> rate_pid) & 1u);
        const bool depth_ramped = depth_pid != kNoPostInput && ((ramp_mask >> depth_pid) & 1u);
        for (uint32_t ff = 0; ff < fe; ++ff) {
            const float rate = rate_ramped ? a.ptab[(size_t)(f0 + ff) * kMaxInputs + rate_pid] : rate0;
            const float depth = depth_ramped ? a.ptab[(size_t)(f0 + ff) * kMaxInputs + depth_pid] : depth0;
            const float *row = tile + ff * kTileW;
            float s = 0.0f;
            for (uint32_t j = 0; j < vpi; ++j) s = s + row[c0 + j];
            float l, r;
            tremolo_process(trem_phase, s, rate, depth, a.u.sr, l, r);
            isum[(ff * kCta + i) * 2 + 0] = l; isum[(ff * kCta + i) * 2 + 1] = r;
            if (a.inst_out != nullptr && inst < a.n_inst) {
                a.inst_out[((size_t)(f0 + ff) * a.n_inst + inst) * 2 + 0] = l;
                a.inst_out[((size_t)(f0 + ff) * a.n_inst + inst) * 2 + 1] = r;
            }
        }
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < fe * 2u; t += kCta) {       /* mix partial: in-order over the CTA's instances */
        const uint32_t ff = t >> 1, c = t & 1u;
        float s = 0.0f;
        for (uint32_t i = 0; i < ipc; ++i) s = s + isum[(ff * kCta + i) * 2 + c];
        a.mix_partial[((size_t)blockIdx.x * kMaxBlock + f0 + ff) * 2 + c] = s;
    }
    __syncthreads();
}

/*
 * Flush one tile of per-voice outputs (mono):
 *   inst_out[f][instance] = sum over the instance's voices in index order (sequential f32
 *     fold == `.iter().map(|n| n.audio_out).sum()`, codegen/emit_node.rs:463-466)
 *   mix_partial[cta][f]   = fixed-shape sum of this CTA's instance outputs
 * tile: [kTileF][kTileW] voice outputs, isum: [kTileF][kCta] scratch.  All threads call.
 */
OSC_D void tile_flush_mono(const KernelArgs &a, const float *tile, float *isum, uint32_t f0, uint32_t fe) {
    const uint32_t vpi = a.vpi;
    const uint32_t ipc = kCta / vpi;                       /* instances per CTA */
    const uint32_t inst0 = blockIdx.x * ipc;
    __syncthreads();
    const uint32_t ipc_shift = 31u - (uint32_t)__clz((int)ipc);          /* ipc is a power of two */
    for (uint32_t t = threadIdx.x; t < fe * ipc; t += kCta) {
        const uint32_t ff = t >> ipc_shift, i = t & (ipc - 1u);
        const float *row = tile + ff * kTileW;
        float s = 0.0f;
        const uint32_t c0 = tile_col(i * vpi, vpi);
        for (uint32_t j = 0; j < vpi; ++j) s = s + row[c0 + j];
        /* rows of isum are rotated by their frame index: the fold below has the lanes of a warp on DIFFERENT frames reading
           the same instance index, which in an unrotated [frame][kCta] array is one bank for all of them (a 32-way conflict
           per load at vpi = 1: 2.6e8 conflict cycles per block in the TwinPeaks profile) */
        isum[ff * kCta + ((i + ff) & (kCta - 1u))] = s;
        if (a.inst_out != nullptr && inst0 + i < a.n_inst)
            a.inst_out[(size_t)(f0 + ff) * a.n_inst + inst0 + i] = s;
    }
    __syncthreads();
    /* mix partial: kCta/kTileF threads per frame, each folds its share of the instances in order,
       then the shares are folded in lane order - fixed shape, run-to-run deterministic */
    {
        constexpr uint32_t TPF = kCta / kTileF;
        const uint32_t ff = threadIdx.x / TPF, qd = threadIdx.x % TPF;
        float s = 0.0f;
        if (ff < fe) {
            const uint32_t per = (ipc + TPF - 1u) / TPF;
            const uint32_t b = qd * per, e = min(b + per, ipc);
            for (uint32_t i = b; i < e; ++i) s = s + isum[ff * kCta + ((i + ff) & (kCta - 1u))];
        }
        float tot = s;
#pragma unroll
        for (uint32_t k = 1; k < TPF; ++k) tot = tot + __shfl_down_sync(0xffffffffu, s, k);
        if (qd == 0 && ff < fe)
            a.mix_partial[(size_t)blockIdx.x * kMaxBlock + f0 + ff] = tot;
    }
    __syncthreads();
}

/* final mix bus: mix[f][c] = fixed-shape sum over CTAs of mix_partial[cta][f][c] */
__global__ void mix_reduce_kernel(const float *__restrict__ partial, float *__restrict__ mix,
                                  uint32_t n_cta, uint32_t channels, const BlockCtl *ctl) {
    __shared__ float red[256];
    const uint32_t fc = blockIdx.x;                        /* frame*channels + c */
    if (fc >= ctl->frames * channels) return;
    float s = 0.0f;
    for (uint32_t c = threadIdx.x; c < n_cta; c += 256) s = s + partial[(size_t)c * kMaxBlock * channels + fc];
    red[threadIdx.x] = s;
    __syncthreads();
    for (uint32_t w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] = red[threadIdx.x] + red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) mix[fc] = red[0];
}

/*
 * Mix bus over NVLink peer memory: compute step (fixed-shape reduction of this rank's CTA
 * partials, one output element per CTA) fused with the exchange: each CTA stores its element into
 * every rank's buffer; the last CTA to finish raises the block's flag on every rank, waits for the
 * peers' flags and folds the ranks' contributions in rank order.
 *   xchg layout per rank: data[2 parity][kMaxRanks][1024] floats, then flags[2][kMaxRanks] u32.
 *   peers[r] = rank r's xchg buffer mapped into this process (peers[rank] = local).
 * A rank can only be two blocks ahead of a peer after that peer finished reading the older
 * block (it needs the peer's flag of the block in between), so two parities suffice.
 */
constexpr uint32_t kMaxRanks = 16;
constexpr uint32_t kXchgFloats = 2u * kMaxRanks * 1024u;
constexpr size_t kXchgBytes = (size_t)kXchgFloats * sizeof(float) + 2u * kMaxRanks * sizeof(uint32_t);
struct PeerTable { float *peer[kMaxRanks]; };

__global__ void __launch_bounds__(256) mix_exchange_kernel(const float *__restrict__ partial, float *__restrict__ mix,
                                                             uint32_t n_cta, uint32_t channels, const BlockCtl *ctl,
                                                             const __grid_constant__ PeerTable peers, uint32_t rank,
                                                             uint32_t n_ranks, uint32_t *seq_word, uint32_t *counter,
                                                             uint32_t *error_flag) {
    __shared__ float red[256];
    __shared__ uint32_t is_last;
    const uint32_t count = ctl->frames * channels;
    /* the block's sequence number lives in device memory (so the launch is identical every block and can sit in a
       captured CUDA graph); the last CTA to finish advances it, after every other CTA of this launch has read it */
    const uint32_t seq = *seq_word + 1u;
    const uint32_t parity = seq & 1u;
    const uint32_t fc = blockIdx.x;                        /* one output element per CTA, as mix_reduce_kernel */
    /* 1. compute: fixed-shape reduction of this rank's CTA partials for element fc */
    if (fc < count) {
        float s = 0.0f;
        for (uint32_t c = threadIdx.x; c < n_cta; c += 256) s = s + partial[(size_t)c * kMaxBlock * channels + fc];
        red[threadIdx.x] = s;
        __syncthreads();
        for (uint32_t w = 128; w > 0; w >>= 1) {
            if (threadIdx.x < w) red[threadIdx.x] = red[threadIdx.x] + red[threadIdx.x + w];
            __syncthreads();
        }
        /* 2. exchange: the element goes straight into every rank's buffer (NVLink peer stores) */
        if (threadIdx.x < n_ranks)
            peers.peer[threadIdx.x][((size_t)parity * kMaxRanks + rank) * 1024u + fc] = red[0];
    }
    /* 3. the last CTA to finish publishes this block's flag and completes the collective */
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();                            /* this CTA's peer stores before the count */
        const uint32_t prev = atomicAdd(counter, 1u);
        is_last = (prev == gridDim.x - 1u) ? 1u : 0u;
        if (is_last) { *counter = 0u; __threadfence_system(); }
    }
    __syncthreads();
    if (!is_last) return;
    if (threadIdx.x < n_ranks) {
        uint32_t *pf = reinterpret_cast<uint32_t *>(peers.peer[threadIdx.x] + kXchgFloats) + parity * kMaxRanks + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pf), "r"(seq) : "memory");
        const uint32_t *mf = reinterpret_cast<const uint32_t *>(peers.peer[rank] + kXchgFloats) + parity * kMaxRanks + threadIdx.x;
        uint32_t got = 0, spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(mf) : "memory");
            if (got == seq) break;
            __nanosleep(32);
        } while (++spins < (1u << 24));                     /* bounded: a dead peer must not hang the GPU */
        if (got != seq) *error_flag = 1u;
    }
    __syncthreads();
    /* 4. fold the ranks' contributions in rank order (identical bits on every rank) */
    const float *mine = peers.peer[rank] + (size_t)parity * kMaxRanks * 1024u;
    for (uint32_t k = threadIdx.x; k < count; k += 256) {
        float s = 0.0f;
        for (uint32_t r = 0; r < n_ranks; ++r) s = s + __ldcv(mine + (size_t)r * 1024u + k);
        mix[k] = s;
    }
    if (threadIdx.x == 0) *seq_word = seq;
}

__global__ void fill_u32_kernel(uint32_t *dst, size_t n, uint32_t value) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = value;
}
/* state[f][v] = init[f] */
__global__ void init_state_kernel(uint32_t *state, const uint32_t *init, uint32_t n_fields, uint32_t n_voices) {
    const size_t n = (size_t)n_fields * n_voices;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        state[i] = init[i / n_voices];
}

}  // namespace oscen
