/*
 * tests/csrc/host_fm_runner.cpp — TEST INFRASTRUCTURE.  Runs one block of the HAND-WRITTEN fused FMVoice kernel
 * (oscen_b200/csrc/kernel_fm.cuh, the kernel bench.py times) on the host, both instantiations, against
 * tests/csrc/host_cuda/cuda_runtime.h — the same execution model as host_kernel_runner.cpp.  The state layout and the
 * input ids of the built-in FM graph equal those of its DSL text (tests/graphs/voice_fm3.graph), so tests/host_kernel.py
 * drives both through one HostEngine.
 */
#include "kernel_fm.cuh"

#include <thread>
#include <vector>

extern "C" int oscen_host_run_fm_block(int general, uint32_t n_voices, uint32_t vpi, float sample_rate, float *state, float *pv,
                                       const oscen::Event *events, uint32_t n_events, const float *ptab, uint32_t ramp_mask,
                                       uint32_t frames, float *voice_out, float *inst_out, float *mix_partial,
                                       const float *uniform_p /* [I_FM_N] */) {
    using namespace oscen;
    if (n_voices == 0 || vpi == 0 || kCta % vpi != 0 || n_voices % vpi != 0 || frames > kMaxBlock) return -1;
    BlockCtl ctl{};
    ctl.frames = frames; ctl.n_events = n_events; ctl.ramp_mask = ramp_mask; ctl.lcg_on = 0;
    KernelArgs a;
    memset(&a, 0, sizeof a);
    a.n_voices = n_voices; a.vpi = vpi; a.n_inst = n_voices / vpi; a.channels = 1;
    a.state = state; a.pv = pv; a.events = events; a.ptab = ptab; a.ctl = &ctl;
    a.voice_out = voice_out; a.inst_out = inst_out; a.mix_partial = mix_partial;
    /* engine.cu fill_args(): uniform values and the host-derived ADSR step data of the uniform envelopes */
    for (uint32_t i = 0; i < I_FM_N; ++i) a.u.p[i] = uniform_p[i];
    a.u.sr = sample_rate;
    const float *p = a.u.p;
    adsr_derive(a.u.env[0], p[I_OP3_A], p[I_OP3_D], p[I_OP3_R], a.u.sr);
    adsr_derive(a.u.env[1], p[I_OP2_A], p[I_OP2_D], p[I_OP2_R], a.u.sr);
    adsr_derive(a.u.env[2], p[I_OP1_A], p[I_OP1_D], p[I_OP1_R], a.u.sr);
    adsr_derive(a.u.env[3], p[I_F_A], p[I_F_D], p[I_F_R], a.u.sr);
    const uint32_t n_cta = (n_voices + kCta - 1) / kCta;
    blockDim.x = kCta; blockDim.y = blockDim.z = 1; gridDim.x = n_cta; gridDim.y = gridDim.z = 1;
    g_host_cta.n_threads = kCta;
    pthread_barrier_init(&g_host_cta.all, nullptr, kCta);
    for (uint32_t w = 0; w < kCta / 32u; ++w) pthread_barrier_init(&g_host_cta.warp[w], nullptr, 32);
    for (uint32_t b = 0; b < n_cta; ++b) {
        std::vector<std::thread> cta;
        for (uint32_t t = 0; t < kCta; ++t)
            cta.emplace_back([&, t, b]() {
                threadIdx.x = t; blockIdx.x = b;
                if (general) fm_voice_kernel<true>(a); else fm_voice_kernel<false>(a);
            });
        for (std::thread &th : cta) th.join();
    }
    pthread_barrier_destroy(&g_host_cta.all);
    for (uint32_t w = 0; w < kCta / 32u; ++w) pthread_barrier_destroy(&g_host_cta.warp[w]);
    return 0;
}
