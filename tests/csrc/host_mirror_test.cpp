// C++ host driving the C ABI through the generated-graph mirror (oscen_b200/host/graph.hpp),
// the way examples/pivot/src/main.rs:599-725 smoke-tests a graph: note on -> audible output.
// Exit code 0 = ok, 77 = no GPU (compile/link/ABI check only).
#include <cmath>
#include <cstdio>
#include "../../oscen_b200/host/graph.hpp"

int main() {
    if (oscen_b200_device_count() == 0) {
        oscen_b200_config cfg{}; cfg.graph_kind = 0; cfg.n_voices = 8; cfg.voices_per_instance = 8; cfg.sample_rate = 48000.f;
        oscen_b200_engine *e = nullptr;
        int rc = oscen_b200_create(&cfg, &e);
        std::printf("no GPU: create rc=%d (%s)\n", rc, oscen_b200_last_error(nullptr));
        return rc == OSCEN_B200_E_NO_DEVICE ? 77 : 1;
    }
    oscen_b200::FMGraphBank synth(4, 48000.0f);
    const uint8_t note_on[3] = {0x90, 60, 100};
    for (uint32_t i = 0; i < 4; ++i) synth.push_midi(i, 0, note_on);
    synth.set_filter_cutoff(3000.0f);
    float peak = 0.0f;
    for (int b = 0; b < 16; ++b) {
        synth.process_block(512);
        for (float x : synth.out_block()) peak = std::fmax(peak, std::fabs(x));
    }
    std::printf("peak = %g\n", peak);
    return peak > 1e-4f ? 0 : 1;
}
