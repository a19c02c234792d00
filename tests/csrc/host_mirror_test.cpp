// C++ host driving the C ABI through the generated-graph mirror (oscen_b200/host/graph.hpp),
// the way examples/pivot/src/main.rs:599-725 smoke-tests a graph: note on -> audible output.
// Exit code 0 = ok, 77 = no GPU (compile/link/ABI check only).
#include <cmath>
#include <cstdio>
#include "../../oscen_b200/host/graph.hpp"

static const char *kSynthDsl = R"(graph! {
    name: Synth;
    input cutoff: value = 1200.0 [20.0..20000.0, log, ramp: 480];
    input gate: event;
    output audio_out: stream;
    nodes {
        osc = PolyBlepOscillator::saw(220.0, 0.5) * 2;
        env = AdsrEnvelope::new(0.005, 0.05, 0.6, 0.1);
        filter = TptFilter::new(1200.0, 0.9);
        vca = Gain::new(1.0);
    }
    connections {
        gate -> env.gate;
        [sinc_iir] osc.output -> filter.input;
        cutoff -> filter.cutoff;
        filter.output -> vca.input;
        env.output -> vca.gain;
        vca.output -> audio_out;
    }
})";

int main() {
    {   /* the DSL front end needs no GPU: a graph compiles (and a broken one reports line:col) everywhere */
        oscen_b200_graph *g = nullptr; char err[256];
        if (oscen_b200_graph_compile(kSynthDsl, &g, err, sizeof err) != 0) { std::printf("compile: %s\n", err); return 1; }
        if (oscen_b200_graph_input_id(g, "cutoff") != 0 || oscen_b200_graph_n_inputs(g) != 1) return 1;
        oscen_b200_graph_free(g);
        if (oscen_b200_graph_compile("output stream out; node o = Nope::new(); connection o.output -> out;", &g, err, sizeof err) == 0) return 1;
        std::printf("front end: ok (a bad graph reports `%s`)\n", err);
    }
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
    if (!(peak > 1e-4f)) return 1;
    /* the same smoke test on a graph given as `graph!` text: multirate oscillator, ramped cutoff, generated kernel */
    oscen_b200::GraphBank dsl(kSynthDsl, 16, 1, 48000.0f);
    std::vector<oscen_b200::GateEvent> on;
    for (uint32_t v = 0; v < 16; ++v) on.push_back({v, v, 0.9f});
    dsl.push_gates(on);
    dsl.set("cutoff", 4000.0f);                  /* ramps over the declared 480 frames */
    float peak2 = 0.0f;
    for (int b = 0; b < 8; ++b) {
        dsl.process_block(256);
        for (float x : dsl.out_block()) peak2 = std::fmax(peak2, std::fabs(x));
    }
    std::printf("dsl graph peak = %g, latency = %u\n", peak2, dsl.latency_samples());
    return peak2 > 1e-4f ? 0 : 1;
}
