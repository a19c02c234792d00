/*
 * graph.hpp — C++ host-side mirror of the API surface Oscen's `graph!` macro generates
 * (oscen-graph-compiler/src/codegen/mod.rs:1292-1392), for the graphs on the hot path, on top of
 * the C ABI in include/oscen_b200.h.  It exists because the reference's host language (Rust,
 * nightly) is not available in this build environment; INTEGRATION.md shows the equivalent Rust
 * shim.  Names follow the generated struct: new()/init() -> constructor, `set_<input>`,
 * `set_<input>_with_ramp`, `<event>.try_push`, `process_block(frames)`, `<out>_block`.
 *
 * A `Bank<...>` owns N independent instances of one generated graph (what N separate
 * `FMGraph::new()` values would be), rendered together on one B200.
 * Header-only; link with oscen_b200/lib/liboscen_b200.so.  Throws std::runtime_error on a
 * non-zero ABI return (the reference's process* are infallible; an error here is a CUDA error
 * or a misuse).
 */
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/oscen_b200.h"

namespace oscen_b200 {

inline void check(int rc, oscen_b200_engine *e, const char *what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + oscen_b200_last_error(e));
}

/* EventInstance { frame_offset, payload: Scalar(v) } for one voice's `gate` input */
struct GateEvent { uint32_t voice, frame_offset; float velocity; };

class Bank {
protected:
    /* for GraphBank: the engine is created from a compiled DSL graph */
    Bank(uint32_t instances, uint32_t voices_per_instance) : instances_(instances), vpi_(voices_per_instance) {}
    void adopt(oscen_b200_engine *e) { e_ = e; channels_ = oscen_b200_channels(e_); }
public:
    Bank(uint32_t graph_kind, uint32_t instances, uint32_t voices_per_instance, float sample_rate,
         int device = 0, uint32_t plan = OSCEN_B200_PLAN_FUSED,
         uint32_t keep = OSCEN_B200_KEEP_INSTANCE_OUT)
        : instances_(instances), vpi_(voices_per_instance) {
        oscen_b200_config cfg{};
        cfg.graph_kind = graph_kind; cfg.n_voices = instances * voices_per_instance;
        cfg.voices_per_instance = voices_per_instance; cfg.sample_rate = sample_rate;   /* G::init(sr) */
        cfg.device = device; cfg.plan = plan; cfg.keep = keep;
        int rc = oscen_b200_create(&cfg, &e_);
        if (rc != 0) throw std::runtime_error(std::string("oscen_b200_create: ") + oscen_b200_last_error(nullptr));
        channels_ = oscen_b200_channels(e_);
    }
    ~Bank() { oscen_b200_destroy(e_); }
    Bank(const Bank &) = delete;
    Bank &operator=(const Bank &) = delete;

    /* `g.<value> = v` on every instance */
    void set_value(uint32_t input, float v) { check(oscen_b200_set_value(e_, input, 0, instances_ * vpi_, &v, 1), e_, "set_value"); }
    /* per-voice values (e.g. MidiVoiceHandler::frequency) */
    void set_values(uint32_t input, const std::vector<float> &v, uint32_t first_voice = 0) {
        check(oscen_b200_set_value(e_, input, first_voice, (uint32_t)v.size(), v.data(), 0), e_, "set_value");
    }
    /* set_<v>_with_ramp(value, frames) */
    void set_value_with_ramp(uint32_t input, float v, uint32_t frames) { check(oscen_b200_set_value_ramp(e_, input, v, frames), e_, "set_value_ramp"); }
    /* gate.try_push(EventInstance{frame_offset, Scalar(velocity)}) */
    void push_gates(const std::vector<GateEvent> &g) {
        std::vector<oscen_b200_event> ev(g.size());
        for (size_t i = 0; i < g.size(); ++i) ev[i] = {g[i].voice, g[i].frame_offset, OSCEN_B200_EVENT_GATE, g[i].velocity};
        check(oscen_b200_push_events(e_, ev.data(), ev.size()), e_, "push_events");
    }
    /* midi_in.try_push(raw_midi_event(bytes)) on one instance */
    void push_midi(uint32_t instance, uint32_t frame_offset, const uint8_t bytes[3]) {
        check(oscen_b200_push_midi(e_, instance, frame_offset, bytes, 3), e_, "push_midi");
    }
    /* writing `audio_in_block` of every instance: [frames][voices] */
    void stream_in(const float *host, uint32_t frames) { check(oscen_b200_stream_in(e_, 0, host, frames, 0), e_, "stream_in"); }

    void process_block(uint32_t frames) { check(oscen_b200_process_block(e_, frames), e_, "process_block"); frames_ = frames; }

    /* `<out>_block` of instances [first, first+count): [frames][count][channels] */
    std::vector<float> out_block(uint32_t first = 0, uint32_t count = 0) {
        if (count == 0) count = instances_ - first;
        std::vector<float> out((size_t)frames_ * count * channels_);
        check(oscen_b200_read_instance_out(e_, out.data(), first, count), e_, "read_instance_out");
        return out;
    }
    /* the bank's mix bus: [frames][channels] */
    std::vector<float> mix() {
        std::vector<float> out((size_t)frames_ * channels_);
        check(oscen_b200_read_mix(e_, out.data()), e_, "read_mix");
        return out;
    }
    /* BlockRender::render (oscen-lib/src/graph/offline.rs:46-90): `frames` frames in chunks of at most MAX_BLOCK_SIZE;
       returns the instances' `<out>_block`s back to back: [frames][instances][channels] */
    std::vector<float> render(uint32_t frames) {
        std::vector<float> out;
        out.reserve((size_t)frames * instances_ * channels_);
        for (uint32_t pos = 0; pos < frames;) {
            const uint32_t chunk = frames - pos < MAX_BLOCK_SIZE ? frames - pos : MAX_BLOCK_SIZE;
            process_block(chunk);
            const std::vector<float> b = out_block();
            out.insert(out.end(), b.begin(), b.end());
            pos += chunk;
        }
        return out;
    }
    uint32_t latency_samples() const { return oscen_b200_latency_samples(e_); }
    uint32_t channels() const { return channels_; }
    oscen_b200_engine *raw() { return e_; }

    static constexpr uint32_t MAX_BLOCK_SIZE = OSCEN_B200_MAX_BLOCK;   /* Self::MAX_BLOCK_SIZE */

private:
    oscen_b200_engine *e_ = nullptr;
    uint32_t instances_, vpi_, channels_ = 1, frames_ = 0;
};

/* FMGraph x N: examples/fm-synth/src/lib.rs:22-131 (8 voices per instance, ramped level /
   feedback / route / cutoff / resonance / env_amount inputs with 2205-frame default ramps) */
class FMGraphBank : public Bank {
public:
    FMGraphBank(uint32_t instances, float sample_rate, int device = 0)
        : Bank(OSCEN_B200_GRAPH_FM_VOICE, instances, 8, sample_rate, device) {}
    static constexpr uint32_t DEFAULT_RAMP = 2205;
    /* set_<v>(value): ramp over the declared default duration for `[ramp: 2205]` inputs */
    void set_op3_level(float v) { set_value_with_ramp(OSCEN_B200_FM_OP3_LEVEL, v, DEFAULT_RAMP); }
    void set_op3_feedback(float v) { set_value_with_ramp(OSCEN_B200_FM_OP3_FEEDBACK, v, DEFAULT_RAMP); }
    void set_op2_level(float v) { set_value_with_ramp(OSCEN_B200_FM_OP2_LEVEL, v, DEFAULT_RAMP); }
    void set_op2_feedback(float v) { set_value_with_ramp(OSCEN_B200_FM_OP2_FEEDBACK, v, DEFAULT_RAMP); }
    void set_route(float v) { set_value_with_ramp(OSCEN_B200_FM_ROUTE, v, DEFAULT_RAMP); }
    void set_filter_cutoff(float v) { set_value_with_ramp(OSCEN_B200_FM_FILTER_CUTOFF, v, DEFAULT_RAMP); }
    void set_filter_resonance(float v) { set_value_with_ramp(OSCEN_B200_FM_FILTER_RESONANCE, v, DEFAULT_RAMP); }
    void set_filter_env_amount(float v) { set_value_with_ramp(OSCEN_B200_FM_FILTER_ENV_AMOUNT, v, DEFAULT_RAMP); }
    /* plain value inputs: direct field assignment in the reference (main.rs:168-194) */
    void set_op3_ratio(float v) { set_value(OSCEN_B200_FM_OP3_RATIO, v); }
    void set_op2_ratio(float v) { set_value(OSCEN_B200_FM_OP2_RATIO, v); }
    void set_op3_attack(float v) { set_value(OSCEN_B200_FM_OP3_ATTACK, v); }
    void set_op3_release(float v) { set_value(OSCEN_B200_FM_OP3_RELEASE, v); }
};

/* TwinPeaksGraph x N: examples/nih-twin-peaks/src/lib.rs:15-48 */
class TwinPeaksBank : public Bank {
public:
    TwinPeaksBank(uint32_t instances, float sample_rate, int device = 0)
        : Bank(OSCEN_B200_GRAPH_TWIN_PEAKS, instances, 1, sample_rate, device) {}
    void set_cutoff_a(float v) { set_value(OSCEN_B200_TWIN_CUTOFF_A, v); }
    void set_cutoff_b(float v) { set_value(OSCEN_B200_TWIN_CUTOFF_B, v); }
    void set_resonance(float v) { set_value(OSCEN_B200_TWIN_RESONANCE, v); }
};

/* Any graph, from its `graph!` text: what `graph! { ... }` expands to at Rust compile time is
   compiled here at run time (oscen_b200_graph_compile: parse -> lower -> plan; PLAN_FUSED additionally
   generates one CUDA kernel for the graph and compiles it with NVRTC).  Inputs are addressed by the
   names the DSL declares; `set(name, v)` ramps over the declared `ramp: N` frames like the generated
   `set_<name>(v)` (codegen/mod.rs:917-976), or assigns when the input has no ramp. */
class GraphBank : public Bank {
public:
    GraphBank(const std::string &dsl_text, uint32_t instances, uint32_t voices_per_instance, float sample_rate, int device = 0,
              uint32_t plan = OSCEN_B200_PLAN_FUSED, uint32_t keep = OSCEN_B200_KEEP_INSTANCE_OUT)
        : Bank(instances, voices_per_instance) {
        char err[1024];
        if (oscen_b200_graph_compile(dsl_text.c_str(), &g_, err, sizeof err) != 0) throw std::runtime_error(std::string("graph!: ") + err);
        oscen_b200_config cfg{};
        cfg.graph_kind = OSCEN_B200_GRAPH_CUSTOM; cfg.n_voices = instances * voices_per_instance;
        cfg.voices_per_instance = voices_per_instance; cfg.sample_rate = sample_rate;
        cfg.device = device; cfg.plan = plan; cfg.keep = keep;
        oscen_b200_engine *e = nullptr;
        const int rc = oscen_b200_create_from_graph(g_, &cfg, &e);
        if (rc != 0) { oscen_b200_graph_free(g_); g_ = nullptr; throw std::runtime_error(std::string("oscen_b200_create_from_graph: ") + oscen_b200_last_error(nullptr)); }
        adopt(e);
    }
    ~GraphBank() { oscen_b200_graph_free(g_); }
    uint32_t input(const std::string &name) const {
        const int id = oscen_b200_graph_input_id(g_, name.c_str());
        if (id < 0) throw std::runtime_error("graph has no value input `" + name + "`");
        return (uint32_t)id;
    }
    /* set_<name>(v) */
    void set(const std::string &name, float v) {
        const uint32_t id = input(name);
        oscen_b200_input_info info{};
        oscen_b200_graph_input_info(g_, id, &info);
        if (info.ramp_frames) set_value_with_ramp(id, v, info.ramp_frames); else set_value(id, v);
    }
    /* set_<name>_immediate(v) / set_<name>_with_ramp(v, frames) */
    void set_immediate(const std::string &name, float v) { set_value(input(name), v); }
    void set_with_ramp(const std::string &name, float v, uint32_t frames) { set_value_with_ramp(input(name), v, frames); }
    /* graph.<slot>.load(..) of an `external <slot>: AudioAsset;` (asset/mod.rs): frames x channels f32 at the bank's sample
       rate; the SamplePlayers bound to the slot restart at the next block */
    void load_asset(const std::string &slot, const std::vector<float> &interleaved, uint32_t channels = 1) {
        const int id = oscen_b200_graph_asset_id(g_, slot.c_str());
        if (id < 0) throw std::runtime_error("graph has no external slot `" + slot + "`");
        const uint32_t ch = channels ? channels : 1u;
        check(oscen_b200_load_asset(raw(), (uint32_t)id, interleaved.data(), (uint32_t)(interleaved.size() / ch), ch), raw(), "load_asset");
    }
    /* (no reference counterpart: one reference instance has one f32 per value input) compile the kernel variant that reads the
       value inputs which hold one value for the whole bank from the constant bank; returns the mask of the ids taken.  Runs
       NVRTC: not for the audio thread */
    uint32_t specialize() {
        uint32_t mask = 0;
        check(oscen_b200_specialize(raw(), &mask), raw(), "specialize");
        return mask;
    }
    std::string describe() const {
        std::string s(1 << 16, '\0');
        oscen_b200_graph_describe(g_, &s[0], s.size());
        s.resize(std::char_traits<char>::length(s.c_str()));
        return s;
    }
private:
    oscen_b200_graph *g_ = nullptr;
};

}  // namespace oscen_b200
