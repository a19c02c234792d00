/*
 * midi_front.hpp — host-side MIDI front of an instance: MidiParser -> VoiceAllocator<N> ->
 * [MidiVoiceHandler; N], the integer/event logic that precedes the voices in the reference's
 * polyphonic graphs (examples/fm-synth/src/lib.rs:72-90, examples/electric-piano/src/main.rs:55-75).
 * It stays on the CPU (it only runs on events) and emits, per voice, frame-accurate
 * `frequency` value events and `gate` scalar events for the device kernels.
 *
 * Reference semantics kept: handler order "all note_on events, then all note_off events" per
 * node (derive(Node) field order, oscen-macros/src/lib.rs:266-295); 32-slot event queues whose
 * overflow is silently dropped (graph/types.rs:18, `let _ = try_push`); LRU allocation with
 * "steal a released voice first" (voice_allocator.rs:57-91).
 */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace oscen {

constexpr int kMaxPolyVoices = 24;      /* MAX_VOICES, voice_allocator.rs:5 */
constexpr int kQueueCap = 32;           /* MAX_STATIC_EVENTS_PER_ENDPOINT */

/* MidiVoiceHandler::midi_note_to_freq, midi.rs:69-72 */
inline float midi_note_to_freq(uint8_t note) {
    const float semitone_offset = (float)note - 69.0f;
    return 440.0f * powf(2.0f, semitone_offset / 12.0f);
}

/* MidiParser::parse_bytes, midi.rs:147-171: 0 none, 1 note-on, 2 note-off */
inline int midi_parse(const uint8_t *data, uint32_t len, uint8_t &note, float &velocity) {
    if (len < 3) return 0;
    const uint8_t status = data[0] & 0xF0;
    note = data[1];
    const uint8_t vel = data[2];
    if (status == 0x80) return 2;
    if (status == 0x90) {
        if (vel == 0) return 2;
        float f = (float)vel / 127.0f;
        if (f < 0.0f) f = 0.0f;
        if (f > 1.0f) f = 1.0f;
        velocity = f;
        return 1;
    }
    return 0;
}

struct VoiceAllocator {                 /* voice_allocator.rs:29-110 */
    int num_voices = 0;
    bool active[kMaxPolyVoices] = {}, released[kMaxPolyVoices] = {};
    int note[kMaxPolyVoices];
    uint32_t age[kMaxPolyVoices] = {};
    uint32_t current_age = 0;
    explicit VoiceAllocator(int n = 0) : num_voices(n) { for (int &x : note) x = -1; }
    int allocate(uint8_t nt) {
        for (int i = 0; i < num_voices; ++i)
            if (!active[i]) { take(i, nt); return i; }
        int best = 0; bool have = false; uint32_t bp = 0, ba = 0;
        for (int i = 0; i < num_voices; ++i) {       /* min_by_key((released?0:1, age)), first minimum */
            const uint32_t p = released[i] ? 0u : 1u;
            if (!have || p < bp || (p == bp && age[i] < ba)) { have = true; best = i; bp = p; ba = age[i]; }
        }
        take(best, nt);
        return best;
    }
    int find(uint8_t nt) const {
        for (int i = 0; i < num_voices; ++i)
            if (active[i] && !released[i] && note[i] == (int)nt) return i;
        return -1;
    }
    void release(int i) { if (i >= 0 && i < num_voices) { released[i] = true; note[i] = -1; } }
private:
    void take(int i, uint8_t nt) { active[i] = true; released[i] = false; note[i] = nt; age[i] = current_age++; }
};

struct MidiMessage { uint32_t frame; uint8_t bytes[3]; uint8_t len; uint32_t seq; };
struct VoiceEvent { uint32_t voice_in_instance, frame; bool is_gate; float value; };

struct MidiFront {
    VoiceAllocator alloc;
    int cur_note[kMaxPolyVoices];
    float cur_freq[kMaxPolyVoices];
    std::vector<MidiMessage> pending;
    explicit MidiFront(int n = 0) : alloc(n) {
        for (int i = 0; i < kMaxPolyVoices; ++i) { cur_note[i] = -1; cur_freq[i] = 440.0f; }
    }
    /* midi_in.try_push(): at most 32 pending messages per block */
    bool push(uint32_t frame, const uint8_t *bytes, uint32_t len, uint32_t seq) {
        if ((int)pending.size() >= kQueueCap) return false;
        MidiMessage m{};
        m.frame = frame; m.len = (uint8_t)(len > 3 ? 3 : len); m.seq = seq;
        for (uint32_t i = 0; i < m.len; ++i) m.bytes[i] = bytes[i];
        pending.push_back(m);
        return true;
    }
    /* run the front for one block; appends per-voice events (frame order); drops the rest */
    void run(uint32_t frames, std::vector<VoiceEvent> &out) {
        if (pending.empty()) return;
        std::stable_sort(pending.begin(), pending.end(),
                         [](const MidiMessage &a, const MidiMessage &b) { return a.frame < b.frame; });
        size_t i = 0;
        while (i < pending.size() && pending[i].frame < frames) {
            const uint32_t f = pending[i].frame;
            uint8_t on_note[kQueueCap]; float on_vel[kQueueCap]; int n_on = 0;
            uint8_t off_note[kQueueCap]; int n_off = 0;
            for (; i < pending.size() && pending[i].frame == f; ++i) {          /* MidiParser */
                uint8_t nt = 0; float vel = 0.0f;
                const int r = midi_parse(pending[i].bytes, pending[i].len, nt, vel);
                if (r == 1 && n_on < kQueueCap) { on_note[n_on] = nt; on_vel[n_on] = vel; ++n_on; }
                else if (r == 2 && n_off < kQueueCap) off_note[n_off++] = nt;
            }
            struct Q { uint8_t kind, note; float vel; };
            Q vq[kMaxPolyVoices][kQueueCap]; int nvq[kMaxPolyVoices] = {};
            for (int k = 0; k < n_on; ++k) {                                    /* VoiceAllocator::on_note_on */
                const int vi = alloc.allocate(on_note[k]);
                if (nvq[vi] < kQueueCap) vq[vi][nvq[vi]++] = Q{1, on_note[k], on_vel[k]};
            }
            for (int k = 0; k < n_off; ++k) {                                   /* VoiceAllocator::on_note_off */
                const int vi = alloc.find(off_note[k]);
                if (vi >= 0) {
                    if (nvq[vi] < kQueueCap) vq[vi][nvq[vi]++] = Q{2, off_note[k], 0.0f};
                    alloc.release(vi);
                }
            }
            for (int v = 0; v < alloc.num_voices; ++v) {                        /* MidiVoiceHandler */
                int n_gate = 0;
                bool touched = false;
                for (int k = 0; k < nvq[v]; ++k) if (vq[v][k].kind == 1) {
                    cur_note[v] = vq[v][k].note;
                    cur_freq[v] = midi_note_to_freq(vq[v][k].note);
                    touched = true;
                    if (n_gate < kQueueCap) { out.push_back(VoiceEvent{(uint32_t)v, f, true, vq[v][k].vel}); ++n_gate; }
                }
                for (int k = 0; k < nvq[v]; ++k) if (vq[v][k].kind == 2) {
                    if (cur_note[v] == (int)vq[v][k].note) {
                        if (n_gate < kQueueCap) { out.push_back(VoiceEvent{(uint32_t)v, f, true, 0.0f}); ++n_gate; }
                        cur_note[v] = -1;
                    }
                }
                if (touched) out.push_back(VoiceEvent{(uint32_t)v, f, false, cur_freq[v]});   /* frequency value edge */
            }
        }
        pending.clear();
    }
};

}  // namespace oscen
