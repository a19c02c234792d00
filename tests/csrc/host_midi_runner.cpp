/*
 * tests/csrc/host_midi_runner.cpp — TEST INFRASTRUCTURE.  Runs the DEVICE-side MIDI front (midi_device.cuh:
 * MidiParser -> VoiceAllocator<N> -> [MidiVoiceHandler; N], one thread per instance) on the host, with the pipeline of
 * engine.cu's midi_device_run() restated: keys, a stable sort by instance (cub::DeviceRadixSort there, std::stable_sort
 * here), the dry run, an exclusive scan, the emitting run.  The allocator / handler state lives in the handle across
 * blocks, initialised as midi_device_init() does.
 */
#include <algorithm>
#include <numeric>
#include <vector>

#include "midi_front.hpp"
#include "midi_device.cuh"

namespace {
using namespace oscen;
struct HostMidi {
    uint32_t n = 0, n_inst = 0, vpi = 0;
    std::vector<uint32_t> word, age, cur_age, ev_count, ev_offset;
    std::vector<int> note;
    std::vector<float> freq, tab;
};
}  // namespace

extern "C" void *oscen_host_midi_create(uint32_t n_voices, uint32_t vpi) {
    if (vpi == 0 || n_voices % vpi != 0 || vpi > (uint32_t)kDevMaxPoly) return nullptr;
    HostMidi *h = new HostMidi();
    h->n = n_voices; h->vpi = vpi; h->n_inst = n_voices / vpi;
    h->word.assign(n_voices, 0u); h->age.assign(n_voices, 0u); h->cur_age.assign(h->n_inst, 0u);
    h->note.assign(n_voices, -1); h->freq.assign(n_voices, 440.0f);
    h->ev_count.assign(n_voices, 0u); h->ev_offset.assign(n_voices, 0u);
    h->tab.resize(256);
    for (int i = 0; i < 256; ++i) h->tab[i] = midi_note_to_freq((uint8_t)i);
    return h;
}
extern "C" void oscen_host_midi_free(void *h) { delete (HostMidi *)h; }

/* returns the number of events written to `events` (sorted by voice, frame, order), or -1 when `cap` is too small */
extern "C" int oscen_host_midi_run(void *handle, const oscen::MidiMsg *msgs, uint32_t n_msgs, uint32_t frames, uint32_t freq_input,
                                   oscen::Event *events, uint32_t cap, uint32_t *dropped) {
    HostMidi *h = (HostMidi *)handle;
    std::vector<uint32_t> keys_in(n_msgs), idx_in(n_msgs), order(n_msgs), keys(n_msgs), idx(n_msgs);
    if (n_msgs) host_launch((n_msgs + 63u) / 64u, 1, 64, [&]() { midi_keys_kernel(msgs, n_msgs, keys_in.data(), idx_in.data()); });
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return keys_in[x] < keys_in[y]; });
    for (uint32_t k = 0; k < n_msgs; ++k) { keys[k] = keys_in[order[k]]; idx[k] = idx_in[order[k]]; }
    std::fill(h->ev_count.begin(), h->ev_count.end(), 0u);
    uint32_t n_dropped = 0;
    MidiFrontArgs a{};
    a.msgs = msgs; a.keys = keys.data(); a.idx = idx.data();
    a.n_msgs = n_msgs; a.n_inst = h->n_inst; a.vpi = h->vpi; a.frames = frames;
    a.alloc_word = h->word.data(); a.alloc_age = h->age.data(); a.alloc_cur_age = h->cur_age.data();
    a.h_note = h->note.data(); a.h_freq = h->freq.data(); a.note_freq = h->tab.data();
    a.ev_count = h->ev_count.data(); a.ev_offset = h->ev_offset.data(); a.events = events; a.dropped = &n_dropped;
    a.freq_input = freq_input;
    const uint32_t grid = (h->n_inst + 127u) / 128u;
    host_launch(grid, 1, 128, [&]() { midi_front_kernel<false>(a); });
    uint32_t total = 0;
    for (uint32_t v = 0; v < h->n; ++v) { h->ev_offset[v] = total; total += h->ev_count[v]; }
    if (total > cap) return -1;
    host_launch(grid, 1, 128, [&]() { midi_front_kernel<true>(a); });
    if (dropped) *dropped = n_dropped;
    return (int)total;
}
