/*
 * midi_device.cuh — the MIDI front of every instance on the GPU: one thread per
 * instance runs MidiParser -> VoiceAllocator<N> -> [MidiVoiceHandler; N] for the block's messages
 * and writes the per-voice `gate` / `frequency` events straight into the device event list the
 * voice kernels read, so a bank of 10^5..10^6 instances takes note streams without a host loop
 * over instances.  Integer / event logic only; results are identical to the host front
 * (midi_front.hpp), which the CPU tests pin on the oracle.
 *
 *   MidiParser::on_midi_in / parse_bytes        oscen-lib/src/midi.rs:147-224
 *   VoiceAllocator::{allocate_voice, find_voice_for_note, release_voice, on_note_on, on_note_off}
 *                                               oscen-lib/src/voice_allocator.rs:57-136
 *   MidiVoiceHandler::{on_note_on, on_note_off, process}   oscen-lib/src/midi.rs:82-121
 *   handler order (all note_on, then all note_off)         oscen-macros/src/lib.rs:266-295
 *   32-slot queues, try_push drops                         oscen-lib/src/graph/types.rs:18,183-188
 *
 * Pipeline per block (all on the engine stream, no host synchronisation):
 *   1. midi_keys_kernel      key[i] = instance of message i, idx[i] = i
 *   2. cub::DeviceRadixSort  stable sort by instance: each instance's messages stay in PUSH order
 *   3. midi_front_kernel<false>  dry run on a private copy of the state: events per voice
 *   4. cub::DeviceScan       exclusive sum -> first event slot of every voice
 *   5. midi_front_kernel<true>   same walk, writes events (sorted by voice, frame, order) and
 *                                commits allocator / handler state; midi_finish_kernel writes
 *                                BlockCtl.n_events
 */
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "engine_types.h"

namespace oscen {

struct MidiMsg { uint32_t instance, frame; uint8_t bytes[3]; uint8_t len; };    /* == oscen_b200_midi_msg */

constexpr int kDevMaxPoly = 16;          /* voices_per_instance is a power of two; MAX_VOICES = 24 (voice_allocator.rs:5) */
constexpr uint32_t kMidiQueue = 32;      /* MAX_STATIC_EVENTS_PER_ENDPOINT */

struct MidiFrontArgs {
    const MidiMsg *msgs;                 /* push order */
    const uint32_t *keys, *idx;          /* sorted by instance (stable) */
    uint32_t n_msgs, n_inst, vpi, frames;
    /* allocator: word = (note + 1) [9 bits: note is any u8, midi.rs:8-35] | active << 9 | released << 10 per voice, age per voice, current_age per instance */
    uint32_t *alloc_word, *alloc_age, *alloc_cur_age;
    int *h_note; float *h_freq;          /* MidiVoiceHandler.current_note / current_frequency per voice */
    const float *note_freq;              /* [128] midi_note_to_freq, tabulated by the host libm (midi.rs:69-72) */
    uint32_t *ev_count;                  /* [n_voices] zeroed before the dry run */
    const uint32_t *ev_offset;           /* [n_voices] exclusive scan of ev_count */
    Event *events;
    uint32_t *dropped;                   /* messages that found the instance's midi_in queue full */
    uint32_t freq_input;                 /* value-input id MidiVoiceHandler.frequency is connected to (0 for the built-in graphs) */
};

__global__ void midi_keys_kernel(const MidiMsg *msgs, uint32_t n, uint32_t *keys, uint32_t *idx) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = msgs[i].instance; idx[i] = i; }
}

__global__ void midi_finish_kernel(const uint32_t *ev_count, const uint32_t *ev_offset, uint32_t n_voices, BlockCtl *ctl) {
    ctl->n_events = ev_offset[n_voices - 1] + ev_count[n_voices - 1];
}

template <bool EMIT>
__global__ void __launch_bounds__(128) midi_front_kernel(const MidiFrontArgs a) {
    const uint32_t inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= a.n_inst) return;
    uint32_t lo = 0, hi = a.n_msgs;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a.keys[mid] < inst) lo = mid + 1; else hi = mid; }
    if (lo >= a.n_msgs || a.keys[lo] != inst) return;

    /* midi_in.try_push: the first 32 messages pushed this block, the rest found the queue full */
    uint32_t fr[kMidiQueue], raw[kMidiQueue], cnt = 0, over = 0;
    for (uint32_t k = lo; k < a.n_msgs && a.keys[k] == inst; ++k) {
        if (cnt < kMidiQueue) {
            const MidiMsg m = a.msgs[a.idx[k]];
            fr[cnt] = m.frame;
            raw[cnt] = (uint32_t)m.bytes[0] | ((uint32_t)m.bytes[1] << 8) | ((uint32_t)m.bytes[2] << 16) | ((uint32_t)m.len << 24);
            ++cnt;
        } else ++over;
    }
    if (EMIT && over) atomicAdd(a.dropped, over);
    /* process_block stages the queue by frame_offset, stable (codegen/mod.rs:836-872) */
    for (uint32_t i = 1; i < cnt; ++i) {
        const uint32_t kf = fr[i], kr = raw[i];
        uint32_t j = i;
        while (j > 0 && fr[j - 1] > kf) { fr[j] = fr[j - 1]; raw[j] = raw[j - 1]; --j; }
        fr[j] = kf; raw[j] = kr;
    }

    const uint32_t vpi = a.vpi, v0 = inst * vpi;
    uint32_t word[kDevMaxPoly], age[kDevMaxPoly], out[kDevMaxPoly];
    int cur_note[kDevMaxPoly]; float cur_freq[kDevMaxPoly];
    for (uint32_t v = 0; v < vpi; ++v) {
        word[v] = a.alloc_word[v0 + v]; age[v] = a.alloc_age[v0 + v];
        cur_note[v] = a.h_note[v0 + v]; cur_freq[v] = a.h_freq[v0 + v];
        out[v] = EMIT ? a.ev_offset[v0 + v] : 0u;
    }
    uint32_t current_age = a.alloc_cur_age[inst];
    auto emit = [&](uint32_t v, uint32_t frame, uint32_t input_id, float scalar) {
        if (EMIT) a.events[out[v]] = Event{v0 + v, frame, input_id, scalar};
        out[v] += 1u;
    };

    uint32_t i = 0;
    while (i < cnt && fr[i] < a.frames) {            /* offsets >= frames are never reached: dropped with the queue */
        const uint32_t f = fr[i];
        /* per-voice output queues of the allocator for this frame, in push order: (voice, kind, note, vel) */
        uint8_t q_voice[kMidiQueue], q_kind[kMidiQueue], q_note[kMidiQueue]; float q_vel[kMidiQueue];
        uint32_t nq = 0;
        uint32_t j = i;
        /* VoiceAllocator::on_note_on for every note-on of this frame (voice_allocator.rs:114-124) */
        for (j = i; j < cnt && fr[j] == f; ++j) {
            const uint32_t r = raw[j];
            const uint32_t len = r >> 24, status = r & 0xF0u, note = (r >> 8) & 0xFFu, vel = (r >> 16) & 0xFFu;
            if (len < 3u || status != 0x90u || vel == 0u) continue;                 /* midi.rs:147-171 */
            float fv = (float)vel / 127.0f;
            fv = fv < 0.0f ? 0.0f : (fv > 1.0f ? 1.0f : fv);
            uint32_t pick = vpi;
            for (uint32_t v = 0; v < vpi; ++v) if (!((word[v] >> 9) & 1u)) { pick = v; break; }     /* first inactive */
            if (pick == vpi) {                       /* min_by_key((released ? 0 : 1, age)), first minimum */
                uint32_t bp = 2u, ba = 0u;
                for (uint32_t v = 0; v < vpi; ++v) {
                    const uint32_t p = ((word[v] >> 10) & 1u) ? 0u : 1u;
                    if (p < bp || (p == bp && age[v] < ba)) { pick = v; bp = p; ba = age[v]; }
                }
            }
            word[pick] = (note + 1u) | (1u << 9); age[pick] = current_age++;
            q_voice[nq] = (uint8_t)pick; q_kind[nq] = 1; q_note[nq] = (uint8_t)note; q_vel[nq] = fv; ++nq;
        }
        /* then VoiceAllocator::on_note_off (voice_allocator.rs:126-136) */
        for (j = i; j < cnt && fr[j] == f; ++j) {
            const uint32_t r = raw[j];
            const uint32_t len = r >> 24, status = r & 0xF0u, note = (r >> 8) & 0xFFu, vel = (r >> 16) & 0xFFu;
            if (len < 3u || !(status == 0x80u || (status == 0x90u && vel == 0u))) continue;
            for (uint32_t v = 0; v < vpi; ++v) {
                if (((word[v] >> 9) & 3u) == 1u && (word[v] & 0x1FFu) == note + 1u) {   /* active, not released, this note */
                    q_voice[nq] = (uint8_t)v; q_kind[nq] = 2; q_note[nq] = (uint8_t)note; q_vel[nq] = 0.0f; ++nq;
                    word[v] = (1u << 9) | (1u << 10);                                 /* released, note = None */
                    break;
                }
            }
        }
        i = j;
        /* MidiVoiceHandler per voice: all note_on events, then all note_off events (midi.rs:92-121) */
        for (uint32_t v = 0; v < vpi; ++v) {
            bool touched = false;
            for (uint32_t k = 0; k < nq; ++k) if (q_voice[k] == v && q_kind[k] == 1) {
                cur_note[v] = q_note[k]; cur_freq[v] = a.note_freq[q_note[k]]; touched = true;
                emit(v, f, kEventGate, q_vel[k]);
            }
            for (uint32_t k = 0; k < nq; ++k) if (q_voice[k] == v && q_kind[k] == 2) {
                if (cur_note[v] == (int)q_note[k]) { emit(v, f, kEventGate, 0.0f); cur_note[v] = -1; }
            }
            if (touched) emit(v, f, a.freq_input, cur_freq[v]);                                  /* frequency value edge */
        }
    }

    if (EMIT) {
        for (uint32_t v = 0; v < vpi; ++v) {
            a.alloc_word[v0 + v] = word[v]; a.alloc_age[v0 + v] = age[v];
            a.h_note[v0 + v] = cur_note[v]; a.h_freq[v0 + v] = cur_freq[v];
        }
        a.alloc_cur_age[inst] = current_age;
    } else {
        for (uint32_t v = 0; v < vpi; ++v) a.ev_count[v0 + v] = out[v];
    }
}

}  // namespace oscen
