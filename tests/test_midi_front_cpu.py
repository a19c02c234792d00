"""Host-side MIDI front of the product (MidiParser -> VoiceAllocator<N> -> MidiVoiceHandler,
oscen_b200/csrc/midi_front.hpp) against the oracle's restatement of the same reference chain
(oscen-lib/src/midi.rs, voice_allocator.rs), on random MIDI streams with voice stealing."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = O.lib()


@pytest.fixture(scope="module")
def tracer(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("midi") / "midi_front_trace")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-o", exe,
                           os.path.join(ROOT, "tests", "csrc", "midi_front_trace.cpp"), "-lm"])
    return exe


def reference_chain(n_voices, blocks):
    """the per-frame chain of the generated graph, built from the oracle's
    primitives: returns [(voice, frame, kind, f32 bits)] per block"""
    alloc = O.VoiceAlloc()
    L.orc_voice_alloc_new(C.byref(alloc), n_voices)
    cur_note = [-1] * n_voices
    cur_freq = [440.0] * n_voices
    out_blocks = []
    for frames, msgs in blocks:
        out = []
        msgs = sorted(enumerate(msgs[:32]), key=lambda t: (t[1][0], t[0]))      # 32-slot midi_in queue, then by frame
        i = 0
        while i < len(msgs) and msgs[i][1][0] < frames:
            f = msgs[i][1][0]
            ons, offs = [], []
            while i < len(msgs) and msgs[i][1][0] == f:
                _, (_, b0, b1, b2) = msgs[i]
                data = (C.c_uint8 * 3)(b0, b1, b2)
                note, vel = C.c_uint8(0), C.c_float(0)
                r = L.orc_midi_parse(data, 3, C.byref(note), C.byref(vel))
                if r == 1 and len(ons) < 32:
                    ons.append((note.value, vel.value))
                elif r == 2 and len(offs) < 32:
                    offs.append(note.value)
                i += 1
            vq = [[] for _ in range(n_voices)]
            for note, vel in ons:
                vi = L.orc_voice_alloc_allocate(C.byref(alloc), note)
                if len(vq[vi]) < 32:
                    vq[vi].append((1, note, vel))
            for note in offs:
                vi = L.orc_voice_alloc_find(C.byref(alloc), note)
                if vi >= 0:
                    if len(vq[vi]) < 32:
                        vq[vi].append((2, note, 0.0))
                    L.orc_voice_alloc_release(C.byref(alloc), vi)
            for v in range(n_voices):
                touched, gates = False, 0
                for kind, note, vel in vq[v]:
                    if kind == 1:
                        cur_note[v] = note
                        cur_freq[v] = L.orc_midi_note_to_freq(note)
                        touched = True
                        if gates < 32:
                            out.append((v, f, "G", vel)); gates += 1
                for kind, note, vel in vq[v]:
                    if kind == 2 and cur_note[v] == note:
                        if gates < 32:
                            out.append((v, f, "G", 0.0)); gates += 1
                        cur_note[v] = -1
                if touched:
                    out.append((v, f, "F", cur_freq[v]))
        out_blocks.append([(v, f, k, struct.unpack("<I", struct.pack("<f", x))[0]) for v, f, k, x in out])
    return out_blocks


@pytest.mark.parametrize("n_voices,seed", [(8, 1), (16, 2), (4, 3), (8, 4)])
def test_midi_front_matches_reference_chain(tracer, n_voices, seed):
    rng = np.random.default_rng(seed)
    blocks = []
    for _ in range(12):
        frames = int(rng.choice([64, 256, 512]))
        msgs = []
        for _ in range(int(rng.integers(0, 40))):          # sometimes more than the 32-slot queue
            frame = int(rng.integers(0, frames + 40))        # some beyond the block: dropped
            note = int(rng.integers(48, 48 + 2 * n_voices))   # few notes: repeats, stealing, re-triggers
            kind = rng.random()
            if kind < 0.55:
                msgs.append((frame, 0x90, note, int(rng.integers(1, 128))))
            elif kind < 0.85:
                msgs.append((frame, 0x80, note, 0))
            elif kind < 0.95:
                msgs.append((frame, 0x90, note, 0))            # note-on velocity 0 = note-off
            else:
                msgs.append((frame, 0xB0, 7, 100))              # ignored controller message
        blocks.append((frames, msgs))
    text = ""
    for frames, msgs in blocks:
        for m in msgs:
            text += "%d %d %d %d\n" % m
        text += "block %d\n" % frames
    got_txt = subprocess.run([tracer, str(n_voices)], input=text, capture_output=True, text=True, check=True).stdout
    got_blocks, cur = [], []
    for line in got_txt.splitlines():
        if line == "end":
            got_blocks.append(cur); cur = []
        else:
            v, f, k, bits = line.split()
            cur.append((int(v), int(f), k, int(bits, 16)))
    want_blocks = reference_chain(n_voices, blocks)
    assert len(got_blocks) == len(want_blocks)
    n_events = 0
    for g, w in zip(got_blocks, want_blocks):
        assert g == w
        n_events += len(w)
    assert n_events > 20
