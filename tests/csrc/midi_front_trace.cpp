// Host-only trace of the product's MIDI front (oscen_b200/csrc/midi_front.hpp): reads
// "<frames>" then lines "<frame> <b0> <b1> <b2>" (one instance, N voices) until EOF / "block",
// prints the per-voice events the engine would stage: "<voice> <frame> <G|F> <value bits>".
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../oscen_b200/csrc/midi_front.hpp"

int main(int argc, char **argv) {
    const int n_voices = argc > 1 ? atoi(argv[1]) : 8;
    oscen::MidiFront front(n_voices);
    char line[256];
    unsigned frames = 0, seq = 0;
    while (fgets(line, sizeof line, stdin)) {
        if (strncmp(line, "block", 5) == 0) {
            sscanf(line + 5, "%u", &frames);
            std::vector<oscen::VoiceEvent> out;
            front.run(frames, out);
            for (const auto &e : out) {
                unsigned bits; memcpy(&bits, &e.value, 4);
                printf("%u %u %c %08x\n", e.voice_in_instance, e.frame, e.is_gate ? 'G' : 'F', bits);
            }
            printf("end\n");
            continue;
        }
        unsigned f, b0, b1, b2;
        if (sscanf(line, "%u %u %u %u", &f, &b0, &b1, &b2) == 4) {
            const uint8_t bytes[3] = {(uint8_t)b0, (uint8_t)b1, (uint8_t)b2};
            front.push(f, bytes, 3, seq++);
        }
    }
    return 0;
}
