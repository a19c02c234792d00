/* tests/csrc/host_cuda: the device MIDI front's sort / scan are restated by the host runner (std::stable_sort, a loop) */
