"""Committed golden fixtures (tests/golden/): the reference's own known-answer vectors (reference_kats.json) and the
oracle's waveforms for the BASELINE config graphs, frozen by tests/golden/make_golden.py.

CPU: the oracle must reproduce every fixture bit for bit (it has not drifted: same libm, same flags) and honour the
reference's vectors.  GPU: the CUDA path (fused kernels and node plans, through the C ABI) against the committed bytes,
tolerance |gpu - golden| <= 1e-6 + 1e-5 |golden| as BASELINE.json's north_star states (in practice bit-identical).
(The file name sorts last on purpose: these tests were added after the round's GPU budget was spent.)"""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

import oracle_lib as O
from parity_util import assert_close

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLDEN)
import make_golden as MG  # noqa: E402


def fixture(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return z["output"], int(z["kind"]), int(z["voices"]), int(z["voices_per_instance"]), float(z["sample_rate"])


KATS = json.load(open(os.path.join(GOLDEN, "reference_kats.json")))


# ------------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("name", sorted(MG.CASES))
def test_oracle_reproduces_the_committed_fixture_bit_for_bit(name):
    want, kind, n, vpi, sr = fixture(name)
    assert (kind, n, vpi) == MG.CASES[name]
    got = MG.drive(O.OracleBank(kind, n, vpi, sr), name)
    assert got.shape == want.shape and np.array_equal(got, want), name


def test_oracle_honours_the_reference_vectors_in_the_fixture_file():
    L = O.lib()
    k = KATS["tpt_impulse_response"]
    f = O.Tpt()
    L.orc_tpt_new(C.byref(f), k["cutoff"], k["q"]); L.orc_tpt_set_sample_rate(C.byref(f), k["sample_rate"]); L.orc_tpt_prepare(C.byref(f))
    f.cutoff, f.q, f.f_mod = k["cutoff"], k["q"], 0.0
    for i, want in enumerate(k["output"]):
        f.input = 1.0 if i == 0 else 0.0
        L.orc_tpt_process(C.byref(f))
        assert abs(f.output - want) <= k["epsilon"], (i, f.output, want)
    m = KATS["midi_note_to_freq"]
    for note, hz in m["exact"].items():
        assert L.orc_midi_note_to_freq(int(note)) == hz
    for note, hz in m["approx_0p01"].items():
        assert abs(L.orc_midi_note_to_freq(int(note)) - hz) < 0.01
    assert O.OracleBank(O.K_SAT4, 1, 1).L.orc_bank_latency(O.OracleBank(O.K_SAT4, 1, 1).h) == KATS["latency_samples"]["sat_graph_4x"]


def test_fixtures_are_what_make_golden_generates():
    assert sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz")) == sorted(MG.CASES)


# ------------------------------------------------------------------------------------------------ GPU
NODE_PLANS = {"fm_bank_64", "twin_peaks_16", "sat_1x_16", "synth_adsr_16"}      # graphs that also have a node plan


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MG.CASES))
def test_cuda_path_matches_the_committed_fixture(name):
    from oscen_b200 import abi
    want, kind, n, vpi, sr = fixture(name)
    plans = [abi.PLAN_FUSED] + ([abi.PLAN_NODES] if name in NODE_PLANS else [])
    for plan in plans:
        got = MG.drive(abi.Engine(kind, n, vpi, sr, 0, plan), name)
        max_err, exact = assert_close(got, want, f"{name} plan={plan}")
        print(f"{name} plan={plan}: max_err={max_err:.3e} bit_exact={exact:.4f}")


@pytest.mark.gpu
def test_cuda_tpt_filter_matches_the_reference_impulse_response():
    """the reference's one golden waveform (filters/tpt/mod.rs:152-161,227-264) against the CUDA TptFilter body directly:
    a DSL graph `audio_in -> TptFilter -> out`, both back ends"""
    from oscen_b200 import abi
    k = KATS["tpt_impulse_response"]
    src = f"input audio_in: stream; output stream out; node f = TptFilter::new({k['cutoff']!r}, {k['q']!r}); connection audio_in -> f.input; connection f.output -> out;"
    x = np.zeros((8, 4), dtype=np.float32); x[0, :] = 1.0
    for plan in (abi.PLAN_NODES, abi.PLAN_FUSED):
        eng = abi.Engine.from_graph(src, 4, 1, k["sample_rate"], plan=plan)
        eng.stream_in(0, x); eng.process_block(8)
        y = eng.read_voice_out()
        for v in range(4):
            assert np.abs(y[:, v] - np.array(k["output"], dtype=np.float32)).max() <= k["epsilon"], (plan, y[:, v])
