"""oscen_b200_specialize: the generated kernel's variant that reads the bank-uniform value inputs from the constant bank.
The engine picks it per block (launch_fused, engine.cu) and must fall back to the general kernel in every block where one of
the specialised inputs ramps, receives a value event or was set per voice — with the oracle's bits in both cases.  The same
script runs on the host-executed kernels (CPU suite) and on the device (GPU suite)."""
import os

import numpy as np
import pytest

import oracle_lib as O

GRAPHS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "graphs")
GATE = O.EV_GATE
OP3_LEVEL, OP3_FB, OP2_LEVEL, CUTOFF, ENV_AMOUNT = 2, 3, 9, 21, 27       # ids of voice_fm3.graph = OSCEN_B200_FM_*
EXPECTED = ["U", "U", "G", "G", "U", "G", "G", "U", "G", "G", "U"]


def script(bank, n, which_kernel=None):
    """-> (voice_out of all blocks, instance_out of all blocks, the variant that rendered each block, the two masks)"""
    frames = 100
    vo, io, ran, masks = [], [], [], []
    bank.set_value(0, (110.0 * 2.0 ** ((np.arange(n) % 41) / 12.0)).astype(np.float32))      # per voice: never specialised
    bank.set_value(OP3_FB, 0.25)
    bank.set_value(ENV_AMOUNT, 900.0)                                                         # the cutoff follows its envelope
    if which_kernel:
        masks.append(bank.specialize())
    for blk in range(11):
        if blk == 0:
            bank.push_events([(v, (7 * v) % frames, GATE, 0.3 + 0.7 * ((v * 13) % 17) / 17.0) for v in range(n)])
        if blk == 2:
            bank.set_value_ramp(CUTOFF, 2600.0, 150)                  # runs through blocks 2 and 3
        if blk == 5:
            bank.push_events([(v, 10 + v % 50, OP3_LEVEL, 0.2 + 0.01 * v) for v in range(0, n, 3)])
        if blk == 7:
            bank.set_value(OP3_LEVEL, 0.4)                            # broadcast: uniform again
        if blk == 8:
            bank.set_value(OP2_LEVEL, (0.3 + 0.4 * (np.arange(n) % 5) / 5.0).astype(np.float32))
        if blk == 9:
            bank.push_events([(v, (3 * v) % frames, GATE, 0.0) for v in range(0, n, 2)])
        if blk == 10 and which_kernel:
            masks.append(bank.specialize())                           # again: op2_level is per voice now and leaves the mask
        bank.process_block(frames)
        vo.append(bank.read_voice_out()); io.append(bank.read_instance_out())
        if which_kernel:
            ran.append(which_kernel(bank))
    return np.concatenate(vo), np.concatenate(io), ran, masks


def check(got, ref, n_inputs=28):
    vo, io, ran, masks = got
    assert np.abs(ref[0]).max() > 0.01
    assert np.array_equal(vo, ref[0]), float(np.abs(vo - ref[0]).max())
    assert np.array_equal(io, ref[1])
    assert ran == EXPECTED, ran
    every = (1 << n_inputs) - 1
    assert masks[0] == every & ~1                                      # everything but the per-voice `frequency`
    assert masks[1] == every & ~1 & ~(1 << OP2_LEVEL)


def test_specialised_source_reads_the_constant_bank_and_keeps_no_register_for_those_inputs():
    from oscen_b200 import abi
    g = abi.Graph(open(os.path.join(GRAPHS, "voice_fm3.graph")).read())
    general, uni = g.kernel_source(48000.0, 0), g.kernel_source(48000.0, 0x0FFFFFFE)
    assert general == g.codegen()
    assert "a.u.p[" not in general and "a.u.p[21]" in uni and "a.u.p[0]" not in uni
    assert "P[21] = a.pv[" in general and "P[21] = a.pv[" not in uni and "P[0] = a.pv[" in uni
    assert "P[21] = row[21]" not in uni and "P[0] = row[0]" in uni     # ramps / value events on a specialised input never reach it


def test_host_executed_specialised_kernel_and_its_fallbacks_match_oracle():
    from host_kernel import HostEngine
    n, vpi = 72, 8
    text = open(os.path.join(GRAPHS, "voice_fm3.graph")).read()
    ref = script(O.OracleBank(O.K_FM, n, vpi), n)
    got = script(HostEngine(text, n, vpi), n, lambda b: "U" if b.last_kernel == "uniform" else "G")
    check(got, ref)


def test_specialize_is_refused_where_there_is_no_generated_kernel():
    """error behaviour without a GPU: a NULL engine is an argument error (the E_STATE cases need an engine: GPU suite)"""
    from oscen_b200 import abi
    assert abi.load().oscen_b200_specialize(None, None) == -1                   # OSCEN_B200_E_ARG


@pytest.mark.gpu
def test_specialised_kernel_and_its_fallbacks_match_oracle_on_device():
    from oscen_b200 import abi
    n, vpi = 200, 8
    text = open(os.path.join(GRAPHS, "voice_fm3.graph")).read()
    ref = script(O.OracleBank(O.K_FM, n, vpi), n)
    eng = abi.Engine.from_graph(abi.Graph(text), n, vpi, plan=abi.PLAN_FUSED)
    got = script(eng, n, lambda b: "U" if "<uniform=" in b.last_kernel_name() else "G")
    eng.close()
    check(got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("graph_file,kind,freq_id", [("synth_adsr.graph", O.K_SYNTH_ADSR, 2), ("twin_peaks.graph", O.K_TWIN, None), ("sat_4x.graph", O.K_SAT4, 0)])
def test_specialised_kernels_of_the_other_graphs_match_oracle_on_device(graph_file, kind, freq_id):
    from oscen_b200 import abi
    n = 130
    text = open(os.path.join(GRAPHS, graph_file)).read()
    outs = []
    for which in ("oracle", "gpu"):
        bank = O.OracleBank(kind, n, 1) if which == "oracle" else abi.Engine.from_graph(abi.Graph(text), n, 1, plan=abi.PLAN_FUSED)
        if freq_id is not None:
            bank.set_value(freq_id, (110.0 * 2.0 ** ((np.arange(n) % 31) / 12.0)).astype(np.float32))
        rng = np.random.default_rng(5)
        if which == "gpu":
            assert bank.specialize() != 0
        o = []
        for blk, frames in enumerate((200, 77, 256)):
            if kind == O.K_SYNTH_ADSR and blk in (0, 2):
                bank.push_events([(v, (5 * v) % frames, GATE, 0.8 if blk == 0 else 0.0) for v in range(n)])
            if kind == O.K_TWIN:
                bank.stream_in(0, rng.uniform(-1, 1, size=(frames, n)).astype(np.float32))
            bank.process_block(frames); o.append(bank.read_voice_out())
            if which == "gpu":
                assert "<uniform=" in bank.last_kernel_name()
        outs.append(np.concatenate(o))
    assert np.abs(outs[0]).max() > 0.01
    assert np.array_equal(outs[1], outs[0]), float(np.abs(outs[1] - outs[0]).max())


@pytest.mark.gpu
def test_specialize_is_refused_for_built_in_kinds_and_node_plans():
    from oscen_b200 import abi
    text = open(os.path.join(GRAPHS, "synth_adsr.graph")).read()
    for eng in (abi.Engine(abi.GRAPH_FM_VOICE, 64, 8), abi.Engine.from_graph(abi.Graph(text), 64, 1, plan=abi.PLAN_NODES)):
        with pytest.raises(abi.EngineError, match="specialize"):
            eng.specialize()
        eng.close()


# ---------------------------------------------------------------------------- ramp blocks of the generated kernels
def ramp_script(bank, n):
    """gates, a long ramp on the filter cutoff (nothing derived outside process() hangs on it: ramp SEGMENTS), value events in
    the middle of it, then a ramp on an envelope's attack time (hoisted step data: those blocks stay on the per-frame path)
    together with a second cutoff ramp, note-offs and re-triggers while it runs"""
    frames = 100
    vo, io = [], []
    bank.set_value(0, (110.0 * 2.0 ** ((np.arange(n) % 37) / 12.0)).astype(np.float32))
    bank.set_value(ENV_AMOUNT, 700.0)
    for blk in range(9):
        if blk == 0:
            bank.push_events([(v, (11 * v) % frames, GATE, 0.4 + 0.6 * ((v * 7) % 13) / 13.0) for v in range(n)])
        if blk == 1:
            bank.set_value_ramp(CUTOFF, 2600.0, 330)                  # blocks 1..4
        if blk == 2:
            bank.push_events([(v, 5 + (3 * v) % 90, OP2_LEVEL, 0.3 + 0.005 * v) for v in range(0, n, 4)])
        if blk == 5:
            bank.set_value_ramp(4, 0.03, 150)                         # op3_attack: blocks 5..6
            bank.set_value_ramp(CUTOFF, 1500.0, 250)                  # blocks 5..7
            bank.push_events([(v, (5 * v) % frames, GATE, 0.0) for v in range(0, n, 2)])
        if blk == 6:
            bank.push_events([(v, (7 * v) % frames, GATE, 0.9) for v in range(0, n, 4)])
        bank.process_block(frames)
        vo.append(bank.read_voice_out()); io.append(bank.read_instance_out())
    return np.concatenate(vo), np.concatenate(io)


@pytest.mark.parametrize("ramp_segments", ["1", "0"])
def test_host_executed_generated_kernel_ramp_blocks_with_and_without_ramp_segments(ramp_segments, monkeypatch):
    from host_kernel import HostEngine
    monkeypatch.setenv("OSCEN_B200_JIT_RAMP_SEGMENTS", ramp_segments)
    n, vpi = 72, 8
    eng = HostEngine(open(os.path.join(GRAPHS, "voice_fm3.graph")).read(), n, vpi)
    assert ("const bool ramp_seg" in eng.graph.codegen()) == (ramp_segments == "1")
    ref, got = ramp_script(O.OracleBank(O.K_FM, n, vpi), n), ramp_script(eng, n)
    assert np.abs(ref[0]).max() > 0.01
    assert np.array_equal(got[0], ref[0]), float(np.abs(got[0] - ref[0]).max())
    assert np.array_equal(got[1], ref[1])


@pytest.mark.gpu
@pytest.mark.parametrize("ramp_segments", ["1", "0"])
def test_generated_kernel_ramp_blocks_with_and_without_ramp_segments_on_device(ramp_segments):
    from oscen_b200 import abi
    n, vpi = 200, 8
    old = os.environ.get("OSCEN_B200_JIT_RAMP_SEGMENTS")
    try:
        os.environ["OSCEN_B200_JIT_RAMP_SEGMENTS"] = ramp_segments
        eng = abi.Engine.from_graph(abi.Graph(open(os.path.join(GRAPHS, "voice_fm3.graph")).read()), n, vpi, plan=abi.PLAN_FUSED)
    finally:
        if old is None:
            os.environ.pop("OSCEN_B200_JIT_RAMP_SEGMENTS", None)
        else:
            os.environ["OSCEN_B200_JIT_RAMP_SEGMENTS"] = old
    ref, got = ramp_script(O.OracleBank(O.K_FM, n, vpi), n), ramp_script(eng, n)
    eng.close()
    assert np.abs(ref[0]).max() > 0.01
    assert np.array_equal(got[0], ref[0]), float(np.abs(got[0] - ref[0]).max())
    assert np.array_equal(got[1], ref[1])
