"""GPU parity of the kernel VARIANTS one engine can pick between for the same block — every one must produce the oracle's
bits: node plan (generic kernels / plain register-prefetch kernels / TMA-tiled kernels, kernel_nodes.cuh) and the FMVoice
state handoff (coalesced loads / TMA bulk copies, kernel_fm.cuh).  The variants are selected the way a user would: by bank
shape (n_voices % 4), by what the block carries (events, ramps), and by the documented environment switches."""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

GATE = O.EV_GATE


def fm_script(bank, n, blocks):
    """gate-on burst in block 0, plain blocks, a ramp, plain again, gate-off, plain tail (block sizes not multiples of 8)"""
    out = []
    bank.set_value(0, (110.0 * 2.0 ** ((np.arange(n) % 37) / 12.0)).astype(np.float32))
    for blk, frames in enumerate(blocks):
        if blk == 0:
            bank.push_events([(v, (7 * v) % frames, GATE, 0.3 + 0.7 * ((v * 13) % 17) / 17.0) for v in range(n)])
        if blk == 3:
            bank.set_value_ramp(21, 2400.0, 150)               # filter cutoff ramps across this block and the next
        if blk == 5:
            bank.push_events([(v, (3 * v) % frames, GATE, 0.0) for v in range(0, n, 2)])
        bank.process_block(frames)
        out.append((bank.read_voice_out(), bank.read_instance_out()))
    return out


@pytest.mark.parametrize("n,env", [
    (256, {}),                                   # n % 4 == 0: plain blocks run the TMA-tiled FmOperator / TptFilter kernels
    (256, {"OSCEN_B200_NODE_TMA": "0"}),         # same bank on the register-prefetch plain kernels
    (256, {"OSCEN_B200_NODE_PLAIN": "0"}),       # and on the generic kernels only
    (136, {}),                                   # partly filled last CTA (136 = 2 * 64 + 8), TMA clips the box
    (130, {}),                                   # n % 4 != 0: no descriptors, plain kernels
])
def test_fm_node_plan_kernel_variants_match_oracle(n, env):
    from oscen_b200 import abi
    blocks = (200, 256, 100, 256, 77, 256, 33, 512)
    old = {k: os.environ.get(k) for k in ("OSCEN_B200_NODE_TMA", "OSCEN_B200_NODE_PLAIN")}
    try:
        for k in old:
            os.environ.pop(k, None)
        os.environ.update(env)
        vpi = 8 if n % 8 == 0 else 1
        ref = O.OracleBank(O.K_FM, n, vpi)
        gpu = abi.Engine(abi.GRAPH_FM_VOICE, n, vpi, 48000.0, 0, abi.PLAN_NODES, 3)
        ro, go = fm_script(ref, n, blocks), fm_script(gpu, n, blocks)
        gpu.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert max(float(np.abs(v).max()) for v, _ in ro) > 0.01
    for blk, ((vr, ir), (vg, ig)) in enumerate(zip(ro, go)):
        assert np.array_equal(vg, vr), (blk, float(np.abs(vg - vr).max()))
        assert np.array_equal(ig, ir), blk


def test_fm_node_plan_tma_kernels_are_the_ones_that_ran():
    """the claim above is only worth something if the plain blocks of the n % 4 == 0 bank really launched TMA kernels:
    the engine reports the launch sequence of its last captured plan"""
    from oscen_b200 import abi
    n = 256
    gpu = abi.Engine(abi.GRAPH_FM_VOICE, n, 8, 48000.0, 0, abi.PLAN_NODES, 3)
    gpu.set_value(0, 220.0)
    gpu.push_events([(v, v % 64, GATE, 0.8) for v in range(n)])
    gpu.process_block(128)
    first = gpu.last_kernel_name()
    gpu.process_block(128)
    plain = gpu.last_kernel_name()
    gpu.close()
    assert "fmop_node_kernel" in first and "tma" not in first
    assert "fmop_node_tma_kernel" in plain and "tpt_node_tma_kernel" in plain and "adsr_node_plain_kernel" in plain


@pytest.mark.parametrize("n", [64, 200, 1024])
def test_fm_general_kernel_with_tma_state_handoff_matches_oracle(n):
    """OSCEN_B200_TMA_STATE=1: fm_voice_kernel<true> hands its 29 state rows over with cp.async.bulk + mbarrier"""
    from oscen_b200 import abi
    old = os.environ.get("OSCEN_B200_TMA_STATE")
    try:
        os.environ["OSCEN_B200_TMA_STATE"] = "1"
        ref = O.OracleBank(O.K_FM, n, 8)
        gpu = abi.Engine(abi.GRAPH_FM_VOICE, n, 8, 48000.0, 0, abi.PLAN_FUSED, 3)
        outs = []
        for b in (ref, gpu):
            o = []
            b.set_value(0, (110.0 * 2.0 ** ((np.arange(n) % 29) / 12.0)).astype(np.float32))
            b.set_value(1, (1.0 + (np.arange(n) % 5)).astype(np.float32))       # op3_ratio per voice: the general kernel
            b.push_events([(v, v % 100, GATE, 0.9) for v in range(n)])
            for frames in (256, 100, 256):
                b.process_block(frames); o.append(b.read_voice_out())
            outs.append(np.concatenate(o))
        name = gpu.last_kernel_name()
        gpu.close()
    finally:
        if old is None:
            os.environ.pop("OSCEN_B200_TMA_STATE", None)
        else:
            os.environ["OSCEN_B200_TMA_STATE"] = old
    assert "fm_voice_kernel" in name, name
    assert np.abs(outs[0]).max() > 0.01
    assert np.array_equal(outs[1], outs[0]), float(np.abs(outs[1] - outs[0]).max())



@pytest.mark.parametrize("graph_file,kind,vpi", [("voice_fm3.graph", O.K_FM, 8), ("synth_adsr.graph", O.K_SYNTH_ADSR, 1), ("sat_4x.graph", O.K_SAT4, 1)])
def test_generated_kernel_with_and_without_segments_matches_oracle(graph_file, kind, vpi):
    """OSCEN_B200_JIT_SEGMENTS=0|1: the generator's per-frame loop and its segmented default, against the oracle bank of the
    same graph (gates where the graph has them, a value change between blocks, odd block sizes)"""
    from oscen_b200 import abi
    text = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "graphs", graph_file)).read()
    n = 136 if vpi == 8 else 130
    old = os.environ.get("OSCEN_B200_JIT_SEGMENTS")
    outs = []
    try:
        for seg in ("1", "0", None):
            if seg is None:
                bank = O.OracleBank(kind, n, vpi)
            else:
                os.environ["OSCEN_B200_JIT_SEGMENTS"] = seg
                bank = abi.Engine.from_graph(abi.Graph(text), n, vpi, plan=abi.PLAN_FUSED)
            freq_id = 2 if kind == O.K_SYNTH_ADSR else 0
            bank.set_value(freq_id, (110.0 * 2.0 ** ((np.arange(n) % 31) / 12.0)).astype(np.float32))
            o = []
            for blk, frames in enumerate((200, 77, 256, 130)):
                if kind != O.K_SAT4 and blk in (0, 2):
                    bank.push_events([(v, (5 * v) % frames, GATE, 0.8 if blk == 0 else 0.0) for v in range(n)])
                if blk == 1:
                    bank.push_events([(v, 10 + v % 40, freq_id, 200.0 + 3.0 * v) for v in range(0, n, 3)])
                bank.process_block(frames); o.append(bank.read_voice_out())
            outs.append(np.concatenate(o))
    finally:
        if old is None:
            os.environ.pop("OSCEN_B200_JIT_SEGMENTS", None)
        else:
            os.environ["OSCEN_B200_JIT_SEGMENTS"] = old
    assert np.abs(outs[2]).max() > 0.01
    assert np.array_equal(outs[0], outs[2]), float(np.abs(outs[0] - outs[2]).max())
    assert np.array_equal(outs[1], outs[2]), float(np.abs(outs[1] - outs[2]).max())



@pytest.mark.parametrize("array_smem", ["1", "0", "2"])
def test_generated_electric_piano_with_state_in_shared_memory_and_in_hbm_rows_matches_oracle(array_smem):
    """OSCEN_B200_JIT_ARRAY_SMEM=0|1: where the generated kernel keeps the 259 state words of AmplitudeSource + OscillatorBank
    for the block (HBM rows / a [word][thread] slice of dynamic shared memory): raw MIDI in, stereo frames out, against the
    oracle's electric-piano bank; 70 instances = 1120 voices over 18 CTAs, the last one partly filled"""
    from oscen_b200 import abi
    from test_parity_gpu import _random_midi
    text = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "graphs", "electric_piano.graph")).read()
    n_inst, vpi = 70, 16
    n = n_inst * vpi
    old = os.environ.get("OSCEN_B200_JIT_ARRAY_SMEM")
    try:
        os.environ["OSCEN_B200_JIT_ARRAY_SMEM"] = array_smem
        g = abi.Graph(text)
        gen = abi.Engine.from_graph(g, n, plan=abi.PLAN_FUSED)
        if array_smem != "0":          # and on top of it the variant for the bank-uniform inputs (all but the MIDI front's frequency)
            assert gen.specialize() == (1 << g.n_inputs) - 1 - (1 << g.input_id("frequency"))
    finally:
        if old is None:
            os.environ.pop("OSCEN_B200_JIT_ARRAY_SMEM", None)
        else:
            os.environ["OSCEN_B200_JIT_ARRAY_SMEM"] = old
    ref = O.OracleBank(O.K_EPIANO, n, vpi)
    rng = np.random.default_rng(23)
    for blk, frames in enumerate((256, 130, 256)):
        msgs = _random_midi(rng, n_inst, frames, 200 if blk == 0 else 30, 2 * vpi)
        for b in (ref, gen):
            for inst, frame, data in msgs:
                b.push_midi(inst, frame, data)
            if blk == 1:
                b.set_value(g.input_id("decay_rate"), 55.0)
            b.process_block(frames)
        assert np.array_equal(gen.read_voice_out(), ref.read_voice_out()), blk
        assert np.array_equal(gen.read_instance_out(), ref.read_instance_out()), blk
    assert np.abs(ref.read_instance_out()).max() > 1e-3
    assert ("<uniform=" in gen.last_kernel_name()) == (array_smem != "0")
    gen.close()
