"""GPU parity tests of `external NAME: AudioAsset;` slots and SamplePlayer (oscen_b200_load_asset + the generated
kernel), the same scripts as tests/test_host_kernel_cpu.py runs on the host-executed kernel.
(The file name sorts last on purpose: added after the round's GPU budget was spent — the kernel text is checked on the
host, the engine-side slot table and loads were not yet run on a device.)"""
import os

import numpy as np
import pytest

import graph_walk as W
from test_host_kernel_cpu import PLAYERS_DSL

pytestmark = pytest.mark.gpu
GRAPHS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "graphs")
f32 = np.float32


@pytest.fixture(params=[0, 1], ids=["generated_fused_kernel", "node_plan"])
def plan(request):
    return request.param


def test_sample_player_graph_reproduces_the_reference_graph_test_on_the_gpu(plan):
    """oscen-lib/tests/sample_player_graph.rs:64-124 for every voice of a 1000-voice bank"""
    from oscen_b200 import abi
    n = 1000
    a = np.array([0.1, -0.2, 0.3, -0.4, 0.5], dtype=np.float32)
    b = np.array([0.9, 0.8, 0.7], dtype=np.float32)
    eng = abi.Engine.from_graph(open(os.path.join(GRAPHS, "player.graph")).read(), n, 1, 44100.0, plan=plan)
    eng.process_block(8)
    assert (eng.read_voice_out() == 0.0).all()
    eng.load_asset("clip", a)
    eng.process_block(7); y1 = eng.read_voice_out()
    eng.process_block(3); y2 = eng.read_voice_out()
    assert np.array_equal(np.concatenate([y1, y2]), np.repeat(np.concatenate([a, a])[:, None], n, axis=1))
    eng.load_asset("clip", np.stack([b, -b], axis=1))           # stereo source: a mono player takes channel 0
    eng.process_block(6)
    assert np.array_equal(eng.read_voice_out(), np.repeat(np.concatenate([b, b])[:, None], n, axis=1))
    eng.load_asset("clip", np.zeros(0, dtype=np.float32))
    eng.process_block(4)
    assert (eng.read_voice_out() == 0.0).all()
    with pytest.raises(abi.EngineError):
        eng.load_asset(3, a)                                        # the graph declares one slot


def test_two_sample_players_through_a_filter_match_the_oracle_walk_on_the_gpu(plan):
    from oscen_b200 import abi
    n, sr = 5, 48000.0
    rng = np.random.default_rng(7)
    bufs = {k: rng.uniform(-1, 1, size).astype(np.float32) for k, size in (("a1", 37), ("b1", 11), ("a2", 5))}
    level = (0.3 + 0.1 * np.arange(n)).astype(np.float32)
    cutoff = (500.0 + 300.0 * np.arange(n)).astype(np.float32)
    script = [(20, {}), (50, {"loop_a": "a1"}), (41, {"loop_b": "b1"}), (64, {"loop_a": "a2"})]
    g = abi.Graph(PLAYERS_DSL)
    eng = abi.Engine.from_graph(g, n, 1, sr, plan=plan)
    eng.set_value(g.input_id("level"), level)
    eng.set_value(g.input_id("cutoff"), cutoff)
    got = []
    for frames, loads in script:
        for slot, key in loads.items():
            eng.load_asset(slot, bufs[key])
        eng.process_block(frames); got.append(eng.read_voice_out())
    got = np.concatenate(got)
    want = np.zeros_like(got)
    for v in range(n):
        pa, pb, idle, filt = W.sample_player(), W.sample_player(), W.sample_player(), W.tpt(900.0, 0.8, sr)
        slots = {"loop_a": pa, "loop_b": pb}
        f = 0
        for frames, loads in script:
            for slot, key in loads.items():
                slots[slot].publish(bufs[key])
            for _ in range(frames):
                a, b, i = pa.process(), pb.process(), idle.process()
                filt.set(input=f32(f32(a * level[v]) + i), cutoff=cutoff[v])
                want[f, v] = f32(filt.process() + f32(b * f32(0.25)))
                f += 1
    assert np.array_equal(got, want), float(np.abs(got - want).max())
