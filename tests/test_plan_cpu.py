"""The plan compiler (oscen_b200/csrc/plan.hpp): graph descriptions -> Kahn topological waves ->
launch sequence.  No GPU needed: checked through oscen_b200_plan_describe()."""
import pytest

from oscen_b200 import abi, build

build.build()

# producers of every node, as the reference graph files connect them
FM_EDGES = {  # examples/fm-synth/src/fm_voice.rs:83-155
    "filter_env_gain": ["env_filter"], "cutoff_mod": ["filter_env_gain"], "op3_osc": ["env3"],
    "op3_route": ["op3_osc"], "op2_osc": ["env2", "op3_route"], "op1_mod_mixer": ["op2_osc", "op3_route"],
    "op1_osc": ["env1", "op1_mod_mixer"], "filter": ["op1_osc", "cutoff_mod"], "output_gain": ["filter"],
}
SYNTH_EDGES = {"carrier": ["modulator"], "filter": ["carrier"], "vca": ["filter", "amp_env"]}   # README.md:44-52
TWIN_EDGES = {"filter_a.output - filter_b.output": ["filter_a", "filter_b"]}                      # nih-twin-peaks lib.rs:35-47
SAT_EDGES = {"clip": ["osc"]}


def _check_topological(plan, edges, n_nodes):
    pos = {}
    for k, (_, names) in enumerate(plan):
        for nm in names:
            assert nm not in pos, f"{nm} launched twice"
            pos[nm] = k
    assert len(pos) == n_nodes
    for node, producers in edges.items():
        for p in producers:
            assert pos[p] < pos[node], f"{node} launched before its producer {p}"


def test_fm_voice_plan_is_a_valid_topological_order_with_one_launch_per_node_type_and_wave():
    plan = abi.plan_describe(abi.GRAPH_FM_VOICE)
    _check_topological(plan, FM_EDGES, 13)
    # the four envelopes are independent sources: one AdsrEnvelope launch renders all of them
    assert plan[0] == ("AdsrEnvelope", ["env3", "env2", "env1", "env_filter"])
    assert [t for t, _ in plan].count("FmOperator") == 3          # op3 -> op2 -> op1 is a dependency chain
    assert plan[-1] == ("Gain", ["output_gain"])
    assert len(plan) == 10


@pytest.mark.parametrize("kind,edges,n", [(abi.GRAPH_SYNTH, SYNTH_EDGES, 3), (abi.GRAPH_SYNTH_ADSR, SYNTH_EDGES, 5),
                                          (abi.GRAPH_TWIN_PEAKS, TWIN_EDGES, 3), (abi.GRAPH_SAT_1X, SAT_EDGES, 2)])
def test_other_plans_are_topological(kind, edges, n):
    plan = abi.plan_describe(kind)
    present = {nm for _, names in plan for nm in names}
    _check_topological(plan, {k: [p for p in v if p in present] for k, v in edges.items() if k in present}, n)


def test_twin_peaks_filters_share_one_launch():
    plan = abi.plan_describe(abi.GRAPH_TWIN_PEAKS)
    assert plan[0] == ("LP18Filter", ["filter_a", "filter_b"])
    assert plan[1][0] == "Sub(expr)"


@pytest.mark.parametrize("kind", [abi.GRAPH_SAT_4X, abi.GRAPH_EPIANO])
def test_multirate_and_array_stream_graphs_are_fused_only(kind):
    with pytest.raises(abi.EngineError):
        abi.plan_describe(kind)
