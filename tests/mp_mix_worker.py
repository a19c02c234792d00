"""Worker of tests/test_multigpu_gpu.py (one process per GPU, launched by torch.distributed.run): renders a voice-sharded
FM bank through the C ABI with the mix bus exchanged (a) by the fused reduce+exchange kernel over NVLink peer memory and
(b) by ncclAllReduce, fused plan and node plan, and checks on every rank, every block:

  * every rank holds the SAME BITS of the mix bus (p2p: ranks fold in rank order; gathered and compared);
  * it equals all_reduce(sum) of the ranks' local (no-comm) mix buses to 1e-5;
  * rank 0: it equals the oracle's mix of the UNION bank within the summation-order tolerance (2e-6 * sum |instance|),
    the reference's `voices.audio_out -> audio_out` sum (oscen-graph-compiler/src/codegen/emit_node.rs:463-466);
  * close + re-import of the peer exchange starts a clean sequence (no stale flag folds old data).

Prints one JSON line from rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import oracle_lib as O  # noqa: E402
from oscen_b200 import abi, shard  # noqa: E402


def gather_bits(x):
    """every rank's copy of x (float32 array) on every rank"""
    t = torch.from_numpy(np.ascontiguousarray(x).view(np.int32).copy()).cuda()
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.cpu().numpy() for o in out]


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", lr))
    n_total, vpi = 8192 * world, 8
    first, n = shard.shard_range(n_total, rank, world, vpi)
    wl = bench.Workload(n_total, first, n)
    blocks = [256, 256, 100, 256]
    result = {"world": world, "cases": []}

    # the oracle's union bank (rank 0 only) and every rank's local, comm-less render
    ref_mix, ref_scale = [], []
    if rank == 0:
        whole = bench.Workload(n_total, 0, n_total)
        ob = O.OracleBank(O.K_FM, n_total, vpi, 48000.0, threads=os.cpu_count() or 1)
        for b, fr in enumerate(blocks):
            whole.apply(ob, b, abi.EVENT_DTYPE, O.EV_GATE)
            ob.process_block(fr)
            inst = ob.read_instance_out().astype(np.float64)[:, :, 0]
            ref_mix.append(inst.sum(axis=1)); ref_scale.append(np.abs(inst).sum(axis=1))
    local = abi.Engine(abi.GRAPH_FM_VOICE, n, vpi, 48000.0, lr, abi.PLAN_FUSED, abi.KEEP_INSTANCE_OUT)
    want = []
    for b, fr in enumerate(blocks):
        wl.apply(local, b, abi.EVENT_DTYPE, abi.EVENT_GATE)
        local.process_block(fr)
        w = torch.from_numpy(local.read_mix().astype(np.float64)).cuda()
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        want.append(w.cpu().numpy()[:, 0])
    local.close()

    for plan, mode, reimport in ((abi.PLAN_FUSED, "p2p", False), (abi.PLAN_FUSED, "nccl", False), (abi.PLAN_NODES, "p2p", False),
                                 (abi.PLAN_NODES, "nccl", False), (abi.PLAN_FUSED, "p2p", True)):
        eng = abi.Engine(abi.GRAPH_FM_VOICE, n, vpi, 48000.0, lr, plan, abi.KEEP_INSTANCE_OUT)
        got_mode = shard.connect_mix_bus(eng, rank, world, mode)
        case = {"plan": "fused" if plan == abi.PLAN_FUSED else "nodes", "mode": got_mode, "reimport": reimport,
                "ranks_identical": True, "max_abs_vs_allreduce": 0.0, "max_rel_vs_oracle_scaled": 0.0}
        for b, fr in enumerate(blocks):
            if reimport and b == 2:
                # drop the exchange and map it again mid-run: sequence numbers restart, old flags / data must not be folded
                eng.synchronize(); dist.barrier()
                eng.comm_p2p_close(); dist.barrier()
                assert shard.connect_mix_bus(eng, rank, world, "p2p") == got_mode
            wl.apply(eng, b, abi.EVENT_DTYPE, abi.EVENT_GATE)
            eng.process_block(fr)
            got = eng.read_mix()[:, 0]
            allr = gather_bits(got)
            if mode == "p2p":
                case["ranks_identical"] &= all(np.array_equal(a, allr[0]) for a in allr)
            err = float(np.abs(got.astype(np.float64) - want[b]).max())
            case["max_abs_vs_allreduce"] = max(case["max_abs_vs_allreduce"], err)
            assert np.allclose(got, want[b], rtol=1e-5, atol=1e-5), (case, b, err)
            if rank == 0:
                d = np.abs(got.astype(np.float64) - ref_mix[b])
                tol = 1e-6 + 2e-6 * ref_scale[b]
                case["max_rel_vs_oracle_scaled"] = max(case["max_rel_vs_oracle_scaled"], float((d / tol).max()))
                assert (d <= tol).all(), (case, b, float((d / tol).max()))
        assert case["ranks_identical"], case
        eng.synchronize(); dist.barrier()
        eng.close()
        result["cases"].append(case)
    if rank == 0:
        result["peak"] = float(max(np.abs(m).max() for m in ref_mix))
        assert result["peak"] > 1.0
        print("MPMIX " + json.dumps(result), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
