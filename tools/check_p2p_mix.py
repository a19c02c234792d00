#!/usr/bin/env python
"""2+ rank check of the NVLink peer-memory mix bus (run under torchrun on a multi-GPU box):
the fused reduce+exchange kernel must give every rank the sum of all ranks' local mix buses.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_p2p_mix.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oscen_b200 import abi  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", lr))
    n, frames = 4096, 256
    wl = bench.Workload(n * world, rank * n, n)
    local = abi.Engine(abi.GRAPH_FM_VOICE, n, 8, 48000.0, lr, abi.PLAN_FUSED, abi.KEEP_INSTANCE_OUT)
    p2p = abi.Engine(abi.GRAPH_FM_VOICE, n, 8, 48000.0, lr, abi.PLAN_FUSED, abi.KEEP_INSTANCE_OUT)
    mine = torch.frombuffer(bytearray(p2p.comm_p2p_export(world)), dtype=torch.uint8).cuda()
    allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(allh, mine)
    p2p.comm_p2p_import(rank, world, b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
    worst = 0.0
    for b in range(6):
        frames_b = frames if b != 3 else 100            # a short block in the middle
        for e in (local, p2p):
            wl.apply(e, b, abi.EVENT_DTYPE, abi.EVENT_GATE)
            e.process_block(frames_b)
        want = torch.from_numpy(local.read_mix().copy()).cuda()
        dist.all_reduce(want, op=dist.ReduceOp.SUM)
        got = p2p.read_mix()
        err = float(np.abs(got - want.cpu().numpy()).max())
        worst = max(worst, err)
        assert np.allclose(got, want.cpu().numpy(), rtol=1e-5, atol=1e-5), (rank, b, err)
        # every rank must hold the same bits
        g = torch.from_numpy(got.copy()).cuda()
        lo, hi = g.clone(), g.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "ranks disagree on the mix bus"
    if rank == 0:
        print(f"p2p mix bus ok on {world} ranks: max |p2p - allreduce(local)| = {worst:.3e}, peak {float(np.abs(got).max()):.3f}")
    dist.barrier()
    local.close(); p2p.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
