"""One process per GPU: wiring of the only exchange step of the path, the mix bus.

Voices (whole instances) are sharded in contiguous ranges, one engine per rank; no state and no per-voice output ever
moves.  `connect_mix_bus` gives an engine its collective: the fused reduce+exchange kernel over NVLink peer memory
(`oscen_b200_comm_p2p_*`, CUDA IPC handles all-gathered through torch.distributed) or `ncclAllReduce`
(`oscen_b200_comm_init`, the unique id broadcast from rank 0).  torch.distributed is the rendezvous plumbing only."""
import sys

from . import abi


def shard_range(n_total, rank, world, granule=1):
    """contiguous [first, first + count) of `n_total` units for `rank`; whole granules (e.g. an instance's voices) stay together"""
    g_total = n_total // granule
    base, extra = divmod(g_total, world)
    first = rank * base + min(rank, extra)
    count = base + (1 if rank < extra else 0)
    return first * granule, count * granule


def connect_mix_bus(eng, rank, world, mode="p2p"):
    """mode "p2p" | "nccl"; returns the mode in effect on ALL ranks ("local" for world == 1).  Collective: every rank calls it."""
    if world <= 1:
        return "local"
    import torch
    import torch.distributed as dist
    if mode == "nccl":
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(abi.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        eng.comm_init(rank, world, bytes(idt.cpu().numpy().tobytes()))
        dist.barrier()
        return "nccl"
    ok = 1
    try:
        mine = torch.frombuffer(bytearray(eng.comm_p2p_export(world)), dtype=torch.uint8).cuda()
    except abi.EngineError as ex:
        print(f"[rank {rank}] p2p mix exchange unavailable ({ex})", file=sys.stderr)
        mine = torch.zeros(64, dtype=torch.uint8, device="cuda"); ok = 0
    allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(allh, mine)
    flag = torch.tensor([ok], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()):
        try:
            eng.comm_p2p_import(rank, world, b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
        except abi.EngineError as ex:          # IPC mapping refused (no peer access): every rank falls back together
            print(f"[rank {rank}] p2p mix exchange unavailable ({ex})", file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()):
            dist.barrier()                     # no rank renders before every rank has reset and mapped its exchange buffer
            return "p2p"
        if ok:
            eng.comm_p2p_close()
    return connect_mix_bus(eng, rank, world, "nccl")
