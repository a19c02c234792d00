"""N>1 host logic on CPU (world_size 2, gloo): voices are sharded in contiguous instance ranges,
no state moves, the only exchange is the all-reduce of the mix bus.  Each rank renders its shard
with the CPU oracle standing in for the GPU engine (same driving surface), exactly as bench.py
shards its workload; the reduced mix must equal the single-process render of the whole bank."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import oracle_lib as O
    import bench
    dist.init_process_group(backend="gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_total, frames = 256, 256
    n = n_total // world
    wl = bench.Workload(n_total, rank * n, n)                 # this rank's shard of the global bank
    bank = O.OracleBank(O.K_FM, n, 8)
    dt = np.dtype([("voice", "<u4"), ("frame_offset", "<u4"), ("input_id", "<u4"), ("scalar", "<f4")])
    mixes = []
    for b in range(3):
        wl.apply(bank, b, dt, O.EV_GATE)
        bank.process_block(frames)
        m = torch.from_numpy(bank.read_mix().copy())
        dist.all_reduce(m, op=dist.ReduceOp.SUM)              # the one exchange step
        mixes.append(m.numpy())
    if rank == 0:
        np.save({out!r}, np.concatenate(mixes))
    dist.destroy_process_group()
""")


def test_two_rank_sharding_matches_single_process(tmp_path):
    out = str(tmp_path / "mix.npy")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, out=out))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)], env=env,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)
    got = np.load(out)

    sys.path.insert(0, ROOT)
    import bench
    import oracle_lib as O
    wl = bench.Workload(256, 0, 256)
    bank = O.OracleBank(O.K_FM, 256, 8)
    dt = np.dtype([("voice", "<u4"), ("frame_offset", "<u4"), ("input_id", "<u4"), ("scalar", "<f4")])
    ref = []
    for b in range(3):
        wl.apply(bank, b, dt, O.EV_GATE)
        bank.process_block(256)
        ref.append(bank.read_mix())
    ref = np.concatenate(ref)
    assert np.abs(ref).max() > 1e-3
    # two partial sums added instead of one sequential sum: tolerance, not bit equality
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-5)


def test_workload_shards_tile_the_global_bank():
    sys.path.insert(0, ROOT)
    import bench
    whole = bench.Workload(512, 0, 512)
    parts = [bench.Workload(512, r * 128, 128) for r in range(4)]
    assert np.array_equal(np.concatenate([p.freq for p in parts]), whole.freq)
    assert np.array_equal(np.concatenate([p.on for p in parts]), whole.on)
    assert np.array_equal(np.concatenate([p.vel for p in parts]), whole.vel)
