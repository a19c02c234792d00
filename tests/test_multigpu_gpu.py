"""N > 1 on real GPUs: the voice-sharded mix bus through the C ABI, both exchange implementations and both plans, against
the oracle's union bank and bit-for-bit across ranks (tests/mp_mix_worker.py).  Needs >= 2 GPUs on the box: skipped on the
1-GPU lease the driver's test tier uses — bench.py runs the same check (`mix_check`) inside every N > 1 bench line, so the
driver's scaling run covers it as well."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import ctypes
        lib = ctypes.CDLL("libcuda.so.1")
        if lib.cuInit(0) != 0:
            return 0
        n = ctypes.c_int(0)
        return n.value if lib.cuDeviceGetCount(ctypes.byref(n)) == 0 else 0
    except OSError:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs on one box")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_mix_bus_across_ranks_matches_oracle_and_is_bit_identical_on_every_rank(world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        os.path.join(ROOT, "tests", "mp_mix_worker.py")], env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("MPMIX ")]
    assert line, p.stdout[-2000:]
    res = json.loads(line[-1][6:])
    assert res["world"] == world and len(res["cases"]) == 5
    modes = {(c["plan"], c["mode"]) for c in res["cases"]}
    assert ("fused", "nccl") in modes and ("nodes", "nccl") in modes
    for c in res["cases"]:
        assert c["ranks_identical"] and c["max_rel_vs_oracle_scaled"] <= 1.0, c
    print(json.dumps(res))
