#!/bin/bash
# round-2 call F (8 GPUs, one box): multi-GPU tests, then the driver's scaling command at N = 1, 2, 4, 8 (20 steps / 5 warm-up, as SCALE_rNN)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q > gpurun_out/r02f_tests.log 2>&1; echo "multi-GPU tests rc=$?"; tail -3 gpurun_out/r02f_tests.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02f_scale_1.json 2> gpurun_out/r02f_scale_1.err; echo "N=1 rc=$?"
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02f_scale_$N.json 2> gpurun_out/r02f_scale_$N.err
  echo "N=$N rc=$?"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 8 --steps 100 --warmup 20 --no-other-configs > gpurun_out/r02f_scale_8_100steps.json 2> gpurun_out/r02f_scale_8_100steps.err
python - <<'PY'
import json
base = None
for n in (1, 2, 4, 8, "8_100steps"):
    try:
        lines = open(f"gpurun_out/r02f_scale_{n}.json").read().strip().splitlines()
        d = json.loads(lines[-1])
        if n == 1: base = d["value"]
        N = d["n_gpus"]
        print(n, "value %.4g" % d["value"], "eff %.3f" % (d["value"] / (N * base)), "ms/step %.4f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], d.get("mix_check"), d["clocks"]["reasons"])
        for o in d.get("other_configs", []):
            print("    ", o["config"][:70], o["scaling"], "%.4g" % o["value"], "ms %.3f" % o["ms_per_step"], o.get("mix_check"))
    except Exception as ex:
        print(n, "FAILED", ex)
PY
tail -3 gpurun_out/r02f_scale_8.err
