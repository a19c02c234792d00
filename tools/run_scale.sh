#!/bin/bash
# 1/2/4/8-GPU weak-scaling check of bench.py on one box (run under gpurun --gpus 8)
mkdir -p gpurun_out
timeout 200 python bench.py --gpus 1 --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
for N in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --steps 100 --warmup 20 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  echo "N=$N rc=$?"
done
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    try:
        lines = open(f"gpurun_out/scale_{n}.json").read().strip().splitlines()
        d = json.loads(lines[-1])
        print(n, "lines", len(lines), "value %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], d["config"]["parallelism"], d["clocks"]["reasons"])
    except Exception as ex:
        print(n, "FAILED", ex)
PY
tail -3 gpurun_out/scale_8.err
