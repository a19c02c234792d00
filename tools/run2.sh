#!/bin/bash
# 2-GPU validation: p2p mix check + bench with both mix-bus modes (run under gpurun --gpus 2)
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_p2p_mix.py > gpurun_out/p2p_check.log 2>&1
grep "p2p mix bus\|Error\|error\|assert" gpurun_out/p2p_check.log | tail -5
for m in p2p nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 100 --warmup 20 --mix $m > gpurun_out/b2_$m.json 2> gpurun_out/b2_$m.err
  python - "$m" <<'PY'
import json, sys
m = sys.argv[1]
lines = open(f"gpurun_out/b2_{m}.json").read().strip().splitlines()
d = json.loads(lines[-1])
print(m, "stdout_lines", len(lines), "value %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"], "e2e %.4g" % d["e2e"]["value"])
PY
done
