#!/bin/bash
# why is the device-timed step at N=8 twice the N=4 one?  (run under gpurun --gpus 8)
mkdir -p gpurun_out
run() { # name, extra env, extra args
  env $2 OSCEN_BENCH_DUMP=gpurun_out/diag_$1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 8 --steps 100 --warmup 20 $4 > gpurun_out/diag_$1.json 2> gpurun_out/diag_$1.err
  python - "$1" <<'PY'
import json, sys, glob
import numpy as np
name = sys.argv[1]
d = json.loads(open(f"gpurun_out/diag_{name}.json").read().strip().splitlines()[-1])
print(name, "value %.4g ms/step %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
for f in sorted(glob.glob(f"gpurun_out/diag_{name}.rank*.json")):
    r = json.load(open(f)); s = np.array(r["step_ms"]); k = np.array(r["kern_ms"])
    print("   ", f[-10:-5], "step p50 %.3f p90 %.3f max %.3f | kernel p50 %.3f p90 %.3f max %.3f | wall/step %.3f" % (np.median(s), np.percentile(s, 90), s.max(), np.median(k), np.percentile(k, 90), k.max(), 1e3 * r["wall_s"] / len(s)))
PY
}
run p2p "A=1" 29711 ""
run p2p_nosampler "OSCEN_BENCH_NO_SAMPLER=1" 29712 ""
run nccl "A=1" 29713 "--mix nccl"
