#!/bin/bash
# round-2 call N (2 GPUs): multi-GPU tests and the N = 2 bench line with the final library (p2p and NCCL mix bus)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q > gpurun_out/r02n_tests.log 2>&1; echo "multi-GPU tests rc=$?"; tail -2 gpurun_out/r02n_tests.log
for m in p2p nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --mix $m --no-other-configs > gpurun_out/r02n_bench_n2_$m.json 2> gpurun_out/r02n_bench_n2_$m.err; echo "N=2 $m rc=$?"
python - "$m" <<'PY'
import json, sys
m = sys.argv[1]
d = json.loads(open(f"gpurun_out/r02n_bench_n2_{m}.json").read().strip().splitlines()[-1])
print(m, "value %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], d.get("mix_check"))
PY
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1 same box: value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'])"
