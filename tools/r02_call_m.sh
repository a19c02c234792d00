#!/bin/bash
# round-2 call M: the three ways (generated kernel / node plan / hand-written kernel) of every reference graph with the final library;
# compute-sanitizer (memcheck, racecheck) over the kernel-variant tests and the DSL electric piano
mkdir -p gpurun_out
for V in 65536 1048576; do python tools/bench_dsl.py 10 $V 2>/dev/null; done > gpurun_out/r02m_bench_dsl_three_ways.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/r02m_bench_dsl_three_ways.jsonl"):
    d=json.loads(l); print(d["voices"], d["graph"], d["way"][:24], "%.4g v-s/s %.3f ms" % (d["voice_samples_per_s"], d["kernel_ms"]))
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_variants_gpu.py -m gpu -q -x -k "256-env0 or 136 or tma_kernels or handoff_matches_oracle[64]" > gpurun_out/r02m_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02m_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_dsl_gpu.py tests/test_parity_gpu.py -m gpu -q -x -k "electric_piano or saturator or twin_peaks_host_stream or clip_oversampled" > gpurun_out/r02m_memcheck2.log 2>&1; echo "memcheck2 rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02m_memcheck2.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_variants_gpu.py -m gpu -q -x -k "136 or tma_kernels" > gpurun_out/r02m_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r02m_racecheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_dsl_gpu.py tests/test_parity_gpu.py -m gpu -q -x -k "electric_piano" > gpurun_out/r02m_racecheck2.log 2>&1; echo "racecheck2 rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r02m_racecheck2.log | tail -3
