#!/bin/bash
# round-2 call AA: compute-sanitizer over what changed after call M — specialised generated kernels, array-node state in shared
# memory (generated ElectricPiano), segmented Synth kernel, pipelined instance_mix_node_kernel (node plan)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_specialize.py tests/test_variants_gpu.py -m gpu -q -x -k "specialis or shared_memory or 130 or segments" > gpurun_out/r02aa_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02aa_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "synth or node_plan or nodes" > gpurun_out/r02aa_memcheck2.log 2>&1; echo "memcheck2 rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02aa_memcheck2.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_variants_gpu.py -m gpu -q -x -k "shared_memory" > gpurun_out/r02aa_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r02aa_racecheck.log | tail -3
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_variants_gpu.py -m gpu -q -x -k "node_plan_kernel_variants and 136" > gpurun_out/r02aa_racecheck2.log 2>&1; echo "racecheck2 rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r02aa_racecheck2.log | tail -3
