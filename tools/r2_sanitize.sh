#!/bin/bash
# compute-sanitizer passes over the round-2 kernels (small shapes; one tool per run).  --report-api-errors no: with lazy module
# loading the runtime probes cuKernelGetFunction on the first launch of each kernel and the tool reports that (benign) API
# return code as an error.
mkdir -p gpurun_out
O=gpurun_out
K='gate_stage and (1500 or 2500 or 375 or 1001 or 131 or 300) or conv_family_matches_torch and (2500 or 2503 or 375 or 1001 or 1111) or residual_unit_composed or fused_residual_unit_through'
timeout 900 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -k "$K" > $O/r2_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> $O/r2_sanitizer_memcheck.txt
tail -6 $O/r2_sanitizer_memcheck.txt | cut -c1-200
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -k "conv_family_matches_torch and (2500 or 375 or 1001)" > $O/r2_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> $O/r2_sanitizer_racecheck.txt
tail -6 $O/r2_sanitizer_racecheck.txt | cut -c1-200
