#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "unit_weight_gradient and 1000" > gpurun_out/p18_pytest_small.txt 2>&1
tail -15 gpurun_out/p18_pytest_small.txt | cut -c1-250
if grep -q "passed" gpurun_out/p18_pytest_small.txt && ! grep -q "failed\|error" gpurun_out/p18_pytest_small.txt; then
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "unit_weight_gradient or fused_residual" > gpurun_out/p18_pytest.txt 2>&1
tail -5 gpurun_out/p18_pytest.txt | cut -c1-250
timeout 120 python tools/wg_bench.py > gpurun_out/p18_wgbench.txt 2>&1; cat gpurun_out/p18_wgbench.txt
fi
