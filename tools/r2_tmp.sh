#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tensor_core_conv_family" > gpurun_out/p21_pytest.txt 2>&1
tail -4 gpurun_out/p21_pytest.txt | cut -c1-300
if grep -q "passed" gpurun_out/p21_pytest.txt && ! grep -q "failed\|rror" gpurun_out/p21_pytest.txt; then
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-micro 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['ms_per_step'], d['value'], d['e2e']['losses'])" > gpurun_out/p21_bench.txt; cat gpurun_out/p21_bench.txt
timeout 300 python tools/layer_table.py 2>/dev/null | grep -E "wgrad|summed" | head -40 > gpurun_out/p21_wgrad_table.txt; cat gpurun_out/p21_wgrad_table.txt
fi
