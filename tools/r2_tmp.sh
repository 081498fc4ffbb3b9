#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "conv_family_matches_torch or gate_stage" 2>&1 | tail -2 | cut -c1-300
timeout 300 python tools/layer_table.py 2>/dev/null | grep -E "summed|1>16 " | cut -c1-160
B="python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-cpu-baseline --no-micro"
timeout 300 $B 2> /dev/null | grep -o '"ms_per_step": [0-9.]*'
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -2 | cut -c1-300
timeout 300 python tools/step_timeline.py graph > $O/c6_step_timeline.txt 2>&1; rm -f $O/step_trace.json
