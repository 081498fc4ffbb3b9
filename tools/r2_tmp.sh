#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "residual_unit" 2>&1 | tail -1 | cut -c1-300
B="python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-cpu-baseline --no-micro"
timeout 300 $B 2> /dev/null | grep -o '"ms_per_step": [0-9.]*'
