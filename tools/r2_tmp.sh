#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "residual_unit" 2>&1 | tail -3 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3 | cut -c1-300
B="python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-cpu-baseline --no-micro"
echo composed; timeout 300 $B 2> /dev/null | grep -o '"ms_per_step": [0-9.]*'
echo layerwise; VBX_UNIT_COMPOSED=0 timeout 300 $B 2> /dev/null | grep -o '"ms_per_step": [0-9.]*'
