#!/bin/bash
timeout 300 python tools/layer_table.py 2>/dev/null > /tmp/lt.txt; grep -E "summed|1024>1024 k41|256>1024 k41|16>64 k41" /tmp/lt.txt | cut -c1-150
timeout 300 python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-cpu-baseline --no-micro 2> /dev/null | grep -o '"ms_per_step": [0-9.]*'
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "tensor_core_conv_family or persistent or gate_stage" 2>&1 | tail -1
