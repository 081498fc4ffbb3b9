#!/bin/bash
B="python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-cpu-baseline --no-micro"
echo "overlap on"; timeout 300 $B 2> /dev/null | grep -o '"ms_per_step": [0-9.]*'
echo "overlap off"; VBX_PHASE_OVERLAP=0 timeout 300 $B 2> /dev/null | grep -o '"ms_per_step": [0-9.]*'
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | grep -E "^E|FAILED|Error|passed|failed" | head -30 | cut -c1-400
