#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "conv_family_matches_torch" 2>&1 | tail -2 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_chain" 2>&1 | grep -E "assert|Error|passed|failed|launches" | head -20 | cut -c1-300
timeout 300 python tools/layer_table.py 2>/dev/null | grep -E "summed|768>1 |1024>1 |32>4 |4>24 |1>16 " | grep -E "summed|wgrad|fwd" | cut -c1-160
B="python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-cpu-baseline --no-micro"
echo "skinny on";  timeout 300 $B 2> /dev/null | grep -o '"ms_per_step": [0-9.]*'
cat > /tmp/d0.py <<'PY'
import sys; sys.path.insert(0,'.')
import torch
from vibravox_b200 import ops
g = ops.ConvGeom(1, 16, 15, 1, 1, 7, 7, 1)
x = torch.randn(32, 1, 47840, device='cuda'); w = torch.randn(16, 1, 15, device='cuda'); b = torch.zeros(16, device='cuda')
for _ in range(3): y = ops.conv_fwd(x, w, g, bias=b, slope=0.2)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:direct_fwd_kernel -s 1 -c 1 -o $O/c4_direct_fwd python /tmp/d0.py > /dev/null 2>&1
ls -la $O/c4_direct_fwd.ncu-rep
