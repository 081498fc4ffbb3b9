#!/bin/bash
# The `ncu --set full` captures behind profiles/r2_ncu_full_kernels.csv (extract here with tools/ncu_extract.py).
mkdir -p gpurun_out
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ru_fwd_kernel -s 3 -c 1 -o $O/r2_ncu_ru_fwd_c32 python tools/ru_bench.py 32 11968 3 6 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ru_fwd_kernel -s 3 -c 1 -o $O/r2_ncu_ru_fwd_c64 python tools/ru_bench.py 64 5984 9 6 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_slab_kernel -c 1 -o $O/r2_ncu_slab_melgan4 python -c "
import sys; sys.path.insert(0,'.')
import torch
from vibravox_b200 import ops
g = ops.ConvGeom(1024, 1024, 41, 4, 1, 20, 0, 4)
x = torch.randn(32, 1024, 748, device='cuda'); w = torch.randn(1024, 256, 41, device='cuda') * 0.01; b = torch.zeros(1024, device='cuda')
for _ in range(3): ops.conv_fwd(x, w, g, bias=b, slope=0.2)
torch.cuda.synchronize()" > /dev/null 2>&1
python tools/ru_bench.py > $O/r2_ru_bench.txt 2>&1; RU_TRAIN=1 python tools/ru_bench.py >> $O/r2_ru_bench.txt 2>&1
python tools/wg_bench.py > $O/r2_wg_bench.txt 2>&1
python tools/ru_timeline.py 32 11968 3 > $O/r2_ru_timeline_c32.txt 2>&1
python tools/ru_timeline.py 64 5984 9 > $O/r2_ru_timeline_c64.txt 2>&1
