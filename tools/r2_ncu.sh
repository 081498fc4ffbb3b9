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
# late-round kernels: MelGAN stage 0 quad kernels, a certainty conv (forward + weight gradient), MelGAN stage 1 forward
cat > /tmp/late.py <<'PY'
import sys; sys.path.insert(0,'.')
import torch
from vibravox_b200 import ops
g0 = ops.ConvGeom(1, 16, 15, 1, 1, 7, 7, 1)
x = torch.randn(32, 1, 47840, device='cuda'); w = torch.randn(16, 1, 15, device='cuda'); b = torch.zeros(16, device='cuda')
for _ in range(2): y = ops.conv_fwd(x, w, g0, bias=b, slope=0.2)
dy = torch.randn_like(y)
for _ in range(2): dx = ops.conv_dgrad(dy, w, None, g0, 47840)
gc = ops.ConvGeom(768, 1, 3, 1, 1, 1, 0, 1)
xc = torch.randn(32, 768, 375, device='cuda'); wc = torch.randn(1, 768, 3, device='cuda') * 0.02; bc = torch.zeros(1, device='cuda')
for _ in range(2): yc = ops.conv_fwd(xc, wc, gc, bias=bc)
for _ in range(2): dwc = ops.conv_wgrad(xc, torch.randn_like(yc), gc)
g1 = ops.ConvGeom(16, 64, 41, 4, 1, 20, 0, 4)
x1 = torch.randn(32, 16, 47840, device='cuda'); w1 = torch.randn(64, 4, 41, device='cuda') * 0.1; b1 = torch.zeros(64, device='cuda')
for _ in range(2): y1 = ops.conv_fwd(x1, w1, g1, bias=b1, slope=0.2)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:"direct_fwd4_kernel|direct_dgrad4_kernel|skinny_fwd_kernel|skinny_wgrad_kernel|tc_slab_kernel|tc_pslab_kernel" -o $O/r2_ncu_late python /tmp/late.py > /dev/null 2>&1
ls -la $O/r2_ncu_late.ncu-rep
