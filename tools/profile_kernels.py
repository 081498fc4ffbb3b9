"""Launch a few representative kernels in isolation (for `ncu --set full`): the dense MelGAN stage-4
conv (fwd / dgrad / wgrad) and the generator residual-unit convs at bs=32 x 3 s shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vibravox_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda"
torch.manual_seed(0)
# MelGAN stage 4 (1024->1024, k41, s4, g4) on T=748
g = ops.ConvGeom(1024, 1024, 41, 4, 1, 20, 0, 4)
x = torch.randn(B, 1024, 748, device=dev)
w = torch.randn(1024, 256, 41, device=dev) * 0.01
wt = ops.transpose_weight(w, 4)
bias = torch.zeros(1024, device=dev)
for _ in range(2):
    y = ops.conv_fwd(x, w, g, bias=bias, slope=0.2)
    dx = ops.conv_dgrad(y, w, wt, g, 748)
    dw = ops.conv_wgrad(x, y, g)
# generator residual unit at C=32, T=11968
C, T = 32, 11968
xg = torch.randn(B, C, T, device=dev)
w1 = torch.randn(C, C, 3, device=dev) * 0.1
w2 = torch.randn(C, C, 1, device=dev) * 0.1
g1 = ops.ConvGeom(C, C, 3, 1, 3, 3, 3, 1)
g2 = ops.ConvGeom(C, C, 1, 1, 1, 0, 0, 1)
for _ in range(2):
    h = ops.conv_fwd(xg, w1, g1)
    o = ops.conv_fwd(h, w2, g2, res=xg, slope=0.01)
    dh = ops.conv_dgrad(o, w2, None, g2, T)
    dwg = ops.conv_wgrad(xg, dh, g1)
torch.cuda.synchronize()
print("done")
