import sys, os
sys.path.insert(0, "/root/repo")
import torch
from vibravox_b200 import ops
B=32
for (ci, co, k, s, d, p, r, gr, T) in [(1024, 1024, 41, 4, 1, 20, 0, 4, 748), (256, 1024, 41, 4, 1, 20, 0, 4, 2990)]:
    g = ops.ConvGeom(ci, co, k, s, d, p, r, gr)
    x = torch.randn(B, ci, T, device="cuda"); y = torch.randn(B, co, g.tout(T), device="cuda")
    for _ in range(2): dw = ops.conv_wgrad(x, y, g)
    torch.cuda.synchronize()
