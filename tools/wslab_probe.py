"""One wide weight-gradient launch (MelGAN stage 4: 1024 -> 1024, k41, s4, groups 4, B x 748) for ncu / timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vibravox_b200 import ops
B = 32
g = ops.ConvGeom(1024, 1024, 41, 4, 1, 20, 0, 4)
x = torch.randn(B, 1024, 748, device="cuda")
dy = torch.randn(B, 1024, g.tout(748), device="cuda")
dw = torch.zeros(1024, 256, 41, device="cuda")
for _ in range(3):
    ops.tc_conv1d_wgrad(x, dy, g, dw=dw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.tc_conv1d_wgrad(x, dy, g, dw=dw)
e1.record(); torch.cuda.synchronize()
print(f"melgan-4 wgrad: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us", {k: v for k, v in os.environ.items() if k.startswith('VBX_')})
