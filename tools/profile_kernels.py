"""Launch representative kernels once each in isolation (for `ncu --set full`), at bs=32 x 3 s shapes:
MelGAN stage-4 / stage-1 convs, an EBEN-discriminator strided grouped conv, the generator residual-unit convs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vibravox_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda"
torch.manual_seed(0)
CASES = [  # Cin, Cout, K, stride, dil, pad, refl, groups, Tin
    (1024, 1024, 41, 4, 1, 20, 0, 4, 748),     # MelGAN stage 4
    (16, 64, 41, 4, 1, 20, 0, 4, 47840),       # MelGAN stage 1
    (384, 768, 7, 2, 2, 6, 0, 4, 743),         # EBEN discriminator, strided grouped
    (32, 32, 3, 1, 3, 3, 3, 1, 11968),         # generator residual unit, dilated k3
    (64, 64, 1, 1, 1, 0, 0, 1, 5984),          # generator residual unit, pointwise
]
torch.cuda.cudart().cudaProfilerStart()
for (ci, co, k, s, d, p, r, gr, T) in CASES:
    g = ops.ConvGeom(ci, co, k, s, d, p, r, gr)
    x = torch.randn(B, ci, T, device=dev)
    w = torch.randn(co, ci // gr, k, device=dev) * 0.05
    y = ops.conv_fwd(x, w, g, slope=0.2)
    dx = ops.conv_dgrad(y, w, None, g, T)
    dw = ops.conv_wgrad(x, y, g)
    torch.cuda.synchronize()
    del x, w, y, dx, dw
torch.cuda.cudart().cudaProfilerStop()
print("done")
