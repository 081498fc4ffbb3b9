"""Bring-up helper for the fused ResidualUnit kernel: one tiny launch, checked against torch fp64.  Run it per
VBX_RU_DBG setting in separate processes (a device-side trap kills the context)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from vibravox_b200 import ops

B, C, T, d = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (1, 32, 256, 1))]
torch.manual_seed(0)
x = torch.randn(B, C, T, device="cuda")
w1 = torch.randn(C, C, 3, device="cuda") / (3 * C) ** 0.5
w2 = torch.randn(C, C, 1, device="cuda") / C ** 0.5
pk = ops.residual_unit_pack(w1, w2)
torch.cuda.synchronize()
print("pack ok", flush=True)
out, h, mask = ops.residual_unit_fwd(x, pk, d, 0.01, want_h=True, want_mask=True)
torch.cuda.synchronize()
print("launch ok", flush=True)
x64 = x.double()
h64 = F.conv1d(F.pad(x64, (d, d), mode="reflect"), w1.double(), None, 1, 0, d)
want = x64 + F.leaky_relu(F.conv1d(h64, w2.double()), 0.01)
print("h err", float((h.double() - h64).norm() / h64.norm()), "out err", float((out.double() - want).norm() / want.norm()), flush=True)
