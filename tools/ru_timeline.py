"""Per-tile pipeline timeline of CTA 0 of the fused ResidualUnit kernel (vbx_ru_set_profile_buffer): cycles between
the pipeline events of each tile.  usage: python tools/ru_timeline.py [C T d]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vibravox_b200 import ops, _lib

C, T, d = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else (32, 11968, 3)
B = 32
x = torch.randn(B, C, T, device="cuda")
w1 = torch.randn(C, C, 3, device="cuda") * 0.1
w2 = torch.randn(C, C, 1, device="cuda") * 0.1
pk = ops.residual_unit_pack(w1, w2)
for _ in range(3):
    ops.residual_unit_fwd(x, pk, d, 0.01)
buf = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
_lib.load().vbx_ru_set_profile_buffer(buf.data_ptr())
ops.residual_unit_fwd(x, pk, d, 0.01)
torch.cuda.synchronize()
_lib.load().vbx_ru_set_profile_buffer(None)
t = buf.cpu().view(64, 16)
names = ["tma_issue", "conv_wait", "conv_start", "conv_done", "mma1_issued", "mid_start", "mid_done", "mma2_issued",
         "epi_start", "epi_done"]
t0 = int(t[0, 0])
print(f"C={C} T={T} d={d}: cycles relative to the first TMA issue of CTA 0")
print("tile " + " ".join(f"{n:>11s}" for n in names))
for i in range(64):
    if int(t[i, 9]) == 0:
        break
    print(f"{i:4d} " + " ".join(f"{int(t[i, e]) - t0:11d}" for e in range(10)))
