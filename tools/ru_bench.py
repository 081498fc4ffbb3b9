"""Times the fused ResidualUnit kernel at the bench shapes (rotating over more than L2 worth of inputs).
usage: python tools/ru_bench.py [C T d [reps]]   (VBX_RU_NR / VBX_RU_NA select the ring depths)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vibravox_b200 import ops

B = 32
shapes = [(32, 11968, 3), (32, 11968, 9), (64, 5984, 3), (64, 5984, 9)]
if len(sys.argv) > 3:
    shapes = [(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 30
train = os.environ.get("RU_TRAIN", "0") == "1"
for C, T, d in shapes:
    nbuf = max(2, int(300e6 // (4 * B * C * T)) + 1)
    xs = [torch.randn(B, C, T, device="cuda") for _ in range(nbuf)]
    w1 = torch.randn(C, C, 3, device="cuda") * 0.1
    w2 = torch.randn(C, C, 1, device="cuda") * 0.1
    pk = ops.residual_unit_pack(w1, w2)
    for i in range(3):
        ops.residual_unit_fwd(xs[i % nbuf], pk, d, 0.01, want_h=train, want_mask=train)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        ops.residual_unit_fwd(xs[i % nbuf], pk, d, 0.01, want_h=train, want_mask=train)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    byt = 4.0 * 2 * B * C * T * (1.625 if train else 1.0)
    print(f"C={C} T={T} d={d} train={int(train)} NR={os.environ.get('VBX_RU_NR','-')} NA={os.environ.get('VBX_RU_NA','-')}: "
          f"{us:7.1f} us  {byt / us / 1e3:6.0f} GB/s", flush=True)
    del xs
