"""Times the TMA weight-gradient kernel of the residual-unit convs against the gather-form kernel it replaces."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vibravox_b200 import ops
B = 32
for C, T, d, K in [(32, 11968, 9, 3), (32, 11968, 1, 1), (64, 5984, 9, 3), (64, 5984, 1, 1)]:
    x, dy = torch.randn(B, C, T, device="cuda"), torch.randn(B, C, T, device="cuda")
    pad = d * (K - 1) // 2
    g = ops.ConvGeom(C, C, K, 1, d, pad, pad, 1)
    def t(fn, reps=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    a = t(lambda: ops.unit_wgrad(x, dy, K, d))
    b = t(lambda: ops.tc_conv1d_wgrad(x, dy, g))
    byt = 4.0 * 2 * B * C * T
    print(f"C={C} T={T} d={d} K={K}: tma {a:6.1f} us ({byt / a / 1e3:5.0f} GB/s)   gather {b:6.1f} us", flush=True)
