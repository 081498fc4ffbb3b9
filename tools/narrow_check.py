"""Times the narrow (HBM-side) conv layers of the step at the bench shapes on the tensor-core path, forward and
input gradient, and checks them against the fp32 FMA kernels.  Run once per setting of the planning knobs
(VBX_TC_DENSE_MAX_CIN, VBX_TC_PS_MAX_KB, VBX_TC_SLAB_MIN_K) to compare kernel forms."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vibravox_b200 import ops

LAYERS = [  # Cin, Cout, K, s, d, pad, refl, groups, Tin, res
    (32, 32, 3, 1, 3, 3, 3, 1, 11968, False), (32, 32, 1, 1, 1, 0, 0, 1, 11968, True),
    (64, 64, 3, 1, 9, 9, 9, 1, 5984, False), (64, 64, 1, 1, 1, 0, 0, 1, 5984, True),
    (128, 128, 3, 1, 1, 1, 1, 1, 1496, False), (128, 128, 1, 1, 1, 0, 0, 1, 1496, True),
    (32, 64, 4, 2, 1, 1, 1, 1, 11968, False), (32, 4, 3, 1, 1, 1, 1, 1, 11968, False),
    (16, 64, 41, 4, 1, 20, 0, 4, 47840, False),
    (24, 48, 7, 2, 1, 3, 0, 4, 11970, False), (24, 48, 7, 2, 3, 3, 0, 4, 11966, False),
    (48, 96, 7, 2, 2, 3, 0, 4, 5981, False),
    (96, 192, 7, 2, 1, 3, 0, 4, 2993, False), (96, 192, 7, 2, 3, 3, 0, 4, 2983, False),
    (192, 384, 7, 2, 2, 3, 0, 4, 1491, False),
    (64, 256, 41, 4, 1, 20, 0, 4, 11960, False),
]
B = 32
ONLY = [int(v) for v in sys.argv[1].split(',')] if len(sys.argv) > 1 else None     # layer indices (for ncu runs)
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 20


def timeit(fn, reps=REPS):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print({k: v for k, v in os.environ.items() if k.startswith("VBX_")})
tot_f = tot_d = 0.0
for Cin, Cout, K, s, d, pad, refl, groups, Tin, with_res in (LAYERS if ONLY is None else [LAYERS[i] for i in ONLY]):
    g = ops.ConvGeom(Cin, Cout, K, s, d, pad, refl, groups)
    To = g.tout(Tin)
    torch.manual_seed(0)
    x = torch.randn(B, Cin, Tin, device="cuda")
    dy = torch.randn(B, Cout, To, device="cuda")
    w = torch.randn(Cout, Cin // groups, K, device="cuda") / (Cin // groups * K) ** 0.5
    bias = torch.randn(Cout, device="cuda")
    res = torch.randn(B, Cout, To, device="cuda") if with_res else None
    pf, pd = ops.tc_pack(w, g, ops.TC_FWD), ops.tc_pack(w, g, ops.TC_DGRAD)
    kw = dict(bias=bias, res=res, slope=0.2, want_mask=with_res)
    y = ops.tc_conv1d_fwd(x, pf, g, **kw)
    y0 = ops.conv1d_fwd(x, w, g, **kw)
    if with_res:
        assert (y[1] != y0[1]).float().mean() < 1e-3
        y, y0 = y[0], y0[0]
    dx = ops.tc_conv1d_dgrad(dy, pd, g, Tin)
    dx0 = ops.conv1d_dgrad(dy, ops.transpose_weight(w, groups), g, Tin)
    ef = float((y - y0).norm() / y0.norm()); ed = float((dx - dx0).norm() / dx0.norm())
    tf = timeit(lambda: ops.tc_conv1d_fwd(x, pf, g, **kw))
    td = timeit(lambda: ops.tc_conv1d_dgrad(dy, pd, g, Tin))
    byt = 4.0 * (x.numel() + dy.numel())
    bf = byt + (5.0 * dy.numel() if with_res else 0)
    tot_f += tf; tot_d += td
    flag = "" if ef < 2e-5 and ed < 2e-5 else "  <-- MISMATCH"
    print(f"{Cin}>{Cout} k{K} s{s} d{d} g{groups} r{refl} T{Tin}: fwd {tf*1e3:7.1f} us ({bf/tf/1e6:6.0f} GB/s, err {ef:.1e})  "
          f"dgrad {td*1e3:7.1f} us ({byt/td/1e6:6.0f} GB/s, err {ed:.1e}){flag}", flush=True)
print(f"sum fwd {tot_f*1e3:.0f} us, dgrad {tot_d*1e3:.0f} us")
