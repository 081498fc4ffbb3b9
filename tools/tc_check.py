"""GPU check of the tcgen05 conv kernels against the fp32 SIMT kernels / fp64 torch (run under `timeout`)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from vibravox_b200 import ops

dev = "cuda"
torch.manual_seed(0)
CASES = [
    # B, Cin, Cout, Tin, K, s, d, pad, refl, groups
    (1, 32, 32, 256, 1, 1, 1, 0, 0, 1),
    (2, 32, 32, 300, 3, 1, 3, 3, 3, 1),
    (2, 64, 256, 403, 41, 4, 1, 20, 0, 4),
    (2, 1024, 1024, 60, 5, 1, 1, 2, 0, 1),
    (3, 768, 768, 50, 5, 1, 2, 2, 0, 4),
    (2, 32, 64, 301, 4, 2, 1, 1, 1, 1),
    (2, 24, 48, 131, 7, 2, 2, 3, 0, 4),
    (32, 1024, 1024, 748, 41, 4, 1, 20, 0, 4),
    (32, 32, 32, 11968, 3, 1, 3, 3, 3, 1),
    (2, 16, 32, 257, 16, 8, 1, 7, 7, 1),
    (2, 24, 48, 131, 7, 2, 3, 3, 0, 4),
    (2, 128, 256, 64, 16, 8, 1, 4, 0, 1),
    (32, 256, 1024, 2990, 41, 4, 1, 20, 0, 4),
    (32, 1024, 1024, 187, 5, 1, 1, 2, 0, 1),
]
which = [int(a) for a in sys.argv[1:]] or range(len(CASES))
for i in which:
    B, Cin, Cout, Tin, K, s, d, pad, refl, groups = CASES[i]
    g = ops.ConvGeom(Cin, Cout, K, s, d, pad, refl, groups)
    x = torch.randn(B, Cin, Tin, device=dev)
    w = torch.randn(Cout, Cin // groups, K, device=dev) / (Cin // groups * K) ** 0.5
    bias = torch.randn(Cout, device=dev)
    ref = ops.conv1d_fwd(x, w, g, bias=bias, slope=0.2)
    packed = ops.tc_pack(w, g, 0)
    y = ops.tc_conv1d_fwd(x, packed, g, bias=bias, slope=0.2)
    torch.cuda.synchronize()
    err = float((y - ref).abs().max()), float((y - ref).norm() / ref.norm())
    def tm(fn, n=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    t_tc = tm(lambda: ops.tc_conv1d_fwd(x, packed, g, bias=bias, slope=0.2))
    t_simt = tm(lambda: ops.conv1d_fwd(x, w, g, bias=bias, slope=0.2))
    flops = 2.0 * B * g.tout(Tin) * Cout * (Cin // groups) * K
    print(f"case {i} {CASES[i]}: max abs {err[0]:.2e} rel-L2 {err[1]:.2e} | tc {t_tc:.3f} ms ({flops/t_tc/1e9:.1f} TF) simt {t_simt:.3f} ms ({flops/t_simt/1e9:.1f} TF)", flush=True)
    # dgrad
    dyv = torch.randn_like(ref)
    wt = ops.transpose_weight(w, groups)
    r2 = torch.randn(B, Cin, Tin, device=dev)
    dref = ops.conv1d_dgrad(dyv, wt, g, Tin, res=r2, slope=0.5)
    pk = ops.tc_pack(w, g, 1)
    dgot = ops.tc_conv1d_dgrad(dyv, pk, g, Tin, res=r2, slope=0.5)
    torch.cuda.synchronize()
    derr = float((dgot - dref).abs().max()), float((dgot - dref).norm() / dref.norm())
    t_dtc = tm(lambda: ops.tc_conv1d_dgrad(dyv, pk, g, Tin))
    t_dsimt = tm(lambda: ops.conv1d_dgrad(dyv, wt, g, Tin))
    print(f"   dgrad: max abs {derr[0]:.2e} rel-L2 {derr[1]:.2e} | tc {t_dtc:.3f} ms ({flops/t_dtc/1e9:.1f} TF) simt {t_dsimt:.3f} ms ({flops/t_dsimt/1e9:.1f} TF)")
    # wgrad
    wref = ops.conv1d_wgrad(x, dyv, g)
    wgot = ops.tc_conv1d_wgrad(x, dyv, g)
    torch.cuda.synchronize()
    werr = float((wgot - wref).abs().max() / wref.abs().max()), float((wgot - wref).norm() / wref.norm())
    t_wtc = tm(lambda: ops.tc_conv1d_wgrad(x, dyv, g, dw=wgot))
    t_wsimt = tm(lambda: ops.conv1d_wgrad(x, dyv, g, dw=wref))
    print(f"   wgrad: max rel {werr[0]:.2e} rel-L2 {werr[1]:.2e} | tc {t_wtc:.3f} ms ({flops/t_wtc/1e9:.1f} TF) simt {t_wsimt:.3f} ms ({flops/t_wsimt/1e9:.1f} TF)")
print("done")
