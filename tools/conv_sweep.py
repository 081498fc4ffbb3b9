"""BASELINE.json configs[4]: dilated/strided Conv1d micro-bench sweep, this repo's kernels vs cuDNN
(torch.nn.functional.conv1d, both fp32 'ieee' and PyTorch's default TF32), B=32, reflect padding.
Writes a markdown table (stdout) with ms, algorithmic GB/s, TFLOP/s and the max-abs deviation of each
implementation from an fp64 reference on a slice of the output."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from vibravox_b200 import ops

dev = "cuda"
B = 32
shapes = [(3, 1, 1), (3, 3, 1), (3, 9, 1), (1, 1, 1), (7, 1, 1), (4, 1, 2), (8, 1, 4), (16, 1, 8)]   # (k, d, s)
Cs = [int(c) for c in os.environ.get("SWEEP_C", "32,64,128,256,512").split(",")]
Ls = [int(l) for l in os.environ.get("SWEEP_L", "4096,16384,65536").split(",")]
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}


def tm(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print(f"| C | L | k,d,s | ours fwd ms | GB/s (% of {peaks['hbm_gbs']:.0f}) | TF/s | cuDNN ieee ms | cuDNN tf32 ms | ours dgrad ms | ours wgrad ms | err ours | err tf32 |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for C in Cs:
    for L in Ls:
        if B * C * L * 4 > 3e9:
            continue
        for (k, d, s) in shapes:
            pad = d * (k - 1) // 2 if s == 1 else s - 1
            g = ops.ConvGeom(C, C, k, s, d, pad, pad, 1)
            x = torch.randn(B, C, L, device=dev)
            w = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
            xp = F.pad(x, (pad, pad), mode="reflect") if pad else x
            t_ours = tm(lambda: ops.conv_fwd(x, w, g))
            y = ops.conv_fwd(x, w, g)
            torch.backends.cudnn.conv.fp32_precision = "ieee"
            t_ieee = tm(lambda: F.conv1d(F.pad(x, (pad, pad), mode="reflect") if pad else x, w, None, s, 0, d))
            torch.backends.cudnn.conv.fp32_precision = "tf32"
            t_tf32 = tm(lambda: F.conv1d(F.pad(x, (pad, pad), mode="reflect") if pad else x, w, None, s, 0, d))
            y_tf32 = F.conv1d(xp, w, None, s, 0, d)
            ref = F.conv1d(xp[:2].double(), w.double(), None, s, 0, d)
            e_ours = float((y[:2].double() - ref).abs().max() / ref.abs().max())
            e_tf32 = float((y_tf32[:2].double() - ref).abs().max() / ref.abs().max())
            dy = torch.randn_like(y)
            wt = ops.transpose_weight(w, 1)
            t_dg = tm(lambda: ops.conv_dgrad(dy, w, wt, g, L))
            dw = torch.zeros_like(w)
            t_wg = tm(lambda: ops.conv_wgrad(x, dy, g, dw=dw))
            To = y.shape[2]
            byts = 4.0 * B * (C * L + C * To) + 4.0 * C * C * k
            fl = 2.0 * B * To * C * C * k
            print(f"| {C} | {L} | {k},{d},{s} | {t_ours:.3f} | {byts / t_ours / 1e6:.0f} ({100 * byts / t_ours / 1e6 / peaks['hbm_gbs']:.0f}%) | "
                  f"{fl / t_ours / 1e9:.1f} | {t_ieee:.3f} | {t_tf32:.3f} | {t_dg:.3f} | {t_wg:.3f} | {e_ours:.1e} | {e_tf32:.1e} |", flush=True)
            del x, w, y, dy, xp, y_tf32
