"""Per-layer roofline table of one training step: records every conv-family call (op, geometry, batch, length,
epilogue) of an eager step at the bench configuration, times each distinct call alone with CUDA events, and ranks
by count x time.  Columns: algorithmic HBM bytes (operands read once, result written once), flops, achieved GB/s and
TF/s, and the time the measured rooflines allow: max(bytes / HBM copy peak, 3 x flops / bf16 peak) - the kernels
issue three bf16 MMAs per fp32 product - so `x` = measured / allowed."""
import os, sys, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vibravox_b200
from vibravox_b200 import ops
from vibravox_b200 import data as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(float(sys.argv[2]) * 16000) if len(sys.argv) > 2 else 48000
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
def _find(d, *names):
    for k, v in d.items():
        if isinstance(v, dict):
            r = _find(v, *names)
            if r: return r
        elif any(n in k.lower() for n in names) and isinstance(v, (int, float)):
            return float(v)
    return None
HBM = (_find(peaks, "hbm") or 6442.9) * 1e9
TF = (_find(peaks, "burst") or 1620.8) * 1e12

calls = collections.OrderedDict()
def rec(kind, g, Bx, Tin, **kw):
    key = (kind, g, Bx, Tin, tuple(sorted(kw.items())))
    calls[key] = calls.get(key, 0) + 1
orig = {n: getattr(ops, n) for n in ("conv1d_fwd", "conv1d_dgrad", "conv1d_wgrad", "conv1d_dgrad_scatter",
                                      "tc_conv1d_fwd", "tc_conv1d_dgrad", "tc_conv1d_wgrad")}
def h_fwd(name):
    def f(x, w, g, bias=None, res=None, slope=1.0, want_mask=False, **k):
        rec(name, g, x.shape[0], x.shape[2], bias=bias is not None, res=res is not None, mask=want_mask,
            nsplit=k.get("nsplit", 2))
        return orig[name](x, w, g, bias=bias, res=res, slope=slope, want_mask=want_mask, **k)
    return f
def h_dgrad(name):
    def f(dy, w, g, Tin, res=None, slope=1.0, **k):
        rec(name, g, dy.shape[0], Tin, res=res is not None, nsplit=k.get("nsplit", 2))
        return orig[name](dy, w, g, Tin, res=res, slope=slope, **k)
    return f
def h_wgrad(name):
    def f(x, dy, g, dw=None):
        rec(name, g, x.shape[0], x.shape[2])
        return orig[name](x, dy, g, dw=dw)
    return f
def h_scatter(dy, wk, g, Tin, dx=None):
    rec("conv1d_dgrad_scatter", g, dy.shape[0], Tin)
    return orig["conv1d_dgrad_scatter"](dy, wk, g, Tin, dx)
ops.conv1d_fwd, ops.tc_conv1d_fwd = h_fwd("conv1d_fwd"), h_fwd("tc_conv1d_fwd")
ops.conv1d_dgrad, ops.tc_conv1d_dgrad = h_dgrad("conv1d_dgrad"), h_dgrad("tc_conv1d_dgrad")
ops.conv1d_wgrad, ops.tc_conv1d_wgrad = h_wgrad("conv1d_wgrad"), h_wgrad("tc_conv1d_wgrad")
ops.conv1d_dgrad_scatter = h_scatter

body, air = O.synthetic_pairs(B, S, seed=1)
lm = vibravox_b200.build_model(seed=42, device="cuda")
batch = {"audio_body_conducted": body.cuda(), "audio_airborne": air.cuda()}
lm.training_step(batch)
calls.clear()
lm.training_step(batch)
torch.cuda.synchronize()
for n, f in orig.items():
    setattr(ops, n, f)
del lm
torch.cuda.empty_cache()

def timeit(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3

rows = []
for (kind, g, Bx, Tin, kw), cnt in calls.items():
    kw = dict(kw)
    Tout = g.tout(Tin)
    cin_g = g.Cin // g.groups
    x = torch.randn(Bx, g.Cin, Tin, device="cuda")
    dy = torch.randn(Bx, g.Cout, Tout, device="cuda")
    w = torch.randn(g.Cout, cin_g, g.K, device="cuda") * 0.05
    flops = 2.0 * Bx * Tout * g.Cout * cin_g * g.K
    nx, ny = 4.0 * x.numel(), 4.0 * dy.numel()
    tc = kind.startswith("tc_")
    if kind.endswith("fwd"):
        bias = torch.randn(g.Cout, device="cuda") if kw["bias"] else None
        res = torch.randn_like(dy) if kw["res"] else None
        byt = nx + ny + (ny if kw["res"] else 0) + (ny / 4 if kw["mask"] else 0)
        if tc:
            pk = ops.tc_pack(w, g, ops.TC_FWD, kw["nsplit"])
            fn = lambda: ops.tc_conv1d_fwd(x, pk, g, bias=bias, res=res, slope=0.2, want_mask=kw["mask"], nsplit=kw["nsplit"])
        else:
            fn = lambda: ops.conv1d_fwd(x, w, g, bias=bias, res=res, slope=0.2, want_mask=kw["mask"])
    elif kind.endswith("dgrad"):
        res = torch.randn_like(x) if kw["res"] else None
        byt = nx + ny + (nx if kw["res"] else 0)
        if tc:
            pk = ops.tc_pack(w, g, ops.TC_DGRAD, kw["nsplit"])
            fn = lambda: ops.tc_conv1d_dgrad(dy, pk, g, Tin, res=res, nsplit=kw["nsplit"])
        else:
            wt = ops.transpose_weight(w, g.groups)
            fn = lambda: ops.conv1d_dgrad(dy, wt, g, Tin, res=res)
    elif kind.endswith("wgrad"):
        byt = nx + ny
        dw = torch.zeros_like(w)
        fn = (lambda: ops.tc_conv1d_wgrad(x, dy, g, dw=dw)) if tc else (lambda: ops.conv1d_wgrad(x, dy, g, dw=dw))
    else:
        continue
    nsp = 3 if kw.get("nsplit", 2) == 2 else 6
    t = timeit(fn)
    allowed = max(byt / HBM, (nsp if tc else 1) * flops / TF)
    rows.append(dict(kind=kind.replace("conv1d_", ""), geom=f"{g.Cin}>{g.Cout} k{g.K} s{g.stride} d{g.dil} g{g.groups}",
                     B=Bx, Tin=Tin, n=cnt, ms=t * 1e3, tot=t * 1e3 * cnt, gbs=byt / t / 1e9, tfs=flops / t / 1e12,
                     bound="hbm" if byt / HBM > (nsp if tc else 1) * flops / TF else "mma", x=t / allowed,
                     extra=",".join(k for k in ("bias", "res", "mask") if kw.get(k))))
    del x, dy, w
rows.sort(key=lambda r: -r["tot"])
tot = sum(r["tot"] for r in rows)
print(f"bs={B} x {S / 16000:g} s; {len(rows)} distinct calls, {sum(r['n'] for r in rows)} launches, {tot:.2f} ms summed (isolated, warm)")
print("| op | geometry | B x Tin | n | ms each | ms total | GB/s | TF/s | bound | x over roofline | excess ms |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    print(f"| {r['kind']} {r['extra']} | {r['geom']} | {r['B']}x{r['Tin']} | {r['n']} | {r['ms']:.3f} | {r['tot']:.2f} | "
          f"{r['gbs']:.0f} | {r['tfs']:.1f} | {r['bound']} | {r['x']:.1f} | {r['tot'] * (1 - 1 / r['x']):.2f} |")
