"""autograd.Function wrappers: each forward/backward is one or a few libvbx_b200 kernels.

These give the drop-in nn.Modules (torch_modules/) ordinary PyTorch autograd semantics
(`loss.backward()`, `torch.autograd.grad(loss, generator.last_conv.weight)` as used by
lightning_modules/eben.py:225-228 of the reference) while every FLOP runs in the
hand-written sm_100a kernels.  First-order only (once_differentiable).
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops
from .ops import ConvGeom

Tensor = torch.Tensor


UNIT_COMPOSED = __import__("os").environ.get("VBX_UNIT_COMPOSED", "1") != "0"
# unit input gradient as the zero-halo gradient on the slab-form kernels + vbx_reflect_fold_k3 (mirror terms).  Off by
# default: parity-green, but measured 36.5 vs 35.7 ms per step - at C <= 128 the gather-form reflect kernel is the faster one
UNIT_ZERO_HALO_DGRAD = __import__("os").environ.get("VBX_UNIT_ZERO_HALO_DGRAD", "0") == "1"
# bias gradient of a chain stage reduced inside the next stage's gate pass (vbx_epilogue.gate_dbias).  Off by default:
# measured 36.6 vs 35.9 ms per step - the per-channel reduction lengthens a pass that the rest of the chain waits for,
# while the separate read-only reduction runs beside the chain
GATE_DBIAS = __import__("os").environ.get("VBX_GATE_DBIAS", "0") == "1"


def _c(t: Tensor) -> Tensor:
    return t if t.is_contiguous() else t.contiguous()


class Flags:
    """Switches consulted by the backward kernels of the shared-schedule training step
    (lightning_modules/eben.py): one discriminator graph serves both phases, so the generator phase
    walks it with parameter gradients OFF and the discriminator phase walks it with the gradient
    w.r.t. graph-leaf inputs (the detached generator outputs) OFF."""
    param_grads = True
    skip_leaf_input_grad = False
    # Fused backward of a discriminator chain (ConvFn -> ConvFn -> ..., every stage output also read by the
    # feature-matching loss).  While `gated_chain` is set, the discriminator forwards tell each stage the slope of the stage
    # that produced its input; in the backward pass the stage's input-gradient epilogue then multiplies by LeakyReLU'(x)
    # and adds x's feature-matching gradient itself, instead of leaving an aten::add (gradient accumulation), an L1-pair
    # backward pass and a LeakyReLU backward pass to run between the two convs.  The hand-over between backward nodes goes
    # through the two tables below, keyed by the activation's storage: `pending` = feature-matching terms registered by
    # FeatureMatchingFn.backward and not yet applied, `gated` = activations whose incoming gradient already carries their
    # LeakyReLU'.  CONTRACT: a tensor produced under `gated_chain` may only be consumed by the next stage of its chain and
    # by FeatureMatchingFn (vibravox_b200.lightning_modules.eben sets the flag around its discriminator calls only).
    gated_chain = False
    pending: dict = {}
    gated: dict = {}          # activation key -> True when its stage's bias gradient was reduced along with the gate

    @classmethod
    def chain_reset(cls) -> None:
        cls.pending.clear()
        cls.gated.clear()

    @classmethod
    def chain_check(cls) -> None:
        """After a backward pass: every registered feature-matching term must have been applied by some stage."""
        left = len(cls.pending)
        cls.chain_reset()
        if left:
            raise RuntimeError(f"{left} feature-matching gradient terms were registered for fused application but no "
                               "ConvFn stage consumed them")


def _key(t: Tensor):
    return (t.data_ptr(), t.numel())


def grad_slot(p: Tensor) -> Optional[Tensor]:
    """View into a flat gradient bucket (set by vibravox_b200.optim.FlatAdam on its parameters).
    When present, backward kernels accumulate the parameter gradient there directly and autograd
    gets None for that input: no per-parameter grad tensors, no AccumulateGrad adds, one Adam
    launch and one all-reduce per network."""
    return getattr(p, "_vbx_grad", None)


class WeightNormFn(Function):
    """w = g * v / ||v||_(dims 1,2)  (torch_modules/utils.py:4-9, weight_norm dim=0).
    Returns (w, wt): the conv layout and the group-transposed layout the dgrad kernel reads."""

    @staticmethod
    def forward(ctx, g: Tensor, v: Tensor, groups: int):
        w, wt, inv = ops.weight_norm_fwd(_c(g), _c(v), groups, want_wt=True)
        ctx.save_for_backward(g, v, inv)
        ctx.slots = (grad_slot(g), grad_slot(v))
        ctx.mark_non_differentiable(wt)
        return w, wt

    @staticmethod
    @once_differentiable
    def backward(ctx, dw, _dwt):
        g, v, inv = ctx.saved_tensors
        sg, sv = ctx.slots
        if sg is not None and sv is not None:      # accumulate straight into the flat gradient bucket
            ops.weight_norm_bwd(_c(g), _c(v), inv, _c(dw), dg=sg, dv=sv, beta=1.0)
            return None, None, None
        dg, dv = ops.weight_norm_bwd(_c(g), _c(v), inv, _c(dw))
        return dg.view_as(g), dv, None


class TransposeWeightFn(Function):
    """wt for an un-normalised conv weight (first_conv / last_conv); no gradient through wt."""

    @staticmethod
    def forward(ctx, w: Tensor, groups: int):
        wt = ops.transpose_weight(_c(w), groups)
        ctx.mark_non_differentiable(wt)
        return wt

    @staticmethod
    def backward(ctx, _):
        return None, None


class ConvFn(Function):
    """y = LeakyReLU_slope(conv1d(x, w) + bias), halo (reflect and/or zero) folded into the kernel.
    `in_slope` != 1 (set by the discriminator forwards under Flags.gated_chain): x is the LeakyReLU_in_slope output of the
    previous ConvFn of a chain, and this stage's input-gradient kernel finishes that stage's backward (see Flags)."""

    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, wt: Optional[Tensor], bias: Optional[Tensor], geom: ConvGeom,
                slope: float, in_slope: float = 1.0, in_bias: Optional[Tensor] = None):
        x_leaf = x.is_leaf                      # before any copy: is x a graph leaf (e.g. a detached G output)?
        x, w = _c(x), _c(w)
        y = ops.conv_fwd(x, w, geom, bias=bias, slope=slope)
        ctx.geom, ctx.slope, ctx.has_bias = geom, slope, bias is not None
        ctx.in_slope = in_slope
        # bias of the stage that produced x (chain contract): its gradient is reduced where this stage gates dx, when it
        # lives in a flat gradient bucket
        ctx.in_b_slot = grad_slot(in_bias) if (GATE_DBIAS and in_bias is not None and in_bias.requires_grad) else None
        ctx.w_slot = grad_slot(w)
        ctx.b_slot = grad_slot(bias) if bias is not None else None
        ctx.x_leaf = x_leaf
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x, w, wt, y if slope != 1.0 else None)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, w, wt, y = ctx.saved_tensors
        geom, slope = ctx.geom, ctx.slope
        # hand-over from the stages downstream (Flags): is gy already multiplied by this stage's LeakyReLU', and is there a
        # feature-matching term of y that nobody applied (its consumer's input gradient is not part of this pass)?
        pre_gated, pend = False, None
        if y is not None and (Flags.gated or Flags.pending):
            ky = _key(y)
            pre_gated = ky in Flags.gated
            bias_done = Flags.gated.pop(ky, False)
            pend = Flags.pending.pop(ky, None)
        else:
            bias_done = False
        if gy is None and pend is None:
            return (None,) * 8
        gy = _c(gy) if gy is not None else None
        need_x, need_w, _, need_b = ctx.needs_input_grad[:4]
        need_w, need_b = need_w and Flags.param_grads, need_b and Flags.param_grads
        need_x = need_x and not (Flags.skip_leaf_input_grad and ctx.x_leaf)
        dbias = None
        if ctx.has_bias and need_b and not bias_done:
            dbias = ctx.b_slot if ctx.b_slot is not None else \
                torch.zeros((geom.Cout,), device=x.device, dtype=torch.float32)
        if pend is not None:
            other, coef, ev = pend
            ops.wait_event(ev)
            # (a gy that arrives here came from consumers outside the chain contract and is not gated yet)
            assert not pre_gated
            gp = ops.fm_gate_bwd(y, other, coef, slope, gy)
            if dbias is not None:
                ops.leaky_relu_bwd(gp, None, 1.0, dbias=dbias, want_dx=False)
        elif pre_gated:
            gp = gy
            if dbias is not None:
                ops.leaky_relu_bwd(gp, None, 1.0, dbias=dbias, want_dx=False)
        else:
            gp = gy
            if slope != 1.0 or dbias is not None:
                out = ops.leaky_relu_bwd(gy, y, slope, dbias=dbias, want_dx=slope != 1.0)
                gp = out if out is not None else gy
        dx = dw = None
        if need_x:
            gate = None
            if ctx.in_slope != 1.0:
                kx = _key(x)
                term = Flags.pending.pop(kx, None)
                in_db = ctx.in_b_slot if Flags.param_grads else None
                if term is not None:
                    ops.wait_event(term[2])
                    gate = (x, ctx.in_slope, term[0], term[1], in_db)
                else:
                    gate = (x, ctx.in_slope, None, None, in_db)
                Flags.gated[kx] = in_db is not None
            dx = ops.conv_dgrad(gp, w, wt, geom, x.shape[2], gate=gate)
        if need_w:
            dw = ops.conv_wgrad(x, gp, geom, dw=ctx.w_slot)
        return (dx, None if ctx.w_slot is not None else dw, None,
                None if ctx.b_slot is not None else dbias, None, None, None, None)


class ConvTransposeFn(Function):
    """y = LeakyReLU_slope(conv_transpose1d(x + skip, w)) (DecBlock.forward head,
    eben_generator.py:251-253).  `geom` is the geometry of the *equivalent forward conv*
    (Cout = ConvT in_channels, Cin = ConvT out_channels)."""

    @staticmethod
    def forward(ctx, x: Tensor, skip: Optional[Tensor], w: Tensor, wt: Tensor, geom: ConvGeom, slope: float,
                output_padding: int):
        xs = ops.add(_c(x), _c(skip)) if skip is not None else _c(x)
        T = xs.shape[2]
        L = (T - 1) * geom.stride - 2 * geom.pad + geom.dil * (geom.K - 1) + output_padding + 1
        y = ops.conv_dgrad(xs, w, wt, geom, L, slope=slope)
        ctx.geom, ctx.slope, ctx.has_skip = geom, slope, skip is not None
        ctx.save_for_backward(xs, _c(w), y if slope != 1.0 else None)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        xs, w, y = ctx.saved_tensors
        geom, slope = ctx.geom, ctx.slope
        gp = _c(gy)
        if slope != 1.0:
            gp = ops.leaky_relu_bwd(gp, y, slope)
        dxs = dw = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            dxs = ops.conv_fwd(gp, w, geom)
        if ctx.needs_input_grad[2]:
            dw = ops.conv_wgrad(gp, xs, geom)
        return (dxs if ctx.needs_input_grad[0] else None,
                dxs if (ctx.has_skip and ctx.needs_input_grad[1]) else None, dw, None, None, None, None)


class ResidualUnitFn(Function):
    """out = x + LeakyReLU(pointwise(dilated(x)))   (ResidualUnit.forward, eben_generator.py:314-316).
    Forward: one fused kernel where the shape allows (ops.residual_unit_fwd), else two conv kernels whose second epilogue
    does the activation and the residual add; either way the 1-bit / 1-byte activation mask is kept for the backward (it is
    not recoverable from `out`).
    Backward, through the COMPOSED conv (include/vbx.h: vbx_unit_combine): the two bias-free convs are one k-tap conv with
    wf = w2 . w1, so  dx = dgrad(dz, wf) + g  and  dwf = wgrad(x, dz)  are all the heavy work - one input gradient and one
    weight gradient per unit instead of two each - and dw1 = w2^T dwf, dw2 = <dwf, w1> are C x C x 3C products.  The
    intermediate activation h is neither saved nor re-read.  VBX_UNIT_COMPOSED=0 restores the layer-by-layer backward."""

    @staticmethod
    def forward(ctx, x: Tensor, w1: Tensor, wt1: Tensor, w2: Tensor, wt2: Tensor, g1: ConvGeom, g2: ConvGeom,
                slope: float):
        x = _c(x)
        B, C, T = x.shape
        unit = (g1.K == 3 and g1.stride == 1 and g1.groups == 1 and g1.refl == g1.pad == g1.dil and g2.K == 1
                and g1.Cin == g1.Cout == g2.Cin == g2.Cout == C)
        fused = unit and ops.use_fused_unit(C, T, g1.dil, B)
        train = any(ctx.needs_input_grad)
        composed = unit and UNIT_COMPOSED
        wf = None
        if fused:
            # one kernel: x read once, out written once; h / mask only when a backward pass will need them
            pk = ops.residual_unit_pack(_c(w1), _c(w2))
            out, h, mask = ops.residual_unit_fwd(x, pk, g1.dil, slope, want_h=train and not composed, want_mask=train)
        elif composed and ops.TC_ENABLED:
            # no fused kernel for this width: the composed conv is still ONE launch (k-tap conv with wf, LeakyReLU + residual
            # + mask in its epilogue) instead of two with h written and re-read in between.  (Not on the fp32 FMA path, which
            # keeps the reference's operation order in the forward pass: composing the weights moves the output by ~1e-7,
            # which the loss balancing amplifies past that path's tighter trajectory bound.)
            h = None
            wf = ops.unit_combine(_c(w1), _c(w2))
            out, mask = ops.conv_fwd(x, wf, g1, res=x, slope=slope, want_mask=True)
        else:
            h = ops.conv_fwd(x, w1, g1)
            out, mask = ops.conv_fwd(h, w2, g2, res=x, slope=slope, want_mask=True)
        ctx.g1, ctx.g2, ctx.slope, ctx.composed = g1, g2, slope, composed
        ctx.save_for_backward(x, h, mask, wt1, wt2, w1, w2, wf if train else None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, h, mask, wt1, wt2, w1, w2, wf = ctx.saved_tensors
        g1, g2, slope = ctx.g1, ctx.g2, ctx.slope
        g = _c(g)
        T = x.shape[2]
        B, C = x.shape[0], x.shape[1]
        dz = ops.leaky_relu_bwd(g, None, slope, mask=mask)
        unit = (g1.K == 3 and g1.stride == 1 and g1.groups == 1 and g1.refl == g1.pad == g1.dil and g2.K == 1
                and g1.Cin == g1.Cout == g2.Cin == g2.Cout == C)
        tma_wgrad = unit and ops.unit_wgrad_workspace(B, C, T, g1.dil, 3) > 0
        need_w1, need_w2 = ctx.needs_input_grad[1], ctx.needs_input_grad[3]
        if ctx.composed:
            w1c, w2c = _c(w1), _c(w2)
            if wf is None:
                wf = ops.unit_combine(w1c, w2c)
            dw1 = dw2 = None
            if need_w1 or need_w2:
                dwf = ops.unit_wgrad(x, dz, 3, g1.dil) if tma_wgrad else ops.conv_wgrad(x, dz, g1)
                dw1, dw2 = ops.unit_split_grads(dwf, w1c, w2c, need_w1, need_w2)
            dx = None
            if ctx.needs_input_grad[0]:
                if UNIT_ZERO_HALO_DGRAD and ops.TC_ENABLED and T > 2 * g1.dil:
                    # zero-halo input gradient on the slab-form kernels + the mirror terms of the 2*d edge positions
                    gz = ConvGeom(g1.Cin, g1.Cout, g1.K, g1.stride, g1.dil, g1.pad, 0, g1.groups)
                    dx = ops.conv_dgrad(dz, wf, None, gz, T, res=g)
                    ops.reflect_fold_k3(dz, wf, dx, g1.dil)
                else:
                    dx = ops.conv_dgrad(dz, wf, None, g1, T, res=g)
            return dx, dw1, None, dw2, None, None, None, None
        if need_w2:
            dw2 = ops.unit_wgrad(h, dz, 1, 1) if tma_wgrad else ops.conv_wgrad(h, dz, g2)
        else:
            dw2 = None
        dh = ops.conv_dgrad(dz, w2, wt2, g2, T)
        if need_w1:
            dw1 = ops.unit_wgrad(x, dh, 3, g1.dil) if tma_wgrad else ops.conv_wgrad(x, dh, g1)
        else:
            dw1 = None
        dx = ops.conv_dgrad(dh, w1, wt1, g1, T, res=g) if ctx.needs_input_grad[0] else None
        return dx, dw1, None, dw2, None, None, None, None


class LeakyReluFn(Function):
    @staticmethod
    def forward(ctx, x: Tensor, slope: float):
        x = _c(x)
        ctx.slope = slope
        ctx.save_for_backward(x)
        return ops.leaky_relu_fwd(x, slope)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        gy = _c(gy)
        return ops.leaky_relu_bwd(gy.view(x.shape), x, ctx.slope), None


class TanhRecomposeFn(Function):
    """tanh(x + cat(first_bands, zeros))   (eben_generator.py:203-208)."""

    @staticmethod
    def forward(ctx, x: Tensor, first: Tensor, p: int):
        y = ops.tanh_recompose_fwd(_c(x), _c(first), p)
        ctx.p = p
        ctx.save_for_backward(y)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        dx = ops.tanh_bwd(_c(gy), y)
        dfirst = dx[:, :ctx.p].contiguous() if ctx.needs_input_grad[1] else None
        return dx, dfirst, None


class PQMFAnalysisFn(Function):
    """F.conv1d(signal, W_a[:bands], stride=m, padding=n-1)   (pqmf.py:194-202)."""

    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, bands: int):
        x = _c(x)
        ctx.save_for_backward(w)
        ctx.L, ctx.bands = x.shape[2], bands
        return ops.pqmf_analysis(x, w, bands)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        (w,) = ctx.saved_tensors
        dx = ops.pqmf_synthesis(_c(gy), w, sum_bands=True, L=ctx.L) if ctx.needs_input_grad[0] else None
        return dx, None, None


class PQMFSynthesisFn(Function):
    """F.conv_transpose1d(bands, W_s, stride=m, groups=m, padding=n-1, output_padding=m-2), optionally
    fused with the generator's sum over bands   (pqmf.py:204-213, eben_generator.py:209-211)."""

    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, sum_bands: bool):
        x = _c(x)
        ctx.save_for_backward(w)
        ctx.T, ctx.sum_bands, ctx.bands = x.shape[2], sum_bands, x.shape[1]
        return ops.pqmf_synthesis(x, w, sum_bands)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        (w,) = ctx.saved_tensors
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.pqmf_analysis(_c(gy), w, ctx.bands, T=ctx.T, x_per_band=not ctx.sum_bands)
        return dx, None, None


class FeatureMatchingFn(Function):
    """sum_i mean|a_i - b_i| / mean|a_i|, times `scale`   (losses/feature_loss.py:37-50).
    `fused` (every a_i is a stage output of a discriminator chain run under Flags.gated_chain): the backward hands the
    gradient terms of the a_i to the conv stages (Flags.pending) instead of materialising them."""

    @staticmethod
    def forward(ctx, scale: float, n: int, fused: bool, *tensors: Tensor):
        a, b = [_c(t) for t in tensors[:n]], [_c(t) for t in tensors[n:]]
        sums = torch.zeros((2 * n,), device=a[0].device, dtype=torch.float64)
        for i in range(n):
            ops.l1_pair_sums(a[i], b[i], sums[2 * i:2 * i + 2])
        loss = ops.fm_finalize(sums, n, scale)
        ctx.scale, ctx.n, ctx.sums, ctx.fused = scale, n, sums, fused
        ctx.save_for_backward(*a, *b)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, go):
        n, sums = ctx.n, ctx.sums
        a, b = ctx.saved_tensors[:n], ctx.saved_tensors[n:]
        go = _c(go).float()
        grads: List[Optional[Tensor]] = [None] * (2 * n)
        need_a = [ctx.needs_input_grad[3 + i] for i in range(n)]
        need_b = [ctx.needs_input_grad[3 + n + i] and Flags.param_grads for i in range(n)]
        if ctx.fused:
            if any(need_b):
                raise NotImplementedError("fused feature-matching backward: gradients w.r.t. the second (reference) "
                                          "feature list are not part of the shared-schedule step")
            if any(need_a):
                coef = ops.fm_coef(sums, n, go, ctx.scale)
                ev = ops.record_event()
                for i in range(n):
                    if need_a[i]:
                        Flags.pending[_key(a[i])] = (b[i], coef[2 * i:2 * i + 2], ev)
            return (None, None, None, *grads)
        for i in range(n):
            if need_a[i] or need_b[i]:
                grads[i], grads[n + i] = ops.l1_pair_bwd(a[i], b[i], sums[2 * i:2 * i + 2], go, ctx.scale, need_a[i],
                                                         need_b[i])
        return (None, None, None, *grads)


class HingeFn(Function):
    """mean over scales of mean(relu(1 - target*certainty))   (losses/hinge_loss.py:35-43)."""

    @staticmethod
    def forward(ctx, target: float, *certs: Tensor):
        certs = [_c(c) for c in certs]
        acc = torch.zeros((1,), device=certs[0].device, dtype=torch.float64)
        for c in certs:
            ops.hinge_fwd(c, target, 1.0 / (c.numel() * len(certs)), acc)
        ctx.target = target
        ctx.save_for_backward(*certs)
        return ops.d2f(acc).view(())

    @staticmethod
    @once_differentiable
    def backward(ctx, go):
        certs = ctx.saved_tensors
        go = _c(go).float()
        grads = [ops.hinge_bwd(c, ctx.target, 1.0 / (c.numel() * len(certs)), go) if need else None
                 for c, need in zip(certs, ctx.needs_input_grad[1:])]
        return (None, *grads)


class MRSTFTFn(Function):
    """auraloss.freq.MultiResolutionSTFTLoss as configured by multi_stft.yaml:1-18:
    A-weighting FIR, then per resolution an STFT computed as a strided conv against a windowed
    DFT basis, magnitude / spectral-convergence / log-magnitude statistics."""

    @staticmethod
    def forward(ctx, x: Tensor, y: Tensor, spec):
        B, C, L = x.shape
        x2, y2 = _c(x).view(B * C, 1, L), _c(y).view(B * C, 1, L)
        if spec.taps is not None:
            xf = ops.conv1d_fwd(x2, spec.taps, spec.fir_geom)
            yf = ops.conv1d_fwd(y2, spec.taps, spec.fir_geom)
        else:
            xf, yf = x2, y2
        nres = len(spec.res)
        stats = torch.zeros((3 * nres,), device=x.device, dtype=torch.float64)
        counts = []
        saved = []
        for r, (geom, basis, _) in enumerate(spec.res):
            if ops.STFT_VIA_FRAMES:
                # framing + DFT as a pointwise conv over the frame axis on the tensor-core kernel
                pw = ConvGeom(geom.K, geom.Cout, 1)
                wpw = basis.view(geom.Cout, geom.K, 1)
                X = ops.conv_fwd(ops.unfold_frames(xf, geom.K, geom.stride, geom.pad), wpw, pw, nsplit=3)
                Y = ops.conv_fwd(ops.unfold_frames(yf, geom.K, geom.stride, geom.pad), wpw, pw, nsplit=3)
            else:
                X = ops.conv1d_fwd(xf, basis, geom)
                Y = ops.conv1d_fwd(yf, basis, geom)
            ops.stft_stats(X, Y, spec.eps, stats[3 * r:3 * r + 3])
            counts.append(float(X.numel() // 2))
            saved += [X, Y]
        key = (B * C, L)
        counts_t = spec.counts_cache.get(key)
        if counts_t is None:
            counts_t = torch.tensor(counts, dtype=torch.float64).to(x.device)
            spec.counts_cache[key] = counts_t
        loss = ops.stft_finalize(stats, counts_t, nres, 1.0 / nres)
        ctx.spec, ctx.stats, ctx.counts, ctx.shape = spec, stats, counts, (B, C, L)
        ctx.save_for_backward(*saved)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, go):
        spec, stats, counts = ctx.spec, ctx.stats, ctx.counts
        B, C, L = ctx.shape
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("MRSTFT: gradient w.r.t. the target is not implemented")
        if not ctx.needs_input_grad[0]:
            return None, None, None
        go = _c(go).float()
        nres = len(spec.res)
        dxf = torch.zeros((B * C, 1, L), device=go.device, dtype=torch.float32)
        for r, (geom, _, basis_k) in enumerate(spec.res):
            X, Y = ctx.saved_tensors[2 * r], ctx.saved_tensors[2 * r + 1]
            dX = ops.stft_bwd(X, Y, spec.eps, stats[3 * r:3 * r + 3], counts[r], go, 1.0 / nres)
            if ops.STFT_VIA_FRAMES:
                pw = ConvGeom(geom.K, geom.Cout, 1)
                dU = ops.conv_dgrad(dX, spec.res[r][1].view(geom.Cout, geom.K, 1), None, pw, dX.shape[2], nsplit=3)
                ops.fold_frames(dU, L, geom.stride, geom.pad, dx=dxf)
            else:
                ops.conv1d_dgrad_scatter(dX, basis_k, geom, L, dxf)
        dx = ops.conv1d_dgrad(dxf, spec.taps, spec.fir_geom, L) if spec.taps is not None else dxf
        return dx.view(B, C, L), None, None


class WeightedSumFn(Function):
    """total = sum_i lam_i * loss_i on device scalars (eben.py:106 `sum(atomic_losses.values())`
    after the in-place scaling of eben.py:237-239; lam is detached there as well)."""

    @staticmethod
    def forward(ctx, lam: Optional[Tensor], *losses: Tensor):
        total, _ = ops.weighted_sum([_c(l).view(1) for l in losses], lam)
        ctx.lam, ctx.n = lam, len(losses)
        return total

    @staticmethod
    @once_differentiable
    def backward(ctx, go):
        g = ops.scalar_mul(_c(go).float().view(1), ctx.lam, ctx.n)
        return (None, *[g[i].view(()) for i in range(ctx.n)])
