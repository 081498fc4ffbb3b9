"""Pin the oracle against the real reference and mint the golden fixtures.

Runs ONLY in the build container (needs /root/reference, read-only).  It
  1. imports the reference's own EBENGenerator / DiscriminatorEBENMultiScales /
     PseudoQMFBanks / feature + hinge losses unmodified,
  2. imports the reference's own ``lightning_modules/eben.py`` under a minimal stand-in
     for the (absent) ``lightning`` package, so the *reference's* training_step /
     compute_atomic_losses / dynamically_balance_losses code runs as written,
     with auraloss (absent, third-party) replaced by the oracle's restatement,
  3. checks oracle/eben_oracle.py against all of it (bit-exact init, forward within
     fp32 noise, step losses / gradient norms / post-Adam parameters),
  4. writes small fixtures to tests/golden/*.pt for the GPU box, where
     /root/reference does not exist.

Usage:  python -m oracle.make_goldens
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from functools import partial

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import eben_oracle as O  # noqa: E402


def _install_lightning_standin():
    """Minimal LightningModule semantics used by eben.py."""

    class LightningModule(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.logged = {}
            self._toggled = {}

        def optimizers(self, use_pl_optimizer=True):
            return self.configure_optimizers()

        def log(self, name, value, **kw):
            self.logged[name] = float(value)

        def manual_backward(self, loss):
            loss.backward()

        def toggle_optimizer(self, optimizer):
            # lightning.pytorch.core.module.LightningModule.toggle_optimizer:
            # freeze every parameter that does not belong to `optimizer`.
            mine = {id(p) for g in optimizer.param_groups for p in g["params"]}
            saved = {}
            for opt in self.configure_optimizers():
                for g in opt.param_groups:
                    for p in g["params"]:
                        if id(p) in saved:
                            continue
                        saved[id(p)] = (p, p.requires_grad)
            for p, flag in saved.values():
                if id(p) not in mine:
                    p.requires_grad = False
            self._toggled = saved

        def untoggle_optimizer(self, optimizer):
            for p, flag in self._toggled.values():
                p.requires_grad = flag
            self._toggled = {}

    lightning = types.ModuleType("lightning")
    lightning.LightningModule = LightningModule
    pt = types.ModuleType("lightning.pytorch")
    ut = types.ModuleType("lightning.pytorch.utilities")
    ty = types.ModuleType("lightning.pytorch.utilities.types")
    ty.STEP_OUTPUT = object
    sys.modules.update({"lightning": lightning, "lightning.pytorch": pt,
                        "lightning.pytorch.utilities": ut,
                        "lightning.pytorch.utilities.types": ty})

    base = types.ModuleType("vibravox.lightning_modules.base_se")

    class BaseSELightningModule(LightningModule):
        def __init__(self, sample_rate, description):
            super().__init__()
            self.sample_rate, self.description = sample_rate, description

    base.BaseSELightningModule = BaseSELightningModule
    sys.modules["vibravox.lightning_modules.base_se"] = base


def _load_reference():
    sys.path.insert(0, REF)
    _install_lightning_standin()
    from vibravox.torch_modules.dnn.eben_generator import EBENGenerator
    from vibravox.torch_modules.dnn.eben_discriminator import DiscriminatorEBENMultiScales
    from vibravox.torch_modules.dsp.pqmf import PseudoQMFBanks
    from vibravox.torch_modules.losses.feature_loss import FeatureLossForDiscriminatorMelganMultiScales
    from vibravox.torch_modules.losses.hinge_loss import HingeLossForDiscriminatorMelganMultiScales
    spec = importlib.util.spec_from_file_location(
        "vibravox.lightning_modules.eben", os.path.join(REF, "vibravox/lightning_modules/eben.py"))
    eben = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(eben)
    return dict(G=EBENGenerator, D=DiscriminatorEBENMultiScales, PQMF=PseudoQMFBanks,
                FM=FeatureLossForDiscriminatorMelganMultiScales,
                HINGE=HingeLossForDiscriminatorMelganMultiScales, LM=eben.EBENLightningModule)


class _OracleMRSTFT(torch.nn.Module):
    """Stand-in for auraloss.freq.MultiResolutionSTFTLoss (absent)."""

    def __init__(self):
        super().__init__()
        self.taps = O.a_weighting_fir()

    def forward(self, x, y):
        return O.mrstft_loss(x, y, self.taps)


def relerr(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(GOLD, exist_ok=True)
    R = _load_reference()
    report = {}

    # ---- 1. PQMF design: bit-exact
    for (m, n) in [(4, 32), (4, 64), (8, 64), (2, 16)]:
        ref = R["PQMF"](decimation=m, kernel_size=n)
        wa, ws, cut = O.pqmf_design(m, n)
        assert torch.equal(wa, ref.analysis_weights.data), (m, n)
        assert torch.equal(ws, ref.synthesis_weights.data), (m, n)
        assert cut == ref._cutoff_ratio
        report[f"pqmf_{m}_{n}_cutoff"] = cut
    print("pqmf design bit-exact")
    ref_pqmf = R["PQMF"](decimation=4, kernel_size=32)
    torch.manual_seed(3)
    sig = torch.rand(4, 1, 48008)
    dec = ref_pqmf(sig, "analysis")
    rec = ref_pqmf(dec, "synthesis").sum(1, keepdim=True)
    snr = 10 * torch.log10((rec ** 2).mean() / ((sig - rec) ** 2).mean()).item()
    report["pqmf_4_32_snr_db_rand_4x48008_seed3"] = snr
    assert torch.equal(O.pqmf_analysis(sig, ref_pqmf.analysis_weights), dec)

    # ---- 2. init: bit-exact for several configs
    for cfg in [dict(m=4, n=32, p=2, q=4, mc=24), dict(m=4, n=32, p=1, q=3, mc=24)]:
        torch.manual_seed(42)
        G = R["G"](m=cfg["m"], n=cfg["n"], p=cfg["p"])
        D = R["D"](q=cfg["q"], min_channels=cfg["mc"])
        torch.manual_seed(42)
        gs = O.init_generator_state(cfg["m"], cfg["n"], cfg["p"])
        ds = O.init_discriminator_state(cfg["q"], cfg["mc"])
        rg, rd = G.state_dict(), D.state_dict()
        assert list(rg.keys()) == list(gs.keys()), "generator key order"
        assert list(rd.keys()) == list(ds.keys()), "discriminator key order"
        for k in rg:
            assert torch.equal(rg[k], gs[k]), k
        for k in rd:
            assert torch.equal(rd[k], ds[k]), k
    print("init bit-exact (keys, order, values)")

    # ---- 3. config 1: G forward on 1x1x16000 (+ D forward on its outputs)
    torch.manual_seed(42)
    G = R["G"](m=4, n=32, p=2)
    D = R["D"](q=4, min_channels=24)
    x = torch.randn(1, 1, 16000)
    with torch.no_grad():
        xc = G.cut_to_valid_length(x)
        y, bands = G(xc)
        emb = D(bands=bands, audio=y)
    torch.manual_seed(42)
    gs = O.init_generator_state(4, 32, 2)
    ds = O.init_discriminator_state(4, 24)
    x2 = torch.randn(1, 1, 16000)
    assert torch.equal(x, x2)
    with torch.no_grad():
        oy, ob = O.generator_forward(gs, O.cut_to_valid_length(x2, 32, 4), 2)
        oe = O.discriminator_forward(ds, ob, oy, 4, 24)
    report["cfg1_oracle_vs_ref_enhanced"] = relerr(oy, y)
    report["cfg1_oracle_vs_ref_bands"] = relerr(ob, bands)
    assert relerr(oy, y) < 1e-6 and relerr(ob, bands) < 1e-6
    for sa, sb in zip(oe, emb):
        assert len(sa) == len(sb)
        for a, b in zip(sa, sb):
            assert a.shape == b.shape and relerr(a, b) < 2e-6, relerr(a, b)
    fm_ref = R["FM"]()(emb, [[t.flip(-1) for t in s] for s in emb])
    fm_or = O.feature_matching_loss(emb, [[t.flip(-1) for t in s] for s in emb])
    assert abs(float(fm_ref) - float(fm_or)) < 1e-6 * abs(float(fm_ref))
    for tgt in (1, -1):
        assert abs(float(R["HINGE"]()(emb, tgt)) - float(O.hinge_loss(emb, tgt))) < 1e-7
    torch.save({
        "seed": 42, "m": 4, "n": 32, "p": 2, "q": 4, "min_channels": 24,
        "x_head": x[0, 0, :8].clone(), "enhanced": y.clone(), "bands": bands.clone(),
        "certainties": [s[-1].clone() for s in emb],
        "emb_absmean": [[float(t.abs().mean()) for t in s] for s in emb],
        "emb_shapes": [[tuple(t.shape) for t in s] for s in emb],
        "first_conv_w0": G.first_conv.weight[0, 0].clone(),
        "analysis_weights": G.pqmf.analysis_weights.data.clone(),
        "synthesis_weights": G.pqmf.synthesis_weights.data.clone(),
        "g_param_sums": {k: float(v.double().sum()) for k, v in G.state_dict().items()},
        "d_param_sums": {k: float(v.double().sum()) for k, v in D.state_dict().items()},
    }, os.path.join(GOLD, "cfg1_forward.pt"))
    print("config-1 forward: oracle == reference", report["cfg1_oracle_vs_ref_enhanced"])

    # ---- 4. training step: the reference's own eben.py code vs the oracle, 2 steps
    B, S = 2, 8000
    body, air = O.synthetic_pairs(B, S, seed=7)
    torch.manual_seed(42)
    G = R["G"](m=4, n=32, p=2)
    D = R["D"](q=4, min_channels=24)
    adam = partial(torch.optim.Adam, lr=3e-4, betas=(0.5, 0.9))
    lm = R["LM"](sample_rate=16000, generator=G, discriminator=D, generator_optimizer=adam,
                 discriminator_optimizer=adam, reconstructive_loss_freq_fn=_OracleMRSTFT(),
                 feature_matching_loss_fn=R["FM"](), adversarial_loss_fn=R["HINGE"](),
                 dynamic_loss_balancing="ema", beta_ema=0.9, update_discriminator_ratio=1.0,
                 description="golden")
    orc = O.OracleEBENStep(seed=42)
    steps = []
    for it in range(2):
        lm.logged = {}
        out = lm.training_step({"audio_body_conducted": body, "audio_airborne": air})
        ref_logs = {k.replace("train/", ""): v for k, v in lm.logged.items()}
        ologs = orc.step(body, air)
        for k, v in ref_logs.items():
            assert abs(v - ologs[k]) <= 2e-5 * max(1.0, abs(v)), (it, k, v, ologs[k])
        steps.append({"logs": ref_logs, "norms_old": [float(t) for t in lm.atomic_norms_old],
                      "enhanced_head": out["enhanced"][0, 0, :64].detach().clone()})
        for a, b in zip(lm.atomic_norms_old, orc.norms_old):
            assert abs(float(a) - float(b)) <= 1e-4 * abs(float(a)), (float(a), float(b))
    # post-Adam parameters: statistical agreement (SURVEY 8c tolerance guidance)
    worst = 0.0
    for k, v in G.state_dict().items():
        worst = max(worst, float((v - orc.g[k].detach()).abs().max()))
    for k, v in D.state_dict().items():
        worst = max(worst, float((v - orc.d[k].detach()).abs().max()))
    report["step_oracle_vs_ref_max_param_abs_diff_after_2_steps"] = worst
    assert worst <= 2 * 2 * 3e-4 + 1e-6
    torch.save({"B": B, "S": S, "data_seed": 7, "model_seed": 42, "steps": steps,
                "g_param_sums_after": {k: float(v.double().sum()) for k, v in G.state_dict().items()},
                "d_param_absmean_after": {k: float(v.double().abs().mean()) for k, v in D.state_dict().items()}},
               os.path.join(GOLD, "train_step.pt"))
    print("training step: oracle == reference eben.py", steps[0]["logs"])

    # ---- 5. fp64 bracketing data for gradients (one step, keep grads)
    o32 = O.OracleEBENStep(seed=42)
    o64 = O.OracleEBENStep(seed=42, dtype=torch.float64)
    o32.step(body, air, keep_grads=True)
    o64.step(body, air, keep_grads=True)
    gnoise = {k: relerr(o32.last["g_grads"][k], o64.last["g_grads"][k]) for k in o32.last["g_grads"]}
    dnoise = {k: relerr(o32.last["d_grads"][k], o64.last["d_grads"][k]) for k in o32.last["d_grads"]}
    torch.save({"B": B, "S": S, "data_seed": 7,
                "g_grad_norm64": {k: float(v.norm()) for k, v in o64.last["g_grads"].items()},
                "d_grad_norm64": {k: float(v.norm()) for k, v in o64.last["d_grads"].items()},
                "g_fp32_vs_fp64": gnoise, "d_fp32_vs_fp64": dnoise,
                "norms64": o64.last["norms"], "lambdas64": o64.last["lambdas"]},
               os.path.join(GOLD, "grad_bracket.pt"))
    report["g_grad_fp32_vs_fp64_median"] = sorted(gnoise.values())[len(gnoise) // 2]
    report["g_grad_fp32_vs_fp64_max"] = max(gnoise.values())
    report["d_grad_fp32_vs_fp64_max"] = max(dnoise.values())

    torch.save(report, os.path.join(GOLD, "oracle_pin_report.pt"))
    for k, v in report.items():
        print(f"  {k}: {v}")
    print("goldens written to", GOLD)


if __name__ == "__main__":
    main()
