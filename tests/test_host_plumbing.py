"""The Python side of the product (autograd.Functions, drop-in modules, FlatAdam gradient slots,
the training-step schedule) driven with torch-CPU stand-ins for the kernels (tests/cpu_shim.py)
and checked against the oracle / the reference goldens.  The kernels themselves are checked on the
GPU (-m gpu) and through the CPU emulator (test_emu_conv.py)."""
import os

import pytest
import torch

from cpu_shim import cpu_ops
from oracle import eben_oracle as O


def relerr(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def test_modules_forward_backward_match_oracle():
    from vibravox_b200.torch_modules.dnn.eben_discriminator import DiscriminatorEBENMultiScales
    from vibravox_b200.torch_modules.dnn.eben_generator import EBENGenerator
    from vibravox_b200.torch_modules.losses.feature_loss import FeatureLossForDiscriminatorMelganMultiScales
    from vibravox_b200.torch_modules.losses.hinge_loss import HingeLossForDiscriminatorMelganMultiScales
    torch.manual_seed(7)
    G, D = EBENGenerator(m=4, n=32, p=1), DiscriminatorEBENMultiScales(q=3, min_channels=24)
    gs = {k: v.detach().clone() for k, v in G.state_dict().items()}
    ds = {k: v.detach().clone() for k, v in D.state_dict().items()}
    body, air = O.synthetic_pairs(1, 4000, seed=3)
    g32 = {k: v.clone().requires_grad_(not k.startswith("pqmf.")) for k, v in gs.items()}
    d32 = {k: v.clone().requires_grad_(True) for k, v in ds.items()}
    x = O.cut_to_valid_length(body, 32, 4)
    a = O.cut_to_valid_length(air, 32, 4)
    y0, b0 = O.generator_forward(g32, x, 1)
    e0 = O.discriminator_forward(d32, b0, y0, 3, 24)
    r0 = O.discriminator_forward(d32, O.pqmf_analysis(a, g32["pqmf.analysis_weights"]), a, 3, 24)
    loss0 = O.feature_matching_loss(e0, r0) + O.hinge_loss(e0, 1) + O.hinge_loss(r0, -1)
    names = [k for k in g32 if g32[k].requires_grad]
    want = torch.autograd.grad(loss0, [g32[k] for k in names] + list(d32.values()))
    with cpu_ops():
        y, bands = G(G.cut_to_valid_length(body))
        assert relerr(y, y0) < 1e-5 and relerr(bands, b0) < 1e-5
        e = D(bands=bands, audio=y)
        r = D(bands=G.pqmf(a, "analysis"), audio=a)
        hinge = HingeLossForDiscriminatorMelganMultiScales()
        loss = FeatureLossForDiscriminatorMelganMultiScales()(e, r) + hinge(e, 1) + hinge(r, -1)
        assert float(loss) == pytest.approx(float(loss0), rel=1e-5)
        gp, dp = dict(G.named_parameters()), dict(D.named_parameters())
        got = torch.autograd.grad(loss, [gp[k] for k in names] + [dp[k] for k in d32])
    for n, u, v in zip(names + list(d32), got, want):
        assert relerr(u, v) < 1e-2, (n, relerr(u, v))   # fp32 vs fp32, different summation order


@pytest.mark.parametrize("schedule", ["shared", "reference"])
def test_training_step_schedule_matches_reference_golden(golden_dir, schedule):
    import vibravox_b200
    gold = torch.load(os.path.join(golden_dir, "train_step.pt"))
    body, air = O.synthetic_pairs(gold["B"], gold["S"], seed=gold["data_seed"])
    with cpu_ops():
        lm = vibravox_b200.build_model(seed=gold["model_seed"], device="cpu")
        lm.schedule = schedule
        for it in range(2):
            out = lm.training_step({"audio_body_conducted": body, "audio_airborne": air})
            want = gold["steps"][it]
            for k, v in want["logs"].items():
                got = float(lm.logged["train/" + k])
                assert got == pytest.approx(v, rel=3e-4, abs=3e-5), (it, k, got, v)
            for a, b in zip(lm.atomic_norms_old.tolist(), want["norms_old"]):
                assert a == pytest.approx(b, rel=5e-4)
        # every trainable parameter moved, gradient bucket is clean, D untouched flags restored
        assert all(p.requires_grad for p in lm.discriminator.parameters())
        assert float(lm.generator_optimizer.grad.abs().sum()) == 0.0
        assert int(lm.generator_optimizer.step_count[0]) == 2 and int(lm.discriminator_optimizer.step_count[0]) == 2
    gsd = lm.generator.state_dict()
    for k, v in gold["g_param_sums_after"].items():
        n = gsd[k].numel()
        assert abs(float(gsd[k].double().sum()) - v) <= 5e-3 * n ** 0.5 + 1e-3, k


def test_fused_discriminator_chain_backward_equals_the_separate_passes(golden_dir, monkeypatch):
    """functional.Flags chain contract: with the fusion on, no L1-pair backward and no LeakyReLU backward with an input
    gradient run for the discriminator stages (the conv input-gradient epilogues do that work), every registered
    feature-matching term is consumed, and the step lands on the same parameters as with the fusion off."""
    import vibravox_b200
    from vibravox_b200 import ops
    from vibravox_b200.lightning_modules import eben as eben_mod
    gold = torch.load(os.path.join(golden_dir, "train_step.pt"))
    body, air = O.synthetic_pairs(gold["B"], gold["S"], seed=gold["data_seed"])
    results = {}
    for fused in (True, False):
        monkeypatch.setattr(eben_mod, "_CHAIN_FUSION", fused)
        with cpu_ops():
            calls = {"l1_pair_bwd": 0, "lrelu_dx_from_y": 0, "gate": 0, "fm_gate_bwd": 0}
            real_l1, real_lr, real_dg, real_fg = ops.l1_pair_bwd, ops.leaky_relu_bwd, ops.conv1d_dgrad, ops.fm_gate_bwd

            def l1(*a, **k):
                calls["l1_pair_bwd"] += 1
                return real_l1(*a, **k)

            def lr(dy, ref, slope, mask=None, dbias=None, want_dx=True):
                calls["lrelu_dx_from_y"] += int(want_dx and ref is not None and slope == 0.2)
                return real_lr(dy, ref, slope, mask=mask, dbias=dbias, want_dx=want_dx)

            def dg(*a, gate=None, **k):
                calls["gate"] += int(gate is not None)
                return real_dg(*a, gate=gate, **k)

            def fg(*a, **k):
                calls["fm_gate_bwd"] += 1
                return real_fg(*a, **k)
            ops.l1_pair_bwd, ops.leaky_relu_bwd, ops.conv1d_dgrad, ops.fm_gate_bwd = l1, lr, dg, fg
            lm = vibravox_b200.build_model(seed=gold["model_seed"], device="cpu")
            lm.schedule = "shared"
            lm.training_step({"audio_body_conducted": body, "audio_airborne": air})
            results[fused] = (calls, {k: v.clone() for k, v in lm.state_dict().items()},
                              {k: float(v) for k, v in lm.logged.items()})
    on, off = results[True][0], results[False][0]
    assert off["gate"] == 0 and off["l1_pair_bwd"] > 0 and off["lrelu_dx_from_y"] > 0
    assert on["l1_pair_bwd"] == 0 and on["lrelu_dx_from_y"] == 0 and on["gate"] > 0
    # 4 chains: in the feature-matching pass the last feature of each chain has no downstream conv in the pass
    assert on["fm_gate_bwd"] == 4
    for k, v in results[False][1].items():
        if v.is_floating_point():
            assert relerr(results[True][1][k], v) < 2e-5, k
    for k, v in results[False][2].items():
        assert results[True][2][k] == pytest.approx(v, rel=1e-5, abs=1e-6), k


def test_chain_contract_violations_fail_loudly():
    """functional.Flags: a feature-matching term that no conv stage applies is an error at the end of the backward pass (not
    a silently missing gradient), and the fused feature-matching backward refuses gradients for the reference features."""
    from vibravox_b200.functional import FeatureMatchingFn, Flags
    with cpu_ops():
        a = torch.randn(2, 3, 16, requires_grad=True)
        b = torch.randn(2, 3, 16)
        Flags.chain_reset()
        loss = FeatureMatchingFn.apply(0.5, 1, True, a * 1.0, b)          # fused, but `a * 1.0` is not a ConvFn stage output
        loss.backward()
        with pytest.raises(RuntimeError, match="feature-matching gradient terms"):
            Flags.chain_check()
        assert not Flags.pending                                          # (chain_check leaves the tables clean)
        b2 = torch.randn(2, 3, 16, requires_grad=True)
        loss = FeatureMatchingFn.apply(0.5, 1, True, a * 1.0, b2 * 1.0)
        with pytest.raises(NotImplementedError):
            loss.backward()
        Flags.chain_reset()
        # the unfused form is the ordinary differentiable loss
        a.grad = None
        FeatureMatchingFn.apply(0.5, 1, False, a * 1.0, b).backward()
        ref = a.detach().clone().requires_grad_(True)
        (0.5 * (ref - b).abs().mean() / ref.abs().mean()).backward()
        assert torch.allclose(a.grad, ref.grad, atol=1e-6)


def test_flat_adam_state_dict_is_torch_adam_compatible():
    """FlatAdam.state_dict() is the layout torch.optim.Adam writes: a torch Adam loaded from it takes the same next
    step, and FlatAdam loaded from a torch Adam state continues that optimizer's trajectory."""
    from vibravox_b200.optim import FlatAdam
    torch.manual_seed(0)
    shapes = [(5, 3, 2), (7,), (4, 1, 1)]
    with cpu_ops():
        ps = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
        flat = FlatAdam(ps, lr=3e-4, betas=(0.5, 0.9))
        assert flat.state_dict()["state"] == {}                       # fresh, like torch Adam
        grads = [[torch.randn(s) for s in shapes] for _ in range(4)]
        for gs in grads[:3]:
            for p, g in zip(ps, gs):
                p.grad = g.clone()
            flat.step(); flat.zero_grad()
        sd = flat.state_dict()
        assert sorted(sd["state"]) == [0, 1, 2] and float(sd["state"][0]["step"]) == 3.0
        assert sd["param_groups"][0]["betas"] == (0.5, 0.9) and sd["param_groups"][0]["params"] == [0, 1, 2]
        # torch Adam continues from FlatAdam's state
        qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
        adam = torch.optim.Adam(qs, lr=3e-4, betas=(0.5, 0.9))
        adam.load_state_dict(sd)
        for p, q, g in zip(ps, qs, grads[3]):
            p.grad, q.grad = g.clone(), g.clone()
        flat.step(); adam.step()
        for p, q in zip(ps, qs):
            assert torch.allclose(p, q, rtol=0, atol=2e-7)
        # FlatAdam continues from torch Adam's state (a reference checkpoint), moments copied in place
        rs = [torch.nn.Parameter(q.detach().clone()) for q in qs]
        flat2 = FlatAdam(rs, lr=1.0, betas=(0.9, 0.999))
        flat2.materialize()
        ptr = flat2.exp_avg.data_ptr()
        flat2.load_state_dict(adam.state_dict())
        assert flat2.exp_avg.data_ptr() == ptr and int(flat2.step_count[0]) == 4
        assert flat2.param_groups[0]["lr"] == 3e-4 and flat2.param_groups[0]["betas"] == (0.5, 0.9)
        g5 = [torch.randn(s) for s in shapes]
        for q, r, g in zip(qs, rs, g5):
            q.grad, r.grad = g.clone(), g.clone()
        adam.step(); flat2.step()
        for q, r in zip(qs, rs):
            assert torch.allclose(q, r, rtol=0, atol=2e-7)
        with pytest.raises(ValueError):
            flat2.load_state_dict(torch.optim.Adam(qs[:2]).state_dict())


def test_flat_adam_keeps_frozen_parameters_in_the_index_space_of_torch_adam():
    """generator.parameters() starts with the two requires_grad=False PQMF banks (pqmf.py:51-56): torch.optim.Adam
    indexes all 92 tensors (state keys start at 2); FlatAdam must use the same indices in both directions."""
    from vibravox_b200.optim import FlatAdam
    torch.manual_seed(1)
    shapes = [(4, 1, 8), (4, 1, 8), (6, 2, 3), (5,)]

    def make():
        ps = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
        ps[0].requires_grad_(False); ps[1].requires_grad_(False)
        return ps

    with cpu_ops():
        ps, qs = make(), None
        torch.manual_seed(1)
        qs = make()
        flat, adam = FlatAdam(ps, lr=3e-4, betas=(0.5, 0.9)), torch.optim.Adam(qs, lr=3e-4, betas=(0.5, 0.9))
        assert len(flat.param_groups[0]["params"]) == 4 and len(flat.params) == 2
        # a fresh torch Adam state dict (what the advisor's repro loads) is accepted
        flat.load_state_dict(adam.state_dict())
        for _ in range(2):
            gs = [torch.randn(s) for s in shapes[2:]]
            for p, q, g in zip(ps[2:], qs[2:], gs):
                p.grad, q.grad = g.clone(), g.clone()
            flat.step(); flat.zero_grad(); adam.step()
        sd, ref = flat.state_dict(), adam.state_dict()
        assert sorted(sd["state"]) == sorted(ref["state"]) == [2, 3]
        assert sd["param_groups"][0]["params"] == ref["param_groups"][0]["params"] == [0, 1, 2, 3]
        for i in (2, 3):
            assert torch.allclose(sd["state"][i]["exp_avg"], ref["state"][i]["exp_avg"], atol=1e-7)
        # both directions
        adam2 = torch.optim.Adam(make(), lr=1.0); adam2.load_state_dict(sd)
        flat2 = FlatAdam(make(), lr=1.0); flat2.load_state_dict(ref)
        assert int(flat2.step_count[0]) == 2 and flat2.param_groups[0]["lr"] == 3e-4
        assert torch.equal(ps[0], qs[0])                              # frozen tensors untouched


def test_flat_adam_round_trips_with_adam_over_generator_parameters():
    from vibravox_b200.optim import FlatAdam
    from vibravox_b200.torch_modules.dnn.eben_generator import EBENGenerator
    torch.manual_seed(3)
    G = EBENGenerator(m=4, n=32, p=2)
    with cpu_ops():
        flat = FlatAdam(G.parameters(), lr=3e-4, betas=(0.5, 0.9))
        adam = torch.optim.Adam(G.parameters(), lr=3e-4, betas=(0.5, 0.9))
        flat.load_state_dict(adam.state_dict())                      # the advisor's failing call
        assert len(flat.all_params) == 92 and len(flat.params) == 90
        flat.materialize()
        flat.grad.fill_(1e-3)
        flat.step()
        sd = flat.state_dict()
        assert min(sd["state"]) == 2 and len(sd["state"]) == 90
        adam.load_state_dict(sd)


def test_run_py_trainer_checkpoint_resume(tmp_path):
    """run.py's override grammar -> Trainer.fit -> last.ckpt -> `ckpt_path=last`: three steps in one go and two steps
    + resume + one step end on the same parameters, Adam moments and balancing state."""
    import run
    common = ["lightning_datamodule=bwe", "lightning_module=eben", "lightning_datamodule.batch_size=1",
              "lightning_datamodule.collate_strategy=constant_length-250-ms", "lightning_module.generator.p=1",
              "lightning_module.discriminator.q=3", "++trainer.accelerator=cpu", "++trainer.log_every_n_steps=1000"]
    with cpu_ops():
        a = run.main(common + ["++trainer.max_steps=3", f"++trainer.default_root_dir={tmp_path}/a"])
        run.main(common + ["++trainer.max_steps=2", f"++trainer.default_root_dir={tmp_path}/b"])
        ck = torch.load(tmp_path / "b" / "checkpoints" / "last.ckpt", weights_only=False)
        assert ck["global_step"] == 2 and len(ck["optimizer_states"]) == 2 and "vbx_balancing" in ck
        assert any(k.endswith("parametrizations.weight.original0") for k in ck["state_dict"])
        # generator state indices are torch.optim.Adam's: 0 and 1 are the frozen PQMF banks (no state), 2 = first_conv
        assert min(ck["optimizer_states"][0]["state"]) == 2 and min(ck["optimizer_states"][1]["state"]) == 0
        assert float(ck["optimizer_states"][0]["state"][2]["step"]) == 2.0
        b = run.main(common + ["++trainer.max_steps=3", f"++trainer.default_root_dir={tmp_path}/b", "+ckpt_path=last"])
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert torch.allclose(sa[k], sb[k], rtol=0, atol=1e-6), k
    for oa, ob in zip(a.configure_optimizers(), b.configure_optimizers()):
        assert int(oa.step_count[0]) == int(ob.step_count[0]) == 3
        assert torch.allclose(oa.exp_avg, ob.exp_avg, rtol=0, atol=1e-7)
        assert torch.allclose(oa.exp_avg_sq, ob.exp_avg_sq, rtol=0, atol=1e-9)
    assert torch.allclose(a.atomic_norms_old, b.atomic_norms_old, rtol=1e-6)
    for k in a.logged:
        assert float(a.logged[k]) == pytest.approx(float(b.logged[k]), rel=1e-5), k


def test_eval_step_logs_the_pre_update_losses_of_a_training_step(golden_dir):
    """common_eval_step / validation_step / test_step (eben.py:132-165, base_se.py:132-136): on a fresh model the
    evaluation losses are the atomic losses the first training step logs before it updates anything - and those are
    pinned to the reference by the golden training logs."""
    import vibravox_b200
    gold = torch.load(os.path.join(golden_dir, "train_step.pt"))
    body, air = O.synthetic_pairs(gold["B"], gold["S"], seed=gold["data_seed"])
    batch = {"audio_body_conducted": body, "audio_airborne": air}
    with cpu_ops():
        ev = vibravox_b200.build_model(seed=gold["model_seed"], device="cpu")
        out = ev.validation_step(batch, 0)
        assert set(out) == {"corrupted", "enhanced", "reference"} and out["enhanced"].shape == out["reference"].shape
        assert not out["enhanced"].requires_grad
        keys = {"generator": ("reconstructive_loss_freq", "feature_matching_loss", "adv_loss_gen"),
                "discriminator": ("real_loss", "fake_loss")}
        for net, names in keys.items():
            for n in names:
                want = gold["steps"][0]["logs"][f"{net}/{n}"]
                assert float(ev.logged[f"validation/{net}/{n}"]) == pytest.approx(want, rel=3e-4, abs=3e-5), (net, n)
        ev.dataloader_names = ["speech_clean", "speech_noisy"]
        ev.logged.clear()
        only_body = ev.test_step({"audio_body_conducted": body}, 3, dataloader_idx=1)
        assert set(only_body) == {"corrupted", "enhanced"} and not ev.logged       # no reference: nothing to log
        ev.test_step(batch, 0, dataloader_idx=1)
        assert "test/generator/adv_loss_gen/speech_noisy" in ev.logged


def test_run_py_noisy_bwe_datamodule_steps():
    """BASELINE config 4 through run.py: the noisy-BWE datamodule (noise mix + joint crop on the device, here the
    host stand-in of vbx_noise_mix_crop) feeds the same training step; batches are (B, 1, samples) pairs, differ
    from step to step, and the losses stay finite."""
    import run
    from vibravox_b200.data import SyntheticNoisyBWEDataModule
    with cpu_ops():
        dm = SyntheticNoisyBWEDataModule(sample_rate=16000, batch_size=2, collate_strategy="constant_length-250-ms",
                                         noise_seconds=1.0, seed=3)
        it = dm.batches(torch.device("cpu"), rank=0)
        b0, b1 = next(it), next(it)
        assert b0["audio_body_conducted"].shape == b0["audio_airborne"].shape == (2, 1, 4000)
        assert not torch.equal(b0["audio_body_conducted"], b1["audio_body_conducted"])
        other = next(SyntheticNoisyBWEDataModule(sample_rate=16000, batch_size=2, noise_seconds=1.0, seed=3,
                                                 collate_strategy="constant_length-250-ms").batches("cpu", rank=1))
        assert not torch.equal(other["audio_airborne"], b0["audio_airborne"])        # each rank draws its own stream
        lm = run.main(["lightning_datamodule=noisybwe", "lightning_module=eben", "lightning_datamodule.batch_size=1",
                       "lightning_datamodule.collate_strategy=constant_length-250-ms",
                       "lightning_datamodule.noise_seconds=1", "lightning_module.generator.p=1",
                       "lightning_module.discriminator.q=3", "++trainer.accelerator=cpu", "++trainer.max_steps=2",
                       "++trainer.log_every_n_steps=1000"])
    assert int(lm.generator_optimizer.step_count[0]) == 2
    assert all(torch.isfinite(v).all() for v in lm.logged.values()) and len(lm.logged) == 7


def test_graph_mode_decision_table(monkeypatch):
    """Which launch form training_step_graphed picks: one graph on a single GPU, three graphs split at the gradient
    all-reduces under NCCL data parallelism (the default for world_size > 1), eager launches whenever the step
    takes a per-step decision on the host or the optimizer is not FlatAdam."""
    import vibravox_b200
    with cpu_ops():
        lm = vibravox_b200.build_model(seed=1, device="cpu")
    for k in ("VBX_GRAPH_SEGMENTS", "VBX_GRAPH_DDP"):
        monkeypatch.delenv(k, raising=False)
    assert lm.graph_mode() == "whole" and lm.graph_capturable()
    monkeypatch.setenv("VBX_GRAPH_SEGMENTS", "1")
    assert lm.graph_mode() == "segments"
    monkeypatch.delenv("VBX_GRAPH_SEGMENTS")
    # world_size 2 over NCCL
    import torch.distributed as dist
    monkeypatch.setattr(dist, "is_initialized", lambda: True)
    monkeypatch.setattr(dist, "get_world_size", lambda *a, **k: 2)
    monkeypatch.setattr(dist, "get_backend", lambda *a, **k: "nccl")
    assert lm.graph_mode() == "segments"
    monkeypatch.setenv("VBX_GRAPH_SEGMENTS", "0")
    assert lm.graph_mode() == "eager" and not lm.graph_capturable()
    monkeypatch.delenv("VBX_GRAPH_SEGMENTS")
    monkeypatch.setenv("VBX_GRAPH_DDP", "1")
    assert lm.graph_mode() == "whole"
    monkeypatch.delenv("VBX_GRAPH_DDP")
    monkeypatch.setattr(dist, "get_backend", lambda *a, **k: "gloo")
    assert lm.graph_mode() == "eager"
    monkeypatch.setattr(dist, "get_backend", lambda *a, **k: "nccl")
    # host-side decisions
    lm.update_discriminator_ratio = 0.5
    assert lm.graph_mode() == "eager"
    lm.update_discriminator_ratio = 1.0
    lm._graph_failed = True
    assert lm.graph_mode() == "eager"
    lm._graph_failed = False
    lm.generator_optimizer = torch.optim.Adam(lm.generator.parameters(), lr=3e-4)
    assert lm.graph_mode() == "eager"
