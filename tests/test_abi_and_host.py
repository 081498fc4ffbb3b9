"""CPU-side checks: the C-ABI library loads and exports every symbol include/vbx.h declares, the
ctypes table mirrors the header, the drop-in modules keep the reference's state_dict layout and
initial values (golden sums minted from the reference itself), PQMF taps are bit-exact, conv
geometry extraction, and the data-parallel helpers under gloo with world_size 2."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "vbx.h")).read()
    return sorted(set(re.findall(r"VBX_API\s+[\w\s\*]+?\b(vbx_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from vibravox_b200 import _lib, build
    build.build_lib()
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_lib.SIGNATURES) == syms
    assert lib.vbx_abi_version() == _lib.ABI_VERSION == 2


def test_ops_refuse_cpu_tensors():
    from vibravox_b200 import _lib, ops
    with pytest.raises(_lib.VbxError):
        ops.leaky_relu_fwd(torch.zeros(8), 0.1)


def test_state_dict_layout_and_init_match_reference(golden_dir):
    from vibravox_b200.torch_modules.dnn.eben_discriminator import DiscriminatorEBENMultiScales
    from vibravox_b200.torch_modules.dnn.eben_generator import EBENGenerator
    gold = torch.load(os.path.join(golden_dir, "cfg1_forward.pt"))
    torch.manual_seed(gold["seed"])
    G = EBENGenerator(m=gold["m"], n=gold["n"], p=gold["p"])
    D = DiscriminatorEBENMultiScales(q=gold["q"], min_channels=gold["min_channels"])
    gs, ds = G.state_dict(), D.state_dict()
    assert list(gs) == list(gold["g_param_sums"]) and list(ds) == list(gold["d_param_sums"])
    for k, v in gs.items():
        assert float(v.double().sum()) == pytest.approx(gold["g_param_sums"][k], abs=1e-9), k
    for k, v in ds.items():
        assert float(v.double().sum()) == pytest.approx(gold["d_param_sums"][k], abs=1e-9), k
    assert torch.equal(G.pqmf.analysis_weights.data, gold["analysis_weights"])
    assert torch.equal(G.pqmf.synthesis_weights.data, gold["synthesis_weights"])
    assert torch.equal(torch.randn(1, 1, 16000)[0, 0, :8], gold["x_head"])
    assert G.multiple == 256 and G.p == 2
    x = torch.zeros(1, 1, 16000)
    assert G.cut_to_valid_length(x).shape[-1] == 15840
    assert G.cut_to_valid_length(torch.zeros(1, 1, 48000)).shape[-1] == 47840
    assert sum(p.numel() for p in G.parameters() if p.requires_grad) == 1945984
    assert sum(p.numel() for p in D.parameters()) == 23161344


def test_conv_geometry_extraction():
    from torch import nn
    from vibravox_b200.torch_modules.utils import conv_geom, conv_trans_geom
    g = conv_geom(nn.Conv1d(32, 32, 3, dilation=9, padding="same", padding_mode="reflect", bias=False))
    assert (g.pad, g.refl, g.dil, g.K) == (9, 9, 9, 3)
    g = conv_geom(nn.Conv1d(4, 24, 3, padding=1, dilation=2, groups=4), extra_reflect=1)
    assert (g.pad, g.refl) == (2, 1) and g.tout(11968) == 11968
    g = conv_geom(nn.Conv1d(16, 64, 41, stride=4, padding=20, groups=4))
    assert g.tout(47840) == 11960
    g, op = conv_trans_geom(nn.ConvTranspose1d(256, 128, 16, stride=8, padding=4))
    assert (g.Cin, g.Cout, g.pad, op) == (128, 256, 4, 0) and g.tout(1496) == 187


def test_mrstft_constants():
    from vibravox_b200.torch_modules.losses.mrstft_loss import MultiResolutionSTFTLoss, dft_basis
    L = MultiResolutionSTFTLoss(fft_sizes=(512, 1024, 2048), hop_sizes=(50, 120, 240), win_lengths=(240, 600, 1200),
                                sample_rate=16000, perceptual_weighting=True)
    assert L.fir_taps.shape == (1, 1, 101) and abs(float(L.fir_taps.sum()) - 0.0093) < 2e-4
    # the conv-with-basis formulation equals torch.stft on the CPU
    x = torch.randn(2, 4000)
    for n_fft, hop, win in ((512, 50, 240), (1024, 120, 600)):
        ref = torch.stft(x, n_fft, hop, win, torch.hann_window(win), return_complex=True)
        basis = dft_basis(n_fft, win)
        pad = win // 2
        xp = torch.nn.functional.pad(x.unsqueeze(1), (pad, pad), mode="reflect")
        got = torch.nn.functional.conv1d(xp, basis, stride=hop)
        bins = n_fft // 2 + 1
        assert got.shape == (2, 2 * bins, ref.shape[-1])
        assert (got[:, :bins] - ref.real).abs().max() < 2e-4
        assert (got[:, bins:] - ref.imag).abs().max() < 2e-4
    with pytest.raises(NotImplementedError):
        MultiResolutionSTFTLoss(w_lin_mag=1.0)


_GLOO_WORKER = r"""
import os, sys, torch
sys.path.insert(0, sys.argv[1])
from vibravox_b200 import parallel
rank, local, world = parallel.init_from_env("gloo")
assert world == 2
bucket = torch.full((1000,), float(rank + 1))
scale = parallel.allreduce_sum_(bucket)
assert scale == 0.5 and torch.all(bucket == 3.0)
assert parallel.max_over_ranks(float(rank)) == 1.0
assert parallel.rank_seed(42, rank) == 42 + rank
parallel.barrier()
print("rank", rank, "ok")
"""


def test_data_parallel_helpers_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = 29650 + os.getpid() % 200
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out
        assert "ok" in out
