"""CPU execution of the 16-byte slab staging written for the persistent tensor-core conv (csrc/ps_vec_stage.h,
EXPERIMENTAL, VBX_TC_PS_VEC=1): every unit a tile's MMAs read must hold the bf16 hi / lo split of the right input
sample (reflect / zero halos, batch-item boundaries, partial channel groups, unaligned fallbacks), and a software MMA
over the staged slab - rows read unit `row + tap*dil + a`, exactly the descriptor start addresses the kernel
forms - must reproduce the convolution."""
import ctypes
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from emu_util import ConvDesc, build_slab, ref_padded, tout



@pytest.fixture(scope="module")
def lib():
    return build_slab()


def bf16_rn(x):
    """fp32 -> bf16 bit pattern, round to nearest even (numpy restatement, independent of the C++ one)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) & 0xFFFF).astype(np.uint32)


def bf16_f(h):
    return (h.astype(np.uint32) << 16).view(np.float32)


def map_pos(p, Tin, refl):
    if p < 0:
        return -1 if p < -refl else -p
    if p >= Tin:
        return -1 if p >= Tin + refl else 2 * (Tin - 1) - p
    return p


CASES = [  # B, Cin, Cout, Tin, K, dil, pad, refl, groups, aligned
    (2, 32, 32, 300, 3, 3, 3, 3, 1, True),       # generator residual conv, reflect halo, several tiles per item
    (3, 64, 64, 200, 3, 9, 9, 9, 1, True),       # wide dilation, tiles straddling batch items
    (2, 40, 24, 132, 5, 2, 4, 0, 2, True),       # zero halo, partial channel groups (20 per group), 2 groups
    (2, 32, 32, 301, 3, 1, 1, 1, 1, True),       # length not a multiple of 4: every group takes the per-sample path
    (2, 16, 16, 260, 7, 1, 3, 3, 1, False),      # unaligned base pointer
    (1, 24, 8, 64, 3, 1, 1, 0, 1, True),         # shorter than one tile
    (2, 32, 32, 256, 1, 1, 0, 0, 1, True),       # pointwise (K = 1): every group aligned, no halo at all
]


@pytest.mark.parametrize("case", CASES, ids=str)
def test_vector_staging_fills_the_slab_and_reproduces_the_conv(lib, case):
    B, Cin, Cout, Tin, K, dil, pad, refl, groups, aligned = case
    To = tout(Tin, K, 1, dil, pad)
    d = ConvDesc(B, Cin, Cout, Tin, To, K, 1, dil, pad, refl, groups)
    torch.manual_seed(sum(case[:9]))
    x = torch.randn(B, Cin, Tin)
    w = torch.randn(Cout, Cin // groups, K) / (Cin // groups * K) ** 0.5
    xn = x.numpy()
    Cin_g, Cout_g = Cin // groups, Cout // groups
    want = F.conv1d(ref_padded(x.double(), pad, refl), w.double(), None, 1, 0, dil, groups).numpy()
    R = To + (K - 1) * dil
    tiles = (B * R + 127) // 128
    slot = np.empty(1 << 20, dtype=np.uint8)
    geom = (ctypes.c_int * 10)()
    got = np.zeros_like(want)
    for grp in range(groups):
        for tile in range(tiles):
            slot[:] = 0xAB                                    # garbage: every unit that is read must have been written
            rc = lib.emu_ps_vec_stage(ctypes.byref(d), grp, ctypes.c_void_p(x.data_ptr()), int(aligned), tile,
                                      ctypes.c_void_p(slot.ctypes.data), ctypes.c_longlong(slot.size), geom)
            assert rc == 0
            npos, groups4, units, a_stage, plane, half, R_c, ncg, a, q_origin = list(geom)
            assert R_c == R and npos == 128 + (K - 1) * dil and units >= npos + 3 and 0 <= a < 4
            assert q_origin == tile * 128 - a
            # (1) slab contents, unit by unit, against the numpy restatement
            A = np.zeros((npos, ncg * 16), dtype=np.float64)  # hi + lo of what the MMAs will read
            for j in range(npos):
                q = tile * 128 + j
                b, p = divmod(q, R)
                tau = map_pos(p - pad, Tin, refl)
                for c in range(ncg * 16):
                    val = xn[b, grp * Cin_g + c, tau] if (b < B and tau >= 0 and c < Cin_g) else np.float32(0)
                    hi = bf16_rn(np.float32(val)).reshape(1)
                    lo = bf16_rn(np.float32(val) - bf16_f(hi)).reshape(1)
                    hi, lo = int(hi[0]), int(lo[0])
                    cg, hf, e = c // 16, (c % 16) // 8, c % 8
                    off = cg * a_stage + hf * half + (a + j) * 16 + e * 2
                    ghi = int(slot[off]) | int(slot[off + 1]) << 8
                    glo = int(slot[off + plane]) | int(slot[off + plane + 1]) << 8
                    assert ghi == hi and glo == lo, (tile, j, c, val)
                    A[j, c] = float(bf16_f(np.array([ghi], np.uint32))[0]) + float(bf16_f(np.array([glo], np.uint32))[0])
            # (2) software MMA: output row r of the tile reads unit a + r + tap*dil for every tap
            wg = w[grp * Cout_g:(grp + 1) * Cout_g].double().numpy()          # (Cout_g, Cin_g, K)
            for r in range(128):
                b, t = divmod(tile * 128 + r, R)
                if b >= B or t >= To:
                    continue
                acc = np.zeros(Cout_g)
                for k in range(K):
                    acc += wg[:, :, k] @ A[r + k * dil, :Cin_g]
                got[b, grp * Cout_g:(grp + 1) * Cout_g, t] = acc
    # operands carry 16 mantissa bits (hi + lo): 1e-5-class agreement with the fp64 convolution
    assert np.abs(got - want).max() < 3e-5 * max(1.0, np.abs(want).max())
