"""CPU check of the arithmetic fact the batched path's bf16 mode rests on (vrod_b200/csrc/knn_batched.cu, split3):
an f32 value v splits into three bf16 parts a = bf16(v), b = bf16(v - a), c = bf16(v - a - b) with (a + b) + c == v
exactly (round to nearest, 3 x 8 significant bits), so the thresholds and the ||x||^2/2 terms folded into the
contraction as aux columns stand for exactly the f32 values the finish kernel computed."""
import numpy as np
import torch


def split3(v):
    a = v.to(torch.bfloat16).float()
    r1 = v - a
    b = r1.to(torch.bfloat16).float()
    r2 = r1 - b
    c = r2.to(torch.bfloat16).float()
    return a, b, c


def test_three_bf16_parts_rebuild_an_f32_exactly():
    g = torch.Generator().manual_seed(1)
    v = torch.cat([
        torch.randn(400_000, generator=g) * torch.logspace(-25, 25, 400_000),   # wide range of magnitudes
        torch.rand(200_000, generator=g) * 100.0,                               # typical ||x||^2/2 and thresholds
        -torch.rand(200_000, generator=g),                                      # cosine surrogates
        torch.tensor([0.0, 1.0, -1.0, 3.0e38, 1.0e-30, 2.0 ** -100, 1.0 + 2.0 ** -23]),
    ])
    a, b, c = split3(v)
    assert torch.equal((a + b) + c, v)
    # every part is a bf16 value (the tensor core multiplies it by 1.0 exactly)
    for part in (a, b, c):
        assert torch.equal(part.to(torch.bfloat16).float(), part)


def test_mirror_geometry_matches_the_library():
    """kd = dim rounded up to 16 data columns + 16 aux columns; rows of 2*ld_h bytes keep TMA's 16-byte stride rule."""
    for dim in (1, 8, 15, 16, 17, 64, 100, 128, 130, 768, 1536):
        kd = (dim + 15) // 16 * 16
        ld_h = kd + 16
        assert kd >= dim and ld_h % 16 == 0 and (ld_h * 2) % 16 == 0
        nslab = (ld_h + 63) // 64
        ksteps_last = (ld_h - (nslab - 1) * 64) // 16
        assert 1 <= ksteps_last <= 4 and (nslab - 1) * 4 + ksteps_last == ld_h // 16
