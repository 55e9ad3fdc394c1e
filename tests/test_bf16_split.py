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


def tiled_offset(r, chunk, R):
    """knn_batched.cu tiled_offset: byte offset of 8-column chunk `chunk` of row r inside a tile of R rows."""
    return (chunk >> 1) * R * 32 + (r >> 3) * 256 + (chunk & 1) * 128 + (r & 7) * 16


def test_mirror_geometry_matches_the_library():
    """kd = dim rounded up to 16 data columns + 16 aux columns = T K steps of 16 columns.  The mirrors are TILED: a (tile of
    R rows, K step) block is R * 32 contiguous bytes in the no-swizzle K-major UMMA layout -- 8-row x 16-byte core
    matrices of 128 contiguous bytes, the two 8-column halves of a K step 128 bytes apart (the descriptor's leading byte
    offset), consecutive 8-row groups 256 bytes apart (its stride byte offset) -- so a whole tile is ONE contiguous piece
    that a single bulk copy moves and the MMA descriptor only advances by R * 32 bytes per K step."""
    for dim in (1, 8, 15, 16, 17, 64, 100, 128, 130, 768, 1536):
        kd = (dim + 15) // 16 * 16
        ld_h = kd + 16
        assert kd >= dim and ld_h % 16 == 0
        T = ld_h // 16
        for R in (128, 256):                      # row tiles (BM) and query groups (BN)
            offs = {tiled_offset(r, ch, R) for r in range(R) for ch in range(2 * T)}
            assert len(offs) == R * 2 * T, "two (row, chunk) pairs share a slot"
            assert offs == set(range(0, T * R * 32, 16)), "the tile is not one dense piece of T * R * 32 bytes"
        R = 128
        for step in range(T):                     # every K step is its own contiguous R * 32-byte block
            block = {tiled_offset(r, 2 * step + h, R) for r in range(R) for h in (0, 1)}
            assert min(block) == step * R * 32 and max(block) == (step + 1) * R * 32 - 16
        for g in range(R // 8):                   # core matrices: 8 rows x 16 bytes, contiguous; LBO 128, SBO 256
            base = tiled_offset(8 * g, 0, R)
            assert [tiled_offset(8 * g + i, 0, R) - base for i in range(8)] == [16 * i for i in range(8)]
            assert tiled_offset(8 * g, 1, R) - base == 128
            if g + 1 < R // 8:
                assert tiled_offset(8 * (g + 1), 0, R) - base == 256
    # shared-memory budget at dim 128: 9 K steps -> 72 KB of queries + 4 stages of 36 KB + the 11200-byte control block
    # (21 mbarriers, flags, 256 thresholds and counters, 16 warps x 4 stash slots of 140 bytes) fit 227 KB with 64 bytes left
    T = (128 + 16) // 16
    ctl = 21 * 8 + 16 + 8 + 256 * 4 + 256 * 4 + 16 * 4 * 140
    assert ctl == 11200 and 0 <= 227 * 1024 - (T * 256 * 32 + 4 * T * 128 * 32 + ctl) < 128
