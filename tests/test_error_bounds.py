"""CPU check of the error budgets the batched path's guard rests on (vrod_b200/csrc/knn_batched.cu,
launch_batched_search / batched_finish_kernel): for operands rounded the way the tensor cores see them, the dot
product stays within eps_dot * ||x|| * ||q|| of the exact one --

    bf16 mirror mode   operands rounded to nearest bf16 (8 significant bits)   eps_dot = 1.01 * 2^-7 + ld_h * 2^-22
    tf32 mode          operands truncated to 10 mantissa bits                  eps_dot = 1.01 * 2^-9 + ld   * 2^-22

(the second term pays for the f32 accumulation).  Random data cannot reach a worst-case bound, so adversarial
operands that sit just below a rounding boundary in every component are checked as well, and must come close to it.
The emulation accumulates in f32 in torch's order, the hardware in its own: the bound is order-independent."""
import numpy as np
import pytest
import torch


def to_bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def to_tf32_truncated(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def eps_dot(kind, d):
    if kind == "bf16":
        ld_h = (d + 15) // 16 * 16 + 16
        return 1.01 * 2.0 ** -7 + ld_h * 2.0 ** -22
    ld = (d + 3) // 4 * 4
    return 1.01 * 2.0 ** -9 + ld * 2.0 ** -22


ROUND = {"bf16": to_bf16, "tf32": to_tf32_truncated}


@pytest.mark.parametrize("kind", ["bf16", "tf32"])
@pytest.mark.parametrize("d", [8, 64, 128, 768, 1536])
def test_random_operands_stay_within_the_budget(kind, d):
    g = torch.Generator().manual_seed(d)
    for scale in (1.0, 1e-3, 1e3):
        X = (torch.rand(4096, d, generator=g) * 2 - 1) * scale
        Q = torch.randn(64, d, generator=g)
        approx = ROUND[kind](X) @ ROUND[kind](Q).T                      # f32 accumulate
        exact = X.double() @ Q.double().T
        bound = eps_dot(kind, d) * X.double().norm(dim=1)[:, None] * Q.double().norm(dim=1)[None, :]
        ratio = ((approx.double() - exact).abs() / bound).max().item()
        assert ratio < 1.0, (kind, d, scale, ratio)


@pytest.mark.parametrize("kind", ["bf16", "tf32"])
def test_adversarial_operands_approach_but_respect_the_budget(kind):
    d = 128
    # every component loses (almost) the largest relative amount the rounding can take away, with equal signs, and
    # x is parallel to q (Cauchy-Schwarz is tight): the error is as large as this operand format allows
    if kind == "bf16":
        v = np.float32(1.0) + np.float32(2.0 ** -8) - np.float32(2.0 ** -20)    # just below the tie: rounds down to 1
    else:
        v = np.float32(1.0) + np.float32(2.0 ** -10) - np.float32(2.0 ** -22)   # truncation drops almost one ulp
    X = torch.full((1, d), float(v))
    Q = torch.full((1, d), float(v))
    approx = (ROUND[kind](X) @ ROUND[kind](Q).T).double()
    exact = X.double() @ Q.double().T
    bound = eps_dot(kind, d) * X.double().norm() * Q.double().norm()
    ratio = ((approx - exact).abs() / bound).item()
    assert 0.9 < ratio < 1.0, (kind, ratio)
