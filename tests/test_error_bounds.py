"""CPU check of the error budgets the batched path's guard rests on (vrod_b200/csrc/knn_batched.cu,
launch_batched_search / batched_finish_kernel): for operands rounded the way the tensor cores see them, the dot
product stays within eps_dot * ||x|| * ||q|| of the exact one --

    bf16 mirror mode   operands rounded to nearest bf16 (8 significant bits)   eps_dot = 1.01 * 2^-7 + ld_h * 2^-22
                       (the worst case; since round 2 the guard charges the operand roundings AS MEASURED while the
                       mirrors are built -- measured_operand_error below -- plus ld_h * 2^-22 for the accumulation)
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


def measured_operand_error(Xs, Q):
    """knn_batched.cu operand_error: max||x~|| * ||q~ - q|| + max||x~ - x|| * ||q|| with the maxima over the mirrored rows
    (build_mirror_kernel measures them while it rounds; f32 sums, inflated like the kernel inflates them)."""
    Xh, Qh = to_bf16(Xs), to_bf16(Q)
    err = ((Xh - Xs) ** 2).sum(1).max().sqrt() * 1.00001
    length = (Xh ** 2).sum(1).max().sqrt() * 1.00001
    dq = ((Qh - Q) ** 2).sum(1).sqrt()
    nq = (Q.double() ** 2).sum(1).sqrt().float() * 1.0000002
    return ((length * dq + err * nq) * 1.00002).double() * 1.000001


@pytest.mark.parametrize("d", [8, 128, 1536])
def test_measured_operand_error_bounds_the_bf16_dot_product(d):
    """The bf16 mode's guard charges the operand roundings as measured, not at their worst case: the bound must hold for
    every row/query pair -- and stay well below the worst-case 2^-7 ||x|| ||q|| on ordinary data (half-ulp errors are
    uniform, not maximal: ~0.42 of the worst case), or nothing was gained."""
    g = torch.Generator().manual_seed(100 + d)
    for scale in (1.0, 1e-3, 1e3):
        X = (torch.rand(4096, d, generator=g) * 2 - 1) * scale
        Q = torch.randn(64, d, generator=g)
        approx = (to_bf16(X).double() @ to_bf16(Q).double().T)          # exact products of the rounded operands
        exact = X.double() @ Q.double().T
        bound = measured_operand_error(X, Q)[None, :]
        assert ((approx - exact).abs() / bound).max().item() < 1.0
        worst_case = 2.0 ** -7 * X.double().norm(dim=1).max() * Q.double().norm(dim=1)
        assert (bound[0] / worst_case).max().item() < (0.5 if d >= 128 else 0.7), "the measured bound should be ~2x tighter on random data"


def test_measured_operand_error_reaches_the_worst_case_on_adversarial_operands():
    d = 128
    v = np.float32(1.0) + np.float32(2.0 ** -8) - np.float32(2.0 ** -20)        # just below the tie: rounds down to 1
    X = torch.full((1, d), float(v))
    Q = torch.full((1, d), float(v))
    err = ((to_bf16(X).double() @ to_bf16(Q).double().T) - X.double() @ Q.double().T).abs().item()
    bound = measured_operand_error(X, Q).item()
    assert 0.95 < err / bound < 1.0, err / bound


def _split3_sum(v):
    a = to_bf16(v)
    b = to_bf16(v - a)
    c = to_bf16(v - a - b)
    return (a + b) + c


@pytest.mark.parametrize("metric", ["euclidean", "cosine"])
@pytest.mark.parametrize("d", [24, 128, 384])
def test_folded_surrogate_stays_within_the_guard_budget(metric, d):
    """The whole budget of batched_finish_kernel's guard in the bf16 mode: the surrogate the tile kernel computes
    from the folded contraction, v' = thr - D with D = dot~ + thr - ||x||^2/2 (cosine: D = dot^ + thr), stays within
    E of the exact surrogate (||x||^2/2 - dot, cosine -dot/||x||) for every row and every threshold between the
    smallest surrogate and the start cap."""
    g = torch.Generator().manual_seed(d)
    n, b = 2048, 32
    ld_h = (d + 15) // 16 * 16 + 16
    eps = ld_h * 2.0 ** -22          # the accumulation's share of eps_dot; the operand roundings are charged as measured
    acc_eps = (ld_h // 16 + 2) * 2.0 ** -23
    for scale in (1.0, 30.0, 0.01):
        X = torch.randn(n, d, generator=g) * scale * (10.0 ** (torch.rand(n, 1, generator=g) * 2 - 1))
        Q = torch.randn(b, d, generator=g) * scale
        nx = (X.double() ** 2).sum(1)
        nq = (Q.double() ** 2).sum(1).sqrt()
        M = nx.max().float().double()
        if metric == "euclidean":
            hx = (0.5 * nx).float()                                             # f32 of the canonical f64 norm, halved
            cap = ((0.5 * M + M.sqrt() * nq * 1.0001) * 1.02 + 1e-30).float()
            exact = 0.5 * nx[:, None] - X.double() @ Q.double().T
            dot = to_bf16(X) @ to_bf16(Q).T
            Ed = measured_operand_error(X, Q)[None, :]
        else:
            inv = (1.0 / nx.sqrt()).float()
            cap = (nq * 1.0001 * 1.02 + 1e-30).float()
            exact = -(X.double() @ Q.double().T) / nx.sqrt()[:, None]
            dot = to_bf16(X * inv[:, None]) @ to_bf16(Q).T
            Ed = measured_operand_error(X * inv[:, None], Q)[None, :]
        for frac in (1.0, 0.5, 0.0):                                            # thresholds from the cap down to the best row
            lo = exact.min(dim=0).values.float()
            thr = _split3_sum(lo + (cap - lo) * frac)[None, :]                  # what the three aux parts stand for
            D = (dot + thr) - hx[:, None] if metric == "euclidean" else dot + thr   # f32, one of the possible orders
            v = thr - D
            u = v.double().abs()
            if metric == "euclidean":
                E = Ed + (eps + 2.4e-7) * M.sqrt() * nq[None, :] + 2.4e-7 * (0.5 * M + u) \
                    + acc_eps * (M.sqrt() * nq[None, :] + 0.5 * M + cap.double()[None, :] + u)
            else:
                E = Ed + (eps + 3.0e-7) * nq[None, :] + 1.2e-7 * u + acc_eps * (1.01 * nq[None, :] + cap.double()[None, :] + u)
            ratio = ((v.double() - exact).abs() / E).max().item()
            assert ratio < 1.0, (metric, d, scale, frac, ratio)


@pytest.mark.parametrize("metric", ["euclidean", "cosine"])
@pytest.mark.parametrize("d", [24, 128, 768])
def test_tf32_surrogate_stays_within_the_guard_budget(metric, d):
    """The same for the tf32 mode, where the epilogue forms v = ||x||^2/2 - dot~ (cosine: -dot~ * 1/||x||) itself."""
    g = torch.Generator().manual_seed(1000 + d)
    n, b = 2048, 32
    eps = eps_dot("tf32", d)
    for scale in (1.0, 30.0, 0.01):
        X = torch.randn(n, d, generator=g) * scale * (10.0 ** (torch.rand(n, 1, generator=g) * 2 - 1))
        Q = torch.randn(b, d, generator=g) * scale
        nx = (X.double() ** 2).sum(1)
        nq = (Q.double() ** 2).sum(1).sqrt()
        M = nx.max().float().double()
        dot = to_tf32_truncated(X) @ to_tf32_truncated(Q).T
        if metric == "euclidean":
            v = (0.5 * nx.float())[:, None] - dot
            exact = 0.5 * nx[:, None] - X.double() @ Q.double().T
            E = (eps + 2.4e-7) * M.sqrt() * nq[None, :] + 2.4e-7 * (0.5 * M + v.double().abs())
        else:
            v = -(dot * (1.0 / nx.sqrt()).float()[:, None])
            exact = -(X.double() @ Q.double().T) / nx.sqrt()[:, None]
            E = (eps + 3.0e-7) * nq[None, :] + 1.2e-7 * v.double().abs()
        ratio = ((v.double() - exact).abs() / E).max().item()
        assert ratio < 1.0, (metric, d, scale, ratio)


# fast_scan_kernel variants (vrod_b200/csrc/knn_scan.cu, kVariants): dim -> (lanes per row, float4 chunks per lane)
SCAN_VARIANTS = {32: (8, 1), 64: (16, 1), 128: (32, 1), 256: (32, 2), 384: (32, 3), 512: (32, 4), 768: (32, 6),
                 1024: (32, 8), 1536: (32, 12)}


def scan_eps(ld, cosine):
    """make_scan_plan's budget of the f32 pass (relative for L2, absolute on the cosine similarity)."""
    u = 2.0 ** -24
    eps = 1.5 * ((ld + 31) // 32 * 4 + 12) * u + ld * 2.0 ** -50
    return eps + 4 * u if cosine else eps


def test_scan_budget_covers_the_reduction_depth():
    """Every product enters the row sum through: 1 subtraction (L2), 1 fma, the rest of its lane's fma chain
    (4 * chunks per lane), log2(lanes) butterfly adds -- one rounding each, all relative to partial sums that are
    bounded by SUM |terms|.  (1 + u)^depth - 1 must fit the budget with room for the cosine path's inverse norm."""
    u = 2.0 ** -24
    for d, (lpr, ch) in SCAN_VARIANTS.items():
        depth = 1 + 4 * ch + int(np.log2(lpr))
        assert (1 + u) ** depth - 1 < scan_eps(d, False) / 1.4, (d, depth)
    for d in (4, 20, 100, 200, 1000, 4096):                 # generic variant: 32 lanes, run-time chunk loop
        ld = (d + 3) // 4 * 4
        depth = 1 + 4 * ((ld // 4 + 31) // 32) + 5
        assert (1 + u) ** depth - 1 < scan_eps(ld, False) / 1.4, (d, depth)


@pytest.mark.parametrize("cosine", [False, True], ids=["euclidean", "cosine"])
@pytest.mark.parametrize("d", [32, 64, 128, 768, 1536])
def test_scan_order_emulation_stays_within_the_budget(d, cosine):
    """f32 emulation of the kernel's summation order (lane chains, then the butterfly; numpy has no fma, so every
    term carries one rounding MORE than on the GPU) against f64."""
    lpr, ch = SCAN_VARIANTS[d]
    rng = np.random.default_rng(d)
    n = 4000
    X = rng.uniform(-1, 1, (n, d)).astype(np.float32)
    q = rng.uniform(-1, 1, d).astype(np.float32)
    Xr = X.reshape(n, ch, lpr, 4)                            # element (chunk c of lane l, component e) = 4 * (l + lpr * c) + e
    qr = q.reshape(ch, lpr, 4)
    acc = np.zeros((n, lpr), dtype=np.float32)
    for c in range(ch):
        for e in range(4):
            if cosine:
                term = Xr[:, c, :, e] * qr[c, :, e]
            else:
                a = Xr[:, c, :, e] - qr[c, :, e]
                term = a * a
            acc = (acc + term).astype(np.float32)
    h = lpr // 2
    while h >= 1:                                            # butterfly: lane l adds lane l ^ h
        acc = (acc[:, :h] + acc[:, h:2 * h]).astype(np.float32)
        h //= 2
    got = acc[:, 0].astype(np.float64)
    X64, q64 = X.astype(np.float64), q.astype(np.float64)
    if cosine:
        err = np.abs(got - X64 @ q64) / (np.linalg.norm(X64, axis=1) * np.linalg.norm(q64))
    else:
        exact = ((X64 - q64) ** 2).sum(1)
        err = np.abs(got - exact) / exact
    assert err.max() < scan_eps(d, cosine), (d, cosine, err.max(), scan_eps(d, cosine))
