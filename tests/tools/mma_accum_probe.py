"""How large is the f32 accumulation error of a bf16 tensor-core contraction on this GPU?  cuBLAS bf16 x bf16 -> f32
(the same tcgen05 data path the batched kernel uses, library kernel) against the f64 product of the SAME rounded
operands, relative to ||x|| ||q||, next to the budget the guard reserves for it (ld * 2^-22)."""
import torch
torch.manual_seed(0)
dev = "cuda"
for K in (144, 400, 784, 1552):
    X = (torch.rand(8192, K, device=dev) * 2 - 1).to(torch.bfloat16)
    Q = torch.randn(1024, K, device=dev).to(torch.bfloat16)
    try:
        S = torch.mm(X, Q.T, out_dtype=torch.float32)
        how = "mm(out_dtype=f32)"
    except Exception as e:                                   # older torch: no f32 output from a bf16 GEMM
        print("no f32-output bf16 GEMM available:", type(e).__name__, str(e)[:80]); break
    E = X.double() @ Q.double().T
    nx = X.double().norm(dim=1)[:, None]; nq = Q.double().norm(dim=1)[None, :]
    rel = ((S.double() - E).abs() / (nx * nq)).max().item()
    # same positive operands: no cancellation, the sum is as large as ||x|| ||q|| allows
    Xp, Qp = X.abs(), Q.abs()
    Sp = torch.mm(Xp, Qp.T, out_dtype=torch.float32)
    Ep = Xp.double() @ Qp.double().T
    relp = ((Sp.double() - Ep).abs() / (nx * nq)).max().item()
    print(f"K={K:5d} {how}: max |err| / (||x|| ||q||) = {rel:.3e} (signed data), {relp:.3e} (positive data); "
          f"budget ld*2^-22 = {K * 2.0 ** -22:.3e}; 2^-24 = {2.0 ** -24:.3e}")
