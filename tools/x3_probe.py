"""Accuracy and timing of the fp32-parity tensor-core GEMM (pgf_gemm_bf16x3) on one B200.

  python tools/x3_probe.py            # accuracy vs fp64 for accumulation-chain lengths, then timings at B = 65,536

Prints one JSON object per line (collected under profiles/ by the caller)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eeg_multimodal_b200 import _lib as L  # noqa: E402
from eeg_multimodal_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
bf = torch.bfloat16
torch.backends.cuda.matmul.allow_tf32 = False


def planes(t):
    return ops.split3(t.contiguous(), planes=torch.empty(3, *t.shape, device=dev, dtype=bf))


def err(a, ref):
    d = (a.double() - ref).abs()
    return dict(max_over_max=float(d.max() / ref.abs().max()), rms_over_rms=float(d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()))


def timeit(fn, iters=10):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def accuracy():
    g = torch.Generator(device=dev).manual_seed(0)
    # forward shape: activations in [0,1] + Laplace-like noise, nn.Linear weights
    M, N, K = 2048, 2560, 2560
    A = torch.rand(M, K, device=dev, generator=g) + 0.5 * torch.randn(M, K, device=dev, generator=g)
    W = (torch.rand(N, K, device=dev, generator=g) * 2 - 1) / K ** 0.5
    ref = A.double() @ W.double().T
    Ap, Wp = planes(A), planes(W)
    print(json.dumps(dict(what="fwd 2048x2560x2560", impl="cuBLAS sgemm", **err(A @ W.T, ref))))
    C = torch.empty(M, N, device=dev)
    ops.gemm_bf16x3(Ap, Wp, C, M=M, N=N, K=K, epi=L.EPI_STORE_F32)
    print(json.dumps(dict(what="fwd 2048x2560x2560", impl="x3 one chain (240 k-blocks)", **err(C, ref))))
    for s in (2, 3, 4, 5, 8, 10, 20):
        C.zero_()
        ops.gemm_bf16x3(Ap, Wp, C, M=M, N=N, K=K, epi=L.EPI_ATOMIC_F32, k_slabs=s)
        print(json.dumps(dict(what="fwd 2048x2560x2560", impl=f"x3 {s} slabs ({40 // s} k-blocks per chain)", **err(C, ref))))
    # plain bf16 GEMM of the same operands for scale
    Cb = torch.empty(M, N, device=dev)
    ops.gemm_bf16(Ap[0], Wp[0], Cb, M=M, N=N, K=K, epi=L.EPI_STORE_F32)
    print(json.dumps(dict(what="fwd 2048x2560x2560", impl="bf16 operands", **err(Cb, ref))))
    # weight-gradient shape: K = batch
    Bsz, N2, K2 = 16384, 2560, 2560
    dZ = torch.randn(Bsz, N2, device=dev, generator=g) * (torch.rand(Bsz, N2, device=dev, generator=g) < 0.5) / Bsz
    X = torch.rand(Bsz, K2, device=dev, generator=g) + 0.5 * torch.randn(Bsz, K2, device=dev, generator=g)
    ref = dZ.double().T @ X.double()
    print(json.dumps(dict(what="dW 2560x2560xK=16384", impl="cuBLAS sgemm", **err(dZ.T @ X, ref))))
    dZp, Xp = planes(dZ), planes(X)
    kb = Bsz // 64
    for per in (128, 64, 32, 16, 10, 8, 4):
        dW = torch.zeros(N2, K2, device=dev)
        ops.gemm_bf16x3(dZp, Xp, dW, M=N2, N=K2, K=Bsz, a_mn=True, b_mn=True, epi=L.EPI_ATOMIC_F32, k_slabs=kb // per)
        print(json.dumps(dict(what="dW 2560x2560xK=16384", impl=f"x3, {per} batch k-blocks per chain", **err(dW, ref))))


def timing():
    B, D, H = 65536, 2560, 768
    g = torch.Generator(device=dev).manual_seed(1)
    X = torch.rand(B, D, device=dev, generator=g)
    W1 = (torch.rand(D, D, device=dev, generator=g) * 2 - 1) / D ** 0.5
    W2 = (torch.rand(H, D, device=dev, generator=g) * 2 - 1) / D ** 0.5
    dZ2 = torch.randn(B, H, device=dev, generator=g) / B
    Xp, W1p, W2p, dZ2p = planes(X), planes(W1), planes(W2), planes(dZ2)
    Z = torch.randn(B, D, device=dev, generator=g)
    Tp = torch.empty(3, B, D, device=dev, dtype=bf)
    H2 = torch.empty(B, H, device=dev)
    b1 = torch.zeros(D, device=dev)
    rows = []
    t = timeit(lambda: ops.split3(Z, planes=Tp, act=L.ACT_RELU))
    rows.append(dict(k="split3 relu [65536,2560] -> 3 planes", ms=t, gbs=B * D * 10 / t / 1e6))
    def slabbed(ns):
        def f():
            ops.fill_zero(Z)
            ops.gemm_bf16x3(Xp, W1p, Z, M=B, N=D, K=D, epi=L.EPI_ATOMIC_F32, k_slabs=ns)
        return f

    for name, fn, flop in (
        ("x3 fwd1 M=65536 N=2560 K=2560", lambda: ops.gemm_bf16x3(Xp, W1p, Z, M=B, N=D, K=D, epi=L.EPI_BIAS_F32, bias=b1), 12.0 * B * D * D),
        ("x3 fwd1, zero-fill + 2 K slabs", slabbed(2), 12.0 * B * D * D),
        ("x3 fwd1, zero-fill + 4 K slabs", slabbed(4), 12.0 * B * D * D),
        ("x3 fwd1, zero-fill + 8 K slabs", slabbed(8), 12.0 * B * D * D),
        ("x3 fwd2 M=65536 N=768 K=2560", lambda: ops.gemm_bf16x3(Xp, W2p, H2, M=B, N=H, K=D, epi=L.EPI_STORE_F32), 12.0 * B * D * H),
        ("x3 dH1 M=65536 N=2560 K=768", lambda: ops.gemm_bf16x3(dZ2p, W2p, Z, M=B, N=D, K=H, b_mn=True, epi=L.EPI_STORE_F32), 12.0 * B * D * H),
        ("bf16 fwd1 (same kernel, one plane pair)", lambda: ops.gemm_bf16(Xp[0], W1p[0], Z, M=B, N=D, K=D, epi=L.EPI_STORE_F32), 2.0 * B * D * D),
        ("cuBLAS sgemm fwd1", lambda: torch.matmul(X, W1.T, out=Z), 2.0 * B * D * D),
    ):
        t = timeit(fn, 5)
        rows.append(dict(k=name, ms=t, tflops_bf16_equiv=flop / t / 1e9))
    dW = torch.zeros(D, D, device=dev)
    kb = B // 64
    for per in (4, 8, 10, 16, 32):
        t = timeit(lambda: ops.gemm_bf16x3(Xp, Xp, dW, M=D, N=D, K=B, a_mn=True, b_mn=True, epi=L.EPI_ATOMIC_F32, k_slabs=kb // per), 5)
        rows.append(dict(k=f"x3 dW1 M=2560 N=2560 K=65536, {per} batch k-blocks per chain", ms=t, tflops_bf16_equiv=12.0 * B * D * D / t / 1e9))
    for r in rows:
        print(json.dumps(r))


def step_timing():
    from eeg_multimodal_b200 import HeadEngine

    B, dims = 65536, (2048, 512)
    g = torch.Generator(device=dev).manual_seed(2)
    blocks = [torch.rand(B, d, device=dev, generator=g) for d in dims]
    label = (torch.rand(B, device=dev, generator=g) < 0.66).long()
    for prec in ("fp32x3", "bf16"):
        eng = HeadEngine(n_models=1, feature_dims=dims, eps=1.0, precision=prec)
        t = timeit(lambda: eng.train_step(blocks, label), 5)
        print(json.dumps(dict(k=f"train_step one model B=65536 D=2560 precision={prec}", ms=t, model_samples_per_s=B / t * 1e3)))
        del eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    accuracy()
    timing()
    step_timing()
