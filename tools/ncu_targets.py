"""One kernel per invocation, launched a few times on realistic operands: the target of `ncu --set full` captures of the
round-2 kernels (fp32-parity GEMM, split3, the many-models slab kernels).  python tools/ncu_targets.py <name> [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eeg_multimodal_b200 import _lib as L  # noqa: E402
from eeg_multimodal_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
bf = torch.bfloat16
name = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
g = torch.Generator(device=dev).manual_seed(0)


def planes(t):
    return ops.split3(t.contiguous(), planes=torch.empty(3, *t.shape, device=dev, dtype=bf))


if name in ("x3_fwd1", "x3_dW1", "split3"):
    B, D = 65536, 2560
    X = torch.rand(B, D, device=dev, generator=g) + 0.3 * torch.randn(B, D, device=dev, generator=g)
    W1 = (torch.rand(D, D, device=dev, generator=g) * 2 - 1) / D ** 0.5
    Xp, W1p = planes(X), planes(W1)
    Z = torch.zeros(B, D, device=dev)
    dW = torch.zeros(D, D, device=dev)
    for _ in range(iters):
        if name == "x3_fwd1":
            ops.gemm_bf16x3(Xp, W1p, Z, M=B, N=D, K=D, epi=L.EPI_ATOMIC_F32, k_slabs=4)
        elif name == "x3_dW1":
            ops.gemm_bf16x3(Xp, Xp, dW, M=D, N=D, K=B, a_mn=True, b_mn=True, epi=L.EPI_ATOMIC_F32, k_slabs=102)
        else:
            ops.split3(X, planes=Xp, act=L.ACT_RELU)
else:   # the batch-8 sweep at 128 models per GPU
    M, Bt, D, H = 128, 8, 2304, 768
    N, K = (D, D) if name.endswith("W1") else (H, D)
    dY = torch.randn(M, Bt, N, device=dev, generator=g) * (torch.rand(M, Bt, N, device=dev, generator=g) < 0.5)
    Xa = torch.relu(torch.randn(M, Bt, K, device=dev, generator=g))
    P = N * K + N
    flat, m, v = (torch.randn(M, P, device=dev, generator=g) * 0.02 for _ in range(3))
    v = v.abs()
    view = lambda t: (t[:, :N * K].view(M, N, K), t[:, N * K:])
    for i in range(iters):
        if name.startswith("wide_adam"):
            ops.linear_adam_step(dY, Xa, view(flat)[0], view(m)[0], view(v)[0], view(flat)[1], view(m)[1], view(v)[1], step=i + 1)
        else:
            ops.linear_bwd_dx(dY, view(flat)[0])
torch.cuda.synchronize()
print("ok", name)
