#!/usr/bin/env python
"""Per-kernel micro-benchmarks at the BASELINE config-4 shapes (B=65536, D=2560, H=768): CUDA-event
timings, algorithmic bytes / flops, fraction of the measured peaks.  Inputs are >= 100 MB (larger
than... or comparable to the 126 MB L2; the big ones are several times L2), `--only` filters by
name, `--iters` sets the repeat count.  Used for the ncu captures under profiles/."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eeg_multimodal_b200 import _lib as L, ops  # noqa: E402


def timeit(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def sweep_main(a):
    """Grouped fp32 small-batch kernels at the reference shapes: M models x (B=8, D=2304, H=768)."""
    dev = torch.device("cuda:0")
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks = {"hbm_gbs": 6650.0}
    M, B, D, H = a.sweep, 8, 2304, 768
    P = D * D + D + H * D + H + 2 * H + 8
    flat, m, v, gr = (torch.zeros(M, P, device=dev) for _ in range(4))
    flat.uniform_(-0.02, 0.02)
    W1, b1 = flat[:, :D * D].view(M, D, D), flat[:, D * D:D * D + D]
    o2 = D * D + D
    W2, b2 = flat[:, o2:o2 + H * D].view(M, H, D), flat[:, o2 + H * D:o2 + H * D + H]
    mv = lambda t, off, shape: t[:, off:off + shape[0] * (shape[1] if len(shape) > 1 else 1)].view(M, *shape)
    X = torch.rand(M, B, D, device=dev)
    H1 = torch.relu(torch.randn(M, B, D, device=dev))          # ReLU output: half of the entries are exactly zero
    dZ1 = torch.randn(M, B, D, device=dev) * 1e-3 * (torch.rand(M, B, D, device=dev) < 0.5)
    dZ2 = torch.randn(M, B, H, device=dev) * 1e-3
    Y1, Y2, dX = torch.empty(M, B, D, device=dev), torch.empty(M, B, H, device=dev), torch.empty(M, B, D, device=dev)
    cases = [
        ("linear_fwd_W1", lambda: ops.linear_fwd(X, W1, b1, L.ACT_RELU, out=Y1), M * D * D * 4),
        ("linear_fwd_W2", lambda: ops.linear_fwd(H1, W2, b2, L.ACT_TANH, out=Y2), M * H * D * 4),
        ("linear_dx_W1", lambda: ops.linear_bwd_dx(dZ1, W1, out=dX), M * D * D * 4),
        ("linear_dx_W2", lambda: ops.linear_bwd_dx(dZ2, W2, mask_src=H1, mask_mode=L.ACT_RELU, out=dX), M * H * D * 4),
        ("linear_dw_W1", lambda: ops.linear_bwd_dw(dZ1, X, dW=mv(gr, 0, (D, D)), db=gr[:, D * D:D * D + D]), M * D * D * 4),
        ("linear_adam_W1", lambda: ops.linear_adam_step(dZ1, X, W1, mv(m, 0, (D, D)), mv(v, 0, (D, D)), b1, m[:, D * D:D * D + D],
                                                        v[:, D * D:D * D + D], 1), M * D * D * 24),
        ("linear_adam_W2", lambda: ops.linear_adam_step(dZ2, H1, W2, mv(m, o2, (H, D)), mv(v, o2, (H, D)), b2,
                                                        m[:, o2 + H * D:o2 + H * D + H], v[:, o2 + H * D:o2 + H * D + H], 1), M * H * D * 24),
        ("adam_flat", lambda: ops.adam_step(flat, gr, m, v, 1), M * P * 28),
    ]
    for name, fn, nbytes in cases:
        if a.only and a.only not in name:
            continue
        ms = timeit(fn, a.iters)
        print(json.dumps({"kernel": name, "models": M, "ms": round(ms, 4), "GB/s": round(nbytes / ms / 1e6, 1),
                          "hbm_frac": round(nbytes / ms / 1e6 / peaks["hbm_gbs"], 3)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--sweep", type=int, default=0, help="N models: time the grouped fp32 kernels of the B=8 sweep (D=2304) instead")
    a = ap.parse_args()
    if a.sweep:
        return sweep_main(a)
    dev = torch.device("cuda:0")
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
    B, D, H = a.batch, 2560, 768
    bf = torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(0)
    x0, x1 = torch.rand(B, 2048, device=dev, generator=g), torch.rand(B, 512, device=dev, generator=g)
    DP = torch.zeros(D, device=dev)
    w, eh, deps = ops.dp_coeffs(DP, 2.718281828)
    w6, eh6, _ = ops.dp_coeffs(torch.zeros(6, D, device=dev), [1.105, 2.718, 20.09, 148.4, 2981.0, 22026.0])
    seeds6 = torch.arange(980616, 980622, dtype=torch.int64, device=dev)
    seeds1x6 = torch.full((6,), 980616, dtype=torch.int64, device=dev)
    Xh6 = torch.empty(6, B, D, dtype=bf, device=dev)
    Xh = torch.empty(B, D, dtype=bf, device=dev)
    Xf = torch.empty(B, D, device=dev)
    W1 = ((torch.rand(D, D, device=dev, generator=g) * 2 - 1) / D ** 0.5).to(bf)
    W2 = ((torch.rand(H, D, device=dev, generator=g) * 2 - 1) / D ** 0.5).to(bf)
    b1, b2 = torch.zeros(D, device=dev), torch.zeros(H, device=dev)
    H1 = torch.empty(B, D, dtype=bf, device=dev)
    H2 = torch.empty(B, H, device=dev)
    dZ2 = (torch.randn(B, H, device=dev, generator=g) * 1e-3).to(bf)
    H2.copy_(torch.tanh(torch.randn(B, H, device=dev, generator=g)))
    # realistic operand statistics (tensor-pipe power, hence the clock under the cap, depends on operand toggling)
    dZ1 = ((torch.randn(B, D, device=dev, generator=g) * 1e-3) * (torch.rand(B, D, device=dev, generator=g) < 0.5)).to(bf)
    dW1, dW2 = torch.zeros(D, D, device=dev), torch.zeros(H, D, device=dev)
    Wc, bc = torch.randn(2, H, device=dev, generator=g) * 0.03, torch.zeros(2, device=dev)
    labels = (torch.rand(B, device=dev, generator=g) < 0.66).long()
    bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (B, D // 32), dtype=torch.int32, device=dev)
    dZ1s = torch.empty(B, D, dtype=bf, device=dev)   # scratch output for the dZ1 cases (dZ1 itself stays an input)
    b1g, b2g, dDPo = torch.zeros(D, device=dev), torch.zeros(H, device=dev), torch.zeros(D, device=dev)
    P = D * D + D + H * D + H + 2 * H + 8
    p, gr, m, v = (torch.zeros(P, device=dev) for _ in range(4))
    sh = torch.zeros(P, dtype=bf, device=dev)
    ops.perturb_gate_fwd([x0, x1], w, eh, noise_mode=L.NOISE_PHILOX, seed=1, out=Xh)
    ops.gemm_bf16(Xh, W1, H1, M=B, N=D, K=D, epi=L.EPI_BIAS_RELU_BF16, bias=b1)

    # the older PriGumbel head tail (train_val.py:95-123) at the same batch: z = fc2 output [B,H] fp32
    pg_w = torch.rand(H, device=dev, generator=g) * 0.9 + 0.05
    pg_coef, pg_wloss = ops.prigumbel_coef(pg_w, exp_eps=2.718, tau=0.01, hard=False, seed=1)   # the reference's tau (train_val.py:524)
    pg_out, pg_dz, pg_dw = torch.empty(B, H, device=dev), torch.empty(B, H, device=dev), torch.empty(H, device=dev)
    pg_dout = torch.randn(B, H, device=dev, generator=g) / B

    cases = [
        ("prigumbel_fwd", lambda: ops.prigumbel_fwd(H2, pg_coef, eps=1.0, seed=1, out=pg_out), B * H * 8, 0),
        ("prigumbel_bwd", lambda: ops.prigumbel_bwd(H2, pg_coef, pg_dout, wloss=pg_wloss, exp_eps=2.718, dz=pg_dz, dw=pg_dw), B * H * 12, 0),
        ("perturb_fwd_philox_bf16", lambda: ops.perturb_gate_fwd([x0, x1], w, eh, noise_mode=L.NOISE_PHILOX, seed=1, out=Xh), B * D * 6, 0),
        # one GPU's share of the eps x seed grid: six eps values at ONE seed (noise shared inside the kernel) ...
        ("perturb_fwd_shared6_bf16", lambda: ops.perturb_gate_fwd([x0, x1], w6, eh6, noise_mode=L.NOISE_PHILOX, model_seeds=seeds1x6, out=Xh6, n_models=6), B * D * (4 + 6 * 2), 0),
        # ... and six different seeds (every model regenerates its own noise)
        ("perturb_fwd_6seeds_bf16", lambda: ops.perturb_gate_fwd([x0, x1], w6, eh6, noise_mode=L.NOISE_PHILOX, model_seeds=seeds6, out=Xh6, n_models=6), B * D * (4 + 6 * 2), 0),
        ("perturb_fwd_philox_f32", lambda: ops.perturb_gate_fwd([x0, x1], w, eh, noise_mode=L.NOISE_PHILOX, seed=1, out=Xf), B * D * 8, 0),
        ("perturb_fwd_nonoise_bf16", lambda: ops.perturb_gate_fwd([x0, x1], None, None, noise_mode=L.NOISE_NONE, out=Xh), B * D * 6, 0),
        ("perturb_fwd_philox_gate_bf16", lambda: ops.perturb_gate_fwd([x0, x1], w, eh, noise_mode=L.NOISE_PHILOX, seed=1, out=Xh, want_gate=True), B * D * 6, 0),
        ("perturb_bwd_dp_f32", lambda: ops.perturb_gate_bwd_dp(Xf, deps, noise_mode=L.NOISE_PHILOX, seed=1), B * D * 4, 0),
        ("perturb_bwd_dp_bf16", lambda: ops.perturb_gate_bwd_dp(Xh, deps, noise_mode=L.NOISE_PHILOX, seed=1), B * D * 2, 0),
        ("gemm_fwd1", lambda: ops.gemm_bf16(Xh, W1, H1, M=B, N=D, K=D, epi=L.EPI_BIAS_RELU_BF16, bias=b1), 0, 2 * B * D * D),
        ("gemm_fwd2", lambda: ops.gemm_bf16(H1, W2, H2, M=B, N=H, K=D, epi=L.EPI_BIAS_TANH_F32, bias=b2), 0, 2 * B * H * D),
        ("gemm_dX", lambda: ops.gemm_bf16(dZ1, W1, Xf, M=B, N=D, K=D, b_mn=True, epi=L.EPI_STORE_F32), 0, 2 * B * D * D),
        ("gemm_dX_dDP_fused", lambda: ops.gemm_bf16_ddp(dZ1, W1, M=B, N=D, K=D, b_mn=True, seed=1, offset=0, row0=0, deps_dDP=deps, out=dDPo), 0, 2 * B * D * D),
        ("gemm_fwd1_bits", lambda: ops.gemm_bf16(Xh, W1, H1, M=B, N=D, K=D, epi=L.EPI_BIAS_RELU_BF16, bias=b1, aux=bits), 0, 2 * B * D * D),
        ("gemm_dZ1_bits", lambda: ops.gemm_bf16(dZ2, W2, dZ1s, M=B, N=D, K=H, b_mn=True, epi=L.EPI_BITMASK_BF16, aux=bits), 0, 2 * B * H * D),
        ("gemm_dZ1_bits_db1", lambda: ops.gemm_bf16(dZ2, W2, dZ1s, M=B, N=D, K=H, b_mn=True, epi=L.EPI_BITMASK_BF16, aux=bits, colsum_out=b1g), 0, 2 * B * H * D),
        ("gemm_dW1", lambda: ops.gemm_bf16(dZ1, Xh, dW1, M=D, N=D, K=B, a_mn=True, b_mn=True, epi=L.EPI_ATOMIC_F32, stream_k=True), 0, 2 * B * D * D),
        ("gemm_dW2", lambda: ops.gemm_bf16(dZ2, H1, dW2, M=H, N=D, K=B, a_mn=True, b_mn=True, epi=L.EPI_ATOMIC_F32, stream_k=True), 0, 2 * B * H * D),
        ("cls_ce_pass1_dz", lambda: ops.cls_ce(H2, Wc, bc, labels, loss_scale=1 / B, grad_scale=1 / B, backward=True, dz=dZ2, dz_dtype=bf, want_dw=False), B * H * 6, 0),
        ("cls_ce_pass2_dz_dw_db2", lambda: ops.cls_ce(H2, Wc, bc, labels, loss_scale=1 / B, grad_scale=1 / B, backward=True, dz=dZ2, dz_dtype=bf, dz_colsum=b2g), B * H * 6, 0),
        ("cls_ce_eval", lambda: ops.cls_ce(H2, Wc, bc, labels, loss_scale=1 / B, grad_scale=1 / B, backward=False), B * H * 4, 0),
        ("colsum_bf16_D", lambda: ops.colsum(dZ1), B * D * 2, 0),
        ("adam", lambda: ops.adam_step(p, gr, m, v, 1, bf16_shadow=sh), P * 30, 0),
    ]
    for name, fn, nbytes, flops in cases:
        if a.only and a.only not in name:
            continue
        ms = timeit(fn, a.iters)
        line = {"kernel": name, "ms": round(ms, 4)}
        if nbytes:
            line["GB/s"] = round(nbytes / ms / 1e6, 1)
            line["hbm_frac"] = round(nbytes / ms / 1e6 / peaks["hbm_gbs"], 3)
        if flops:
            line["TFLOP/s"] = round(flops / ms / 1e9, 1)
            line["tensor_frac_burst"] = round(flops / ms / 1e9 / peaks["bf16_tflops"], 3)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
