"""GPU parity tests, kernel by kernel, through the C ABI (ops.* are 1:1 ctypes calls).

Checker = oracle/ (CPU restatement of the reference) on the same seeded inputs.  Bars
(BASELINE.json north_star): fp32 outputs and gradients within 1e-5 relative, bf16-GEMM mode
within 2e-2, gate indices and argmax predictions bit-exact.
"""
import numpy as np
import pytest
import torch

from conftest import golden_dp
from oracle import head_oracle as ho
from oracle import philox_ref

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from eeg_multimodal_b200 import _lib

    _lib.load()
    return torch.device("cuda:0")


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def make_blocks(B, dims, seed, dist="uniform"):
    g = torch.Generator().manual_seed(seed)
    if dist == "uniform":
        return [torch.rand(B, d, generator=g) for d in dims]
    return [torch.randn(B, d, generator=g) * 0.5 for d in dims]


def oracle_perturb(blocks, DP, eps, lap, gum, hard, fixed=True, tau=1.0):
    p = ho.HeadParams(*(torch.zeros(1),) * 6, DP.view(1, -1))
    eps_t = ho.eps_tensor(eps)
    feature = ho.minmax_normalise(torch.cat(blocks, 1))
    w = torch.sigmoid(p.DP)
    eh = ho.eps_hat_of(w, eps_t, fixed)
    fp = feature + lap * eh
    if gum is None:
        return fp, None, feature
    mask, idx = ho.gumbel_mask(w, fp.shape[0], gum, hard, tau)
    return (fp * mask).sum(0), idx, feature


# ------------------------------------------------------------------------------------------------
# kernel (a) forward
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B", [1, 2, 8, 601])
@pytest.mark.parametrize("dims", [(768, 768, 768), (2048, 512), (64,), (1024, 2048, 1024)])
@pytest.mark.parametrize("hard", [True, False])
def test_perturb_gate_injected_matches_oracle(dev, golden, B, dims, hard):
    from eeg_multimodal_b200 import _lib as L, ops

    D = sum(dims)
    blocks = make_blocks(B, dims, 100 + B, "normal")
    lap, gum = ho.replay_reference_draws(7 + B, B, D)
    DP = torch.from_numpy(golden_dp(golden, "wvalues"))[:D] if D <= 2304 else torch.randn(D, generator=torch.Generator().manual_seed(1)) * 0.1
    for eps in (0.1, 1.0, 8.0):
        ref, ref_idx, ref_norm = oracle_perturb(blocks, DP, eps, lap, gum, hard)
        w, eh, _ = ops.dp_coeffs(DP.to(dev), ho.exp_eps_f32(eps))
        out, idx, rmin, rmax = ops.perturb_gate_fwd([b.to(dev) for b in blocks], w, eh, noise_mode=L.NOISE_INJECTED,
                                                    lap=lap.to(dev), gum=gum.to(dev), hard=hard, want_gate=True,
                                                    want_gate_idx=True, want_minmax=True)
        torch.cuda.synchronize()
        assert rel_err(out, ref) < RTOL, (eps, rel_err(out, ref))
        cat = torch.cat(blocks, 1)
        assert torch.equal(rmin.cpu(), cat.min(1)[0]) and torch.equal(rmax.cpu(), cat.max(1)[0])
        # gate index: bit-exact, except where the two logits are within a few ulp (sigmoid/exp ulp differences)
        wcpu = torch.sigmoid(DP)
        z0, z1 = wcpu + gum[0], (1 - wcpu) + gum[1]
        near_tie = (z0 - z1).abs() < 1e-5
        mism = (idx.cpu().long() != ref_idx) & ~near_tie
        assert int(mism.sum()) == 0
        assert int(near_tie.sum()) < 1e-3 * near_tie.numel() + 2


def test_perturb_gate_hard_is_identity_on_device(dev):
    """models.py:77-79 known answer: with hard=True the gated value equals the perturbed value bit for bit."""
    from eeg_multimodal_b200 import _lib as L, ops

    B, dims = 8, (768, 768, 768)
    blocks = [b.to(dev) for b in make_blocks(B, dims, 3)]
    lap, gum = ho.replay_reference_draws(5, B, sum(dims))
    w, eh, _ = ops.dp_coeffs(torch.zeros(sum(dims), device=dev), ho.exp_eps_f32(1.0))
    a, _, _, _ = ops.perturb_gate_fwd(blocks, w, eh, noise_mode=L.NOISE_INJECTED, lap=lap.to(dev), gum=gum.to(dev), hard=True, want_gate=True)
    b, _, _, _ = ops.perturb_gate_fwd(blocks, w, eh, noise_mode=L.NOISE_INJECTED, lap=lap.to(dev), gum=None, hard=True, want_gate=False)
    assert torch.equal(a, b)
    # DP = 0 -> w = 0.5 exactly: gate index must be bit-exact with the oracle, no tolerance
    _, ref_idx, _ = oracle_perturb([x.cpu() for x in blocks], torch.zeros(sum(dims)), 1.0, lap, gum, True)
    _, idx, _, _ = ops.perturb_gate_fwd(blocks, w, eh, noise_mode=L.NOISE_INJECTED, lap=lap.to(dev), gum=gum.to(dev), hard=True,
                                        want_gate=True, want_gate_idx=True)
    assert torch.equal(idx.cpu().long(), ref_idx)


@pytest.mark.parametrize("dims", [(768, 768, 768), (2048, 512)])
def test_perturb_gate_philox_matches_oracle_noise(dev, dims):
    """Philox mode == oracle fed with the numpy restatement of the same counters."""
    from eeg_multimodal_b200 import _lib as L, ops

    B, D, seed, offset, row0 = 37, sum(dims), 980616, 5, 1000
    blocks = make_blocks(B, dims, 11)
    DP = torch.randn(D, generator=torch.Generator().manual_seed(2)) * 0.2
    lap = torch.from_numpy(philox_ref.laplace(seed, offset, row0, B, D))
    gum = torch.from_numpy(philox_ref.gumbel(seed, offset, row0, B, D))
    ref, ref_idx, _ = oracle_perturb(blocks, DP, 1.0, lap, gum, True)
    w, eh, _ = ops.dp_coeffs(DP.to(dev), ho.exp_eps_f32(1.0))
    dblocks = [b.to(dev) for b in blocks]
    out, idx, _, _ = ops.perturb_gate_fwd(dblocks, w, eh, noise_mode=L.NOISE_PHILOX, seed=seed, offset=offset, row0=row0,
                                          hard=True, want_gate=True, want_gate_idx=True)
    assert rel_err(out, ref) < RTOL
    wcpu = torch.sigmoid(DP)
    near_tie = ((wcpu + gum[0]) - ((1 - wcpu) + gum[1])).abs() < 1e-4
    assert int(((idx.cpu().long() != ref_idx) & ~near_tie).sum()) == 0
    # partition invariance: the second half computed alone with row0 shifted is bit-identical
    h = B // 2
    part, _, _, _ = ops.perturb_gate_fwd([b[h:].contiguous() for b in dblocks], w, eh, noise_mode=L.NOISE_PHILOX, seed=seed,
                                         offset=offset, row0=row0 + h, hard=True)
    assert torch.equal(part, out[h:])
    # a different offset gives fresh noise; bf16 output is the rounded fp32 output
    other, _, _, _ = ops.perturb_gate_fwd(dblocks, w, eh, noise_mode=L.NOISE_PHILOX, seed=seed, offset=offset + 1, row0=row0, hard=True)
    assert not torch.equal(other, out)
    ob, _, _, _ = ops.perturb_gate_fwd(dblocks, w, eh, noise_mode=L.NOISE_PHILOX, seed=seed, offset=offset, row0=row0, hard=True,
                                       out_dtype=torch.bfloat16)
    assert torch.equal(ob, out.to(torch.bfloat16))


def test_perturb_gate_edge_cases(dev):
    from eeg_multimodal_b200 import _lib as L, ops

    dims = (64, 64)
    w, eh, _ = ops.dp_coeffs(torch.zeros(128, device=dev), ho.exp_eps_f32(1.0))
    x = [torch.rand(4, 64, device=dev), torch.rand(4, 64, device=dev)]
    x[0][1] = 0.25
    x[1][1] = 0.25            # constant row: max == min -> NaN like the reference (no epsilon guard)
    x[0][2, 5] = float("nan")  # NaN input poisons its row only
    out, _, _, _ = ops.perturb_gate_fwd(x, w, eh, noise_mode=L.NOISE_PHILOX, seed=1)
    assert torch.isnan(out[1]).all() and torch.isnan(out[2]).all()
    assert torch.isfinite(out[0]).all() and torch.isfinite(out[3]).all()
    # non-private path: range exactly [0, 1]
    n, _, _, _ = ops.perturb_gate_fwd([x[0][[0, 3]].contiguous(), x[1][[0, 3]].contiguous()], None, None, noise_mode=L.NOISE_NONE)
    assert float(n.min()) == 0.0 and float(n.max()) == 1.0
    ref = ho.minmax_normalise(torch.cat([x[0][[0, 3]].cpu(), x[1][[0, 3]].cpu()], 1))
    assert torch.equal(n.cpu(), ref)
    # empty batch is a no-op; ragged widths are rejected
    e, _, _, _ = ops.perturb_gate_fwd([torch.empty(0, 64, device=dev)], w[:64].contiguous(), eh[:64].contiguous(), noise_mode=L.NOISE_PHILOX)
    assert e.shape == (0, 64)
    with pytest.raises(RuntimeError, match="multiples of 4"):
        ops.perturb_gate_fwd([torch.rand(2, 66, device=dev)], w, eh, noise_mode=L.NOISE_PHILOX)
    with pytest.raises(RuntimeError, match="4096"):
        ops.perturb_gate_fwd([torch.rand(2, 4100, device=dev)], w, eh, noise_mode=L.NOISE_NONE)


# ------------------------------------------------------------------------------------------------
# kernel (a) backward
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fixed", [True, False])
@pytest.mark.parametrize("B", [1, 8, 601])
def test_dDP_matches_autograd(dev, golden, B, fixed):
    from eeg_multimodal_b200 import _lib as L, ops

    D = 2304
    blocks = make_blocks(B, (768, 768, 768), 5)
    lap, gum = ho.replay_reference_draws(3, B, D)
    dF = torch.randn(B, D, generator=torch.Generator().manual_seed(9)) * 1e-3
    for eps in (0.1, 1.0, 8.0):
        DP = torch.from_numpy(golden_dp(golden, "wvalues")).clone().requires_grad_(True)
        w = torch.sigmoid(DP.view(1, -1))
        fp = ho.minmax_normalise(torch.cat(blocks, 1)) + lap * ho.eps_hat_of(w, ho.eps_tensor(eps), fixed)
        (fp * dF).sum().backward()
        _, _, deps = ops.dp_coeffs(DP.detach().to(dev), ho.exp_eps_f32(eps), fixed)
        got = ops.perturb_gate_bwd_dp(dF.to(dev), deps, noise_mode=L.NOISE_INJECTED, lap=lap.to(dev))
        assert rel_err(got, DP.grad) < RTOL
        got16 = ops.perturb_gate_bwd_dp(dF.to(dev).to(torch.bfloat16), deps, noise_mode=L.NOISE_INJECTED, lap=lap.to(dev))
        assert rel_err(got16, DP.grad) < 2e-2


def test_dDP_philox_regenerates_forward_noise(dev):
    from eeg_multimodal_b200 import _lib as L, ops

    B, D, seed, offset, row0 = 200, 2560, 42, 9, 77
    dF = torch.randn(B, D, generator=torch.Generator().manual_seed(1)).to(dev)
    _, _, deps = ops.dp_coeffs(torch.zeros(D, device=dev), ho.exp_eps_f32(1.0))
    lap = torch.from_numpy(philox_ref.laplace(seed, offset, row0, B, D)).to(dev)
    a = ops.perturb_gate_bwd_dp(dF, deps, noise_mode=L.NOISE_PHILOX, seed=seed, offset=offset, row0=row0)
    b = ops.perturb_gate_bwd_dp(dF, deps, noise_mode=L.NOISE_INJECTED, lap=lap)
    assert rel_err(a, b) < RTOL
    ref = (dF.double() * lap.double()).sum(0) * deps.double()
    assert rel_err(a, ref) < RTOL


def test_minmax_norm_bwd_matches_autograd(dev):
    from eeg_multimodal_b200 import ops

    for dims in ((768, 768, 768), (2048, 512)):
        blocks = [b.requires_grad_(True) for b in make_blocks(9, dims, 21, "normal")]
        dn = torch.randn(9, sum(dims), generator=torch.Generator().manual_seed(4))
        (ho.minmax_normalise(torch.cat(blocks, 1)) * dn).sum().backward()
        got = ops.minmax_norm_bwd([b.detach().to(dev) for b in blocks], dn.to(dev))
        for g, b in zip(got, blocks):
            assert rel_err(g, b.grad) < 5e-5


# ------------------------------------------------------------------------------------------------
# kernel (b) fp32 CUDA-core path (grouped)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_models,B,N,K", [(1, 8, 2304, 2304), (3, 8, 768, 2304), (2, 5, 2, 768), (1, 19, 260, 512), (48, 8, 768, 768),
                                            (1, 4099, 2, 768), (2, 1000, 7, 260), (1, 513, 4, 64)])   # narrow layer, large batch: slab dW
def test_linear_fp32_grouped(dev, n_models, B, N, K):
    from eeg_multimodal_b200 import _lib as L, ops

    g = torch.Generator().manual_seed(N + K)
    X = torch.randn(n_models, B, K, generator=g)
    W = torch.randn(n_models, N, K, generator=g) / K ** 0.5
    b = torch.randn(n_models, N, generator=g)
    dY = torch.randn(n_models, B, N, generator=g)
    Xd, Wd, bd, dYd = (t.to(dev) for t in (X, W, b, dY))
    Z = torch.einsum("mbk,mnk->mbn", X.double(), W.double()) + b.double()[:, None]
    for act, f in ((L.ACT_NONE, lambda z: z), (L.ACT_RELU, torch.relu), (L.ACT_TANH, torch.tanh)):
        assert rel_err(ops.linear_fwd(Xd, Wd, bd, act), f(Z)) < RTOL
    dX = torch.einsum("mbn,mnk->mbk", dY.double(), W.double())
    assert rel_err(ops.linear_bwd_dx(dYd, Wd), dX) < RTOL
    assert rel_err(ops.linear_bwd_dx(dYd, Wd, mask_src=Xd, mask_mode=L.ACT_RELU), dX * (X > 0)) < RTOL
    Xt = torch.tanh(Xd)
    assert rel_err(ops.linear_bwd_dx(dYd, Wd, mask_src=Xt, mask_mode=L.ACT_TANH), dX * (1 - Xt.cpu().double() ** 2)) < RTOL
    dW, db = ops.linear_bwd_dw(dYd, Xd)
    assert rel_err(dW, torch.einsum("mbn,mbk->mnk", dY.double(), X.double())) < RTOL
    assert rel_err(db, dY.double().sum(1)) < RTOL
    dW2, _ = ops.linear_bwd_dw(dYd, Xd, dW=dW.clone(), db=db.clone(), accumulate=True)
    assert rel_err(dW2, 2 * torch.einsum("mbn,mbk->mnk", dY.double(), X.double())) < RTOL
    if n_models == 1:  # ungrouped call shapes
        assert rel_err(ops.linear_fwd(Xd[0], Wd[0], bd[0], L.ACT_NONE), Z[0]) < RTOL


# ------------------------------------------------------------------------------------------------
# kernel (c) classifier + CE + accuracy
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_models,B,H", [(1, 8, 768), (1, 1, 768), (4, 601, 768), (2, 33, 256)])
def test_cls_ce_matches_cal_loss(dev, n_models, B, H):
    from eeg_multimodal_b200 import ops

    g = torch.Generator().manual_seed(B + H)
    h = torch.tanh(torch.randn(n_models, B, H, generator=g))
    Wc = torch.randn(n_models, 2, H, generator=g) / H ** 0.5
    bc = torch.randn(n_models, 2, generator=g) * 0.1
    labels = (torch.rand(B, generator=g) < 0.66).long()
    res = ops.cls_ce(h.to(dev), Wc.to(dev), bc.to(dev), labels.to(dev), loss_scale=1.0 / B, grad_scale=1.0 / B, backward=True)
    for m in range(n_models):
        z = torch.atanh(h[m].double().clamp(-0.999999, 0.999999)).float().requires_grad_(True)
        hh = torch.tanh(z)
        wc, b_ = Wc[m].clone().requires_grad_(True), bc[m].clone().requires_grad_(True)
        pred = torch.nn.functional.linear(hh, wc, b_)
        loss, acc, pid, _ = ho.cal_loss(pred, labels.view(B, 1))
        loss.backward()
        assert rel_err(res["logits"][m], pred.detach()) < RTOL
        assert torch.equal(res["pred"][m].cpu(), pid)                      # argmax: bit-exact
        st = res["stats"][m].cpu()
        assert abs(float(st[0]) - float(loss.detach())) < 1e-5 * max(1.0, float(loss.detach()))
        assert float(st[1]) == float((pid == labels).sum()) and abs(float(st[2]) - float(acc)) < 1e-6
        assert rel_err(res["dWc"][m], wc.grad) < 5e-5 and rel_err(res["dbc"][m], b_.grad) < 5e-5
        # dz = dL/d(pre-tanh); recomputed tanh from atanh loses a little, so 1e-4 here
        assert rel_err(res["dz"][m], z.grad) < 2e-4
    # pass-1 form (dz only: the weight-side gradients are discarded by the reference) and the fused
    # bias gradient of fc_layers.2 (column sums of dz), fp32 and bf16 dz
    p1 = ops.cls_ce(h.to(dev), Wc.to(dev), bc.to(dev), labels.to(dev), loss_scale=1.0 / B, grad_scale=1.0 / B, backward=True,
                    want_dw=False)
    assert p1["dWc"] is None and torch.equal(p1["dz"], res["dz"]) and torch.equal(p1["stats"], res["stats"])
    for dt in (torch.float32, torch.bfloat16):
        cs = torch.zeros(n_models, H, device=dev)
        p2 = ops.cls_ce(h.to(dev), Wc.to(dev), bc.to(dev), labels.to(dev), loss_scale=1.0 / B, grad_scale=1.0 / B,
                        backward=True, dz_dtype=dt, dz_colsum=cs)
        assert rel_err(cs, res["dz"].double().sum(1)) < 5e-5
        assert rel_err(p2["dWc"], res["dWc"]) < 1e-6 and rel_err(p2["dz"].float(), res["dz"]) < (1e-6 if dt == torch.float32 else 8e-3)
    # eval mode: no labels needed for logits/pred; bf16 activations
    ev = ops.cls_ce(h.to(dev).to(torch.bfloat16), Wc.to(dev), bc.to(dev), None, loss_scale=1.0, grad_scale=1.0, backward=False)
    ref16 = torch.einsum("mbh,mch->mbc", h.to(torch.bfloat16).double(), Wc.double()) + bc.double()[:, None]
    assert rel_err(ev["logits"], ref16) < 1e-5


def test_adam_matches_torch(dev):
    from eeg_multimodal_b200 import ops

    n = 10007
    g0 = torch.Generator().manual_seed(0)
    p = torch.randn(n, generator=g0)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3)
    pd, m, v = p.to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    shadow = torch.zeros(n, device=dev, dtype=torch.bfloat16)
    for step in range(1, 6):
        grad = torch.randn(n, generator=g0) * 10 ** float(torch.randint(-6, 1, (1,), generator=g0))
        ref.grad = grad.clone()
        opt.step()
        ops.adam_step(pd, grad.to(dev), m, v, step, lr=1e-3, bf16_shadow=shadow)
        assert float((pd.cpu() - ref.detach()).abs().max()) < 2e-7
    assert torch.equal(shadow, pd.to(torch.bfloat16))


@pytest.mark.parametrize("n_models,B,N,K", [(2, 8, 96, 256), (3, 5, 64, 192), (1, 8, 768, 2304), (24, 8, 96, 256), (30, 3, 64, 192)])
def test_linear_adam_streaming_kernel_is_bit_identical_to_dw_plus_adam(dev, n_models, B, N, K):
    """The TMA-fed gradient+Adam kernel (straight-line IEEE division / square-root sequences, library fallback per warp)
    against linear_bwd_dw + adam_step (library __fdiv_rn / __fsqrt_rn), bit for bit, with moments and gradients spread
    over the whole fp32 range: exact zeros (idle parameters), denormals, tiny and huge values."""
    from eeg_multimodal_b200 import ops

    g = torch.Generator(device=dev).manual_seed(N * K + B)

    def wide(shape, lo=-44.0, hi=12.0, p_zero=0.15):
        e = torch.rand(shape, device=dev, generator=g) * (hi - lo) + lo
        x = torch.pow(torch.tensor(10.0, device=dev), e) * torch.where(torch.rand(shape, device=dev, generator=g) < 0.5, -1.0, 1.0)
        return torch.where(torch.rand(shape, device=dev, generator=g) < p_zero, torch.zeros_like(x), x)

    P = N * K + N
    flat = torch.randn(n_models, P, device=dev, generator=g)
    m = wide((n_models, P), -42.0, 6.0)
    v = wide((n_models, P), -44.0, 10.0).abs()
    zero = torch.rand(n_models, P, device=dev, generator=g) < 0.2          # never-touched parameters: m = v = 0
    m[zero] = 0.0
    v[zero] = 0.0
    dY = wide((n_models, B, N), -30.0, 3.0, p_zero=0.3)
    X = torch.relu(torch.randn(n_models, B, K, device=dev, generator=g))   # layer inputs behind a ReLU: exact zeros
    X[:, :, ::7] = 0.0                                                     # units that are off for the whole batch

    def views(t):
        return t[:, :N * K].view(n_models, N, K), t[:, N * K:]

    for step in (1, 7):
        a = [t.clone() for t in (flat, m, v)]
        b = [t.clone() for t in (flat, m, v)]
        (Wa, ba), (mWa, mba), (vWa, vba) = (views(t) for t in a)
        ops.linear_adam_step(dY, X, Wa, mWa, vWa, ba, mba, vba, step=step, lr=1e-3)
        grad = torch.empty_like(flat)
        gW, gb = views(grad)
        ops.linear_bwd_dw(dY, X, dW=gW, db=gb)
        ops.adam_step(b[0], grad, b[1], b[2], step, lr=1e-3)
        torch.cuda.synchronize()
        for name, ta, tb in zip(("p", "m", "v"), a, b):
            same = (ta == tb) | (ta.isnan() & tb.isnan())
            assert bool(same.all()), f"step {step}: {name} differs in {int((~same).sum())} of {ta.numel()} elements"
        flat, m, v = a


def test_colsum_and_cast(dev):
    from eeg_multimodal_b200 import ops

    x = torch.randn(1000, 768, device=dev)
    assert rel_err(ops.colsum(x), x.double().sum(0)) < RTOL
    xb = ops.cast_bf16(x)
    assert torch.equal(xb, x.to(torch.bfloat16))
    assert rel_err(ops.colsum(xb), xb.double().sum(0)) < RTOL


# ------------------------------------------------------------------------------------------------
# kernel (b) tcgen05 tensor-core path
# ------------------------------------------------------------------------------------------------
def _gemm_ref(A, B, a_mn, b_mn):
    Am = A.double().cpu().t() if a_mn else A.double().cpu()
    Bm = B.double().cpu().t() if b_mn else B.double().cpu()
    return Am @ Bm.t()


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 256), (200, 264, 136), (1000, 768, 2304)])
def test_gemm_bf16_tcgen05_layouts(dev, a_mn, b_mn, M, N, K):
    from eeg_multimodal_b200 import _lib as L, ops

    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn((K, M) if a_mn else (M, K), generator=g)).to(torch.bfloat16).to(dev)
    Bm = (torch.randn((K, N) if b_mn else (N, K), generator=g) / K ** 0.5).to(torch.bfloat16).to(dev)
    ref = _gemm_ref(A, Bm, a_mn, b_mn)
    C = torch.full((M, N), float("nan"), device=dev)
    ops.gemm_bf16(A, Bm, C, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, epi=L.EPI_STORE_F32)
    torch.cuda.synchronize()
    assert rel_err(C, ref) < 1e-5, rel_err(C, ref)
    Cs = torch.zeros(M, N, device=dev)
    ops.gemm_bf16(A, Bm, Cs, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, epi=L.EPI_ATOMIC_F32, stream_k=True)
    assert rel_err(Cs, ref) < 1e-5, rel_err(Cs, ref)


def test_gemm_bf16_epilogues(dev):
    from eeg_multimodal_b200 import _lib as L, ops

    M, N, K = 520, 768, 512
    g = torch.Generator().manual_seed(1)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    aux = torch.randn(M, N, generator=g).to(torch.bfloat16).to(dev)
    z = _gemm_ref(A, W, False, False)
    zb = z + bias.double().cpu()
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    for epi, ref in ((L.EPI_STORE_BF16, z), (L.EPI_BIAS_RELU_BF16, torch.relu(zb)), (L.EPI_BIAS_TANH_BF16, torch.tanh(zb))):
        ops.gemm_bf16(A, W, out, M=M, N=N, K=K, epi=epi, bias=bias)
        assert rel_err(out, ref) < 6e-3, (epi, rel_err(out, ref))   # bf16 output rounding: 2^-8
    with pytest.raises(RuntimeError):                              # 3 (a bf16 mask-source tile) is no epilogue any more
        ops.gemm_bf16(A, W, out, M=M, N=N, K=K, epi=3)
    # ReLU sign bits: written by the forward epilogue (one uint32 per row x 32 columns), applied by the backward one
    bits = torch.full((M, N // 32), -1, dtype=torch.int32, device=dev)
    ops.gemm_bf16(A, W, out, M=M, N=N, K=K, epi=L.EPI_BIAS_RELU_BF16, bias=bias, aux=bits)
    assert rel_err(out, torch.relu(zb)) < 6e-3
    pos = (out.float() > 0).cpu()
    got = ((bits.cpu().view(M, N // 32, 1) >> torch.arange(32, dtype=torch.int32)) & 1).bool().view(M, N)
    assert torch.equal(got, pos)                                   # bit-exact against the stored activations
    back = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    cs = torch.empty(N, device=dev)
    ops.gemm_bf16(A, W, back, M=M, N=N, K=K, epi=L.EPI_BITMASK_BF16, aux=bits, colsum_out=cs)
    ref_back = z * pos.double()
    assert rel_err(back, ref_back) < 6e-3 and rel_err(cs, ref_back.sum(0)) < 2e-5
    o32 = torch.empty(M, N, device=dev)
    ops.gemm_bf16(A, W, o32, M=M, N=N, K=K, epi=L.EPI_BIAS_F32, bias=bias)
    assert rel_err(o32, zb) < 1e-5
    # fused tanh (MUFU.EX2 + MUFU.RCP form): absolute error at the fp32 rounding level, exact saturation
    big = bias.clone()
    big[:8] = 30.0
    big[8:16] = -30.0
    big[16:24] = 1e-4
    ops.gemm_bf16(A, W, o32, M=M, N=N, K=K, epi=L.EPI_BIAS_TANH_F32, bias=big)
    want = torch.tanh(z + big.double().cpu())
    assert float((o32.cpu().double() - want).abs().max()) < 5e-6   # incl. the fp32 accumulation error of z itself
    assert bool((o32[:, :8] == 1.0).all()) and bool((o32[:, 8:16] == -1.0).all())


def test_gemm_bf16_full_size_freivalds(dev):
    """BASELINE config-4 shapes (B=65536, D=2560): check C v == A (B^T v) instead of a full oracle."""
    from eeg_multimodal_b200 import _lib as L, ops

    Bsz, D, H = 65536, 2560, 768
    g = torch.Generator(device=dev).manual_seed(0)
    X = torch.rand(Bsz, D, device=dev, generator=g).to(torch.bfloat16)
    W1 = ((torch.rand(D, D, device=dev, generator=g) * 2 - 1) / D ** 0.5).to(torch.bfloat16)
    dZ = (torch.randn(Bsz, H, device=dev, generator=g) * 1e-3).to(torch.bfloat16)
    v = torch.randn(D, device=dev, generator=g, dtype=torch.float64)
    # forward GEMM, fp32 out
    C = torch.empty(Bsz, D, device=dev)
    ops.gemm_bf16(X, W1, C, M=Bsz, N=D, K=D, epi=L.EPI_STORE_F32)
    lhs = C.double() @ v
    rhs = X.double() @ (W1.double().t() @ v)
    assert float((lhs - rhs).abs().max() / rhs.abs().max()) < 1e-5
    # weight-gradient GEMM (K = batch, both operands MN-major, stream-K + fp32 reductions)
    dW = torch.zeros(H, D, device=dev)
    ops.gemm_bf16(dZ, X, dW, M=H, N=D, K=Bsz, a_mn=True, b_mn=True, epi=L.EPI_ATOMIC_F32, stream_k=True)
    lhs = dW.double() @ v
    rhs = dZ.double().t() @ (X.double() @ v)
    assert float((lhs - rhs).abs().max() / rhs.abs().max()) < 1e-4


@pytest.mark.parametrize("M,N,K", [(520, 768, 512), (128, 384, 64), (1000, 2304, 768), (4096, 2560, 768)])
def test_gemm_fused_bias_gradient(dev, M, N, K):
    """The sign-bit-mask epilogue also reduces its fp32 values over the rows (db1 = colsum(dZ1)): per-32-row
    slab partials inside the kernel, fixed-order sum outside -- no pass over the bf16 output."""
    from eeg_multimodal_b200 import _lib as L, ops

    g = torch.Generator().manual_seed(M + N)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
    W = (torch.randn(K, N, generator=g) / K ** 0.5).to(torch.bfloat16).to(dev)     # [K,N]: MN-major B, as dZ1 = dZ2 . W2
    keep = torch.rand(M, N, generator=g) < 0.5
    weights = (1 << torch.arange(32, dtype=torch.int64))
    words = (keep.view(M, N // 32, 32).long() * weights).sum(-1)
    bits = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32).to(dev)
    ref = _gemm_ref(A, W, False, True) * keep.double()
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    cs = torch.full((N,), float("nan"), device=dev)
    ops.gemm_bf16(A, W, out, M=M, N=N, K=K, b_mn=True, epi=L.EPI_BITMASK_BF16, aux=bits, colsum_out=cs)
    assert rel_err(out, ref) < 6e-3
    assert rel_err(cs, ref.sum(0)) < 2e-5, rel_err(cs, ref.sum(0))
    # deterministic: bit-identical on a second launch
    cs2 = torch.empty(N, device=dev)
    ops.gemm_bf16(A, W, out, M=M, N=N, K=K, b_mn=True, epi=L.EPI_BITMASK_BF16, aux=bits, colsum_out=cs2)
    assert torch.equal(cs, cs2)


@pytest.mark.parametrize("M,N,K,row0", [(520, 768, 512, 0), (100, 264, 64, 7), (1000, 2304, 2304, 123456), (8192, 2560, 2560, 65536 * 3)])
def test_gemm_fused_dDP_matches_unfused(dev, M, N, K, row0):
    """dDP from the fused input-gradient GEMM (dX stays in TMEM, noise regenerated in the epilogue) against
    the unfused pair: fp32 dX store + pgf_perturb_gate_bwd_dp, and against the numpy Philox oracle."""
    from eeg_multimodal_b200 import _lib as L, ops
    from oracle import philox_ref

    g = torch.Generator().manual_seed(M + K)
    dZ = (torch.randn(M, K, generator=g) * 1e-2).to(torch.bfloat16).to(dev)
    W = (torch.randn(K, N, generator=g) / K ** 0.5).to(torch.bfloat16).to(dev)
    deps = (torch.rand(N, generator=g) - 0.7).to(dev)
    seed, offset = 980616 + M, 5
    dX = torch.empty(M, N, device=dev)
    ops.gemm_bf16(dZ, W, dX, M=M, N=N, K=K, b_mn=True, epi=L.EPI_STORE_F32)
    unfused = ops.perturb_gate_bwd_dp(dX, deps, noise_mode=L.NOISE_PHILOX, seed=seed, offset=offset, row0=row0)
    fused = torch.full((N,), float("nan"), device=dev)
    ops.gemm_bf16_ddp(dZ, W, M=M, N=N, K=K, b_mn=True, seed=seed, offset=offset, row0=row0, deps_dDP=deps, out=fused)
    assert rel_err(fused, unfused) < 2e-5, rel_err(fused, unfused)
    if M <= 1000:
        lap = torch.from_numpy(philox_ref.laplace(seed, offset, row0, M, N)).double()
        ref = (dX.cpu().double() * lap).sum(0) * deps.cpu().double()
        assert rel_err(fused, ref) < 2e-5, rel_err(fused, ref)
    acc = fused.clone()
    ops.gemm_bf16_ddp(dZ, W, M=M, N=N, K=K, b_mn=True, seed=seed, offset=offset, row0=row0, deps_dDP=deps, out=acc, accumulate=True)
    assert rel_err(acc, 2 * fused) < 1e-6


@pytest.mark.parametrize("dims,out_dtype", [((2048, 512), torch.bfloat16), ((768, 768, 768), torch.float32), ((2048, 512), torch.float32)])
def test_perturb_ring_kernel_is_bit_identical_to_generic(dev, dims, out_dtype):
    """Large batches go through the TMA-ring variant (bulk async copies into a shared-memory row ring); halves of
    the same batch are small enough for the generic register kernel.  Same counters, same arithmetic: bit-equal."""
    from eeg_multimodal_b200 import _lib as L, ops

    D = sum(dims)
    B = 2 * ((1 << 22) // D // 2 + 37)                     # B*D >= 2^22 -> ring; B/2 * D < 2^22 -> generic
    g = torch.Generator(device=dev).manual_seed(3)
    blocks = [torch.randn(B, d, device=dev, generator=g) * 0.5 for d in dims]
    for b in blocks:
        b[5] = 0.25                                         # a constant row -> NaN row, like the reference
    blocks[-1][7, 3] = float("nan")
    DP = torch.randn(D, device=dev, generator=g) * 0.2
    w, eh, _ = ops.dp_coeffs(DP, ho.exp_eps_f32(1.0))
    full, _, mn, mx = ops.perturb_gate_fwd(blocks, w, eh, noise_mode=L.NOISE_PHILOX, seed=123, offset=9, row0=1 << 20,
                                           out_dtype=out_dtype, want_minmax=True)
    h = B // 2
    for lo in (0, h):
        part, _, pmn, pmx = ops.perturb_gate_fwd([b[lo:lo + h] for b in blocks], w, eh, noise_mode=L.NOISE_PHILOX, seed=123,
                                                 offset=9, row0=(1 << 20) + lo, out_dtype=out_dtype, want_minmax=True)
        assert torch.equal(part.view(torch.int16 if out_dtype == torch.bfloat16 else torch.int32),
                           full[lo:lo + h].view(torch.int16 if out_dtype == torch.bfloat16 else torch.int32))
        assert torch.equal(pmn.view(torch.int32), mn[lo:lo + h].view(torch.int32))
    assert bool(torch.isnan(full[5]).all()) and bool(torch.isnan(full[7]).all()) and bool(torch.isfinite(full[6]).all())
    # non-private path (model.py:53-64) through the ring as well
    nf, _, _, _ = ops.perturb_gate_fwd(blocks, None, None, noise_mode=L.NOISE_NONE, out_dtype=torch.float32)
    np_, _, _, _ = ops.perturb_gate_fwd([b[:h] for b in blocks], None, None, noise_mode=L.NOISE_NONE, out_dtype=torch.float32)
    assert torch.equal(nf[:h].view(torch.int32), np_.view(torch.int32))


def test_grouped_launch_with_per_model_seed_array(dev):
    """An eps x seed grid has arbitrary seeds: the grouped kernels read them from a device array, and every
    model's slice is bit-identical to a single-model launch with that seed (forward and dDP)."""
    from eeg_multimodal_b200 import _lib as L, ops

    M, B, dims = 5, 8, (768, 768, 768)
    D = sum(dims)
    g = torch.Generator(device=dev).manual_seed(4)
    blocks = [torch.rand(B, d, device=dev, generator=g) for d in dims]
    DP = torch.randn(M, D, device=dev, generator=g) * 0.2
    eeps = torch.tensor([ho.exp_eps_f32(e) for e in (0.1, 1.0, 3.0, 5.0, 8.0)], device=dev)
    w, eh, deps = ops.dp_coeffs(DP, eeps)
    seeds = [980616, 7, 2 ** 40 + 3, 980616, 2 ** 63 + 11]
    sdev = torch.tensor([s if s < 2 ** 63 else s - 2 ** 64 for s in seeds], dtype=torch.int64, device=dev)
    out, _, _, _ = ops.perturb_gate_fwd(blocks, w, eh, noise_mode=L.NOISE_PHILOX, model_seeds=sdev, offset=3, row0=16, n_models=M)
    dF = torch.randn(M, B, D, device=dev, generator=g)
    dDP = ops.perturb_gate_bwd_dp(dF, deps, noise_mode=L.NOISE_PHILOX, model_seeds=sdev, offset=3, row0=16)
    for m in range(M):
        one, _, _, _ = ops.perturb_gate_fwd(blocks, w[m], eh[m], noise_mode=L.NOISE_PHILOX, seed=seeds[m], offset=3, row0=16)
        assert torch.equal(one, out[m])
        d1 = ops.perturb_gate_bwd_dp(dF[m], deps[m], noise_mode=L.NOISE_PHILOX, seed=seeds[m], offset=3, row0=16)
        assert rel_err(d1, dDP[m]) < 1e-6
    # models 0 and 3 share a seed: identical noise, scaled by their own eps_hat
    n0 = (out[0] - out[3]).abs().max()
    assert float(n0) > 0 and not torch.equal(out[0], out[1])


@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_perturb_shared_batch_sweep_kernel_matches_single_model(dev, out_dtype):
    """A large batch shared by the sweep goes through the shared-batch ring kernel (row fetched and normalised once,
    perturbed once per model); every model's output is bit-identical to its own single-model launch."""
    from eeg_multimodal_b200 import _lib as L, ops

    dims, M = (2048, 512), 3
    D = sum(dims)
    B = (1 << 22) // D + 61
    g = torch.Generator(device=dev).manual_seed(8)
    blocks = [torch.rand(B, d, device=dev, generator=g) for d in dims]
    DP = torch.randn(M, D, device=dev, generator=g) * 0.2
    w, eh, _ = ops.dp_coeffs(DP, torch.tensor([ho.exp_eps_f32(e) for e in (0.1, 1.0, 8.0)], device=dev))
    seeds = [980616, 2 ** 63 + 5, 12]
    sdev = torch.tensor([s if s < 2 ** 63 else s - 2 ** 64 for s in seeds], dtype=torch.int64, device=dev)
    out, _, mn, _ = ops.perturb_gate_fwd(blocks, w, eh, noise_mode=L.NOISE_PHILOX, model_seeds=sdev, offset=2, row0=B, n_models=M,
                                         out_dtype=out_dtype, want_minmax=True)
    it = torch.int16 if out_dtype == torch.bfloat16 else torch.int32
    for m in range(M):
        one, _, mn1, _ = ops.perturb_gate_fwd(blocks, w[m], eh[m], noise_mode=L.NOISE_PHILOX, seed=seeds[m], offset=2, row0=B,
                                              out_dtype=out_dtype, want_minmax=True)
        assert torch.equal(one.view(it), out[m].view(it))
        assert torch.equal(mn1, mn[m])
    # an eps sweep at ONE seed (how the sweep grid lands on a GPU): the models share the noise computation inside the
    # kernel, and each still equals its own single-model launch bit for bit
    same = torch.tensor([55, 55, 9], dtype=torch.int64, device=dev)
    out3, _, _, _ = ops.perturb_gate_fwd(blocks, w, eh, noise_mode=L.NOISE_PHILOX, model_seeds=same, offset=2, row0=B, n_models=M,
                                         out_dtype=out_dtype)
    for m, sd in enumerate((55, 55, 9)):
        one, _, _, _ = ops.perturb_gate_fwd(blocks, w[m], eh[m], noise_mode=L.NOISE_PHILOX, seed=sd, offset=2, row0=B, out_dtype=out_dtype)
        assert torch.equal(one.view(it), out3[m].view(it))
    # arithmetic-progression seeds (no device array) take the same kernel
    out2, _, _, _ = ops.perturb_gate_fwd(blocks, w, eh, noise_mode=L.NOISE_PHILOX, seed=100, seed_step=7, offset=2, row0=B,
                                         n_models=M, out_dtype=out_dtype)
    one, _, _, _ = ops.perturb_gate_fwd(blocks, w[2], eh[2], noise_mode=L.NOISE_PHILOX, seed=114, offset=2, row0=B, out_dtype=out_dtype)
    assert torch.equal(one.view(it), out2[2].view(it))


def _guarded(shape, dtype, dev, fill):
    """A tensor of `shape` carved out of the middle of a larger sentinel-filled allocation; returns (view, check)."""
    n = int(np.prod(shape))
    pad = 4096
    big = torch.full((n + 2 * pad,), fill, dtype=dtype, device=dev)
    view = big[pad:pad + n].view(*shape)

    def check():
        lo, hi = big[:pad], big[pad + n:]
        ok = (lo == fill) & (hi == fill) if fill == fill else (torch.isnan(lo.float()) & torch.isnan(hi.float()))
        assert bool(ok.all()), "out-of-bounds write next to a kernel output"
    return view, check


def test_no_out_of_bounds_writes_at_ragged_sizes(dev):
    """compute-sanitizer is not available on the GPU pool: every kernel family writes into outputs carved out of
    sentinel-filled allocations at sizes that are not multiples of its tiles, and the guard zones must survive."""
    from eeg_multimodal_b200 import _lib as L, ops

    g = torch.Generator(device=dev).manual_seed(0)
    checks = []

    def G(shape, dtype=torch.float32, fill=float("nan")):
        v, c = _guarded(shape, dtype, dev, fill)
        checks.append(c)
        return v

    # tcgen05 GEMM: M, N not multiples of the 128/256 tiles; fp32 store, bf16 + sign bits, mask + column partials, dDP
    M, N, K = 200, 384, 136
    A = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
    W = torch.randn(N, K, device=dev, generator=g).to(torch.bfloat16)
    Wt = torch.randn(K, N, device=dev, generator=g).to(torch.bfloat16)
    ops.gemm_bf16(A, W, G((M, N)), M=M, N=N, K=K, epi=L.EPI_STORE_F32)
    bits = G((M, N // 32), torch.int32, fill=-7)
    ops.gemm_bf16(A, W, G((M, N), torch.bfloat16), M=M, N=N, K=K, epi=L.EPI_BIAS_RELU_BF16, bias=torch.zeros(N, device=dev), aux=bits)
    ops.gemm_bf16(A, Wt, G((M, N), torch.bfloat16), M=M, N=N, K=K, b_mn=True, epi=L.EPI_BITMASK_BF16, aux=bits, colsum_out=G((N,)))
    ops.gemm_bf16_ddp(A, Wt, M=M, N=N, K=K, b_mn=True, seed=3, offset=1, row0=5, deps_dDP=torch.ones(N, device=dev), out=G((N,)))
    Cs = G((N, 264), fill=0.0)
    ops.gemm_bf16(Wt, torch.randn(K, 264, device=dev, generator=g).to(torch.bfloat16), Cs, M=N, N=264, K=K, a_mn=True, b_mn=True,
                  epi=L.EPI_ATOMIC_F32, stream_k=True)
    # perturb: generic (small), TMA ring (large), shared-batch sweep ring
    dims = (772, 128, 64)
    D = sum(dims)
    for B, nm in ((3, 1), ((1 << 22) // D + 5, 1), ((1 << 22) // D + 5, 2)):
        blocks = [torch.rand(B, d, device=dev, generator=g) for d in dims]
        w, eh, _ = ops.dp_coeffs(torch.zeros(nm, D, device=dev) if nm > 1 else torch.zeros(D, device=dev), [2.7] * nm if nm > 1 else 2.7)
        for dt in (torch.float32, torch.bfloat16):
            ops.perturb_gate_fwd(blocks, w, eh, noise_mode=L.NOISE_PHILOX, seed=1, n_models=nm,
                                 out=G((nm, B, D) if nm > 1 else (B, D), dt))
    # classifier + CE: ragged batch, bf16 and fp32 activations, all modes
    B, H = 33, 264
    h = torch.tanh(torch.randn(2, B, H, device=dev, generator=g))
    Wc, bc = torch.randn(2, 2, H, device=dev, generator=g), torch.zeros(2, 2, device=dev)
    lab = (torch.rand(B, device=dev, generator=g) < 0.5).long()
    ops.cls_ce(h, Wc, bc, lab, loss_scale=1.0, grad_scale=1.0, backward=True, dz=G((2, B, H)), dWc=G((2, 2, H)), dbc=G((2, 2)),
               dz_colsum=G((2, H)), logits=G((2, B, 2)), pred=G((2, B), torch.int64, fill=-1), stats=G((2, 4)))
    ops.cls_ce(h.to(torch.bfloat16), Wc, bc, lab, loss_scale=1.0, grad_scale=1.0, backward=True, want_dw=False,
               dz=G((2, B, H), torch.bfloat16), logits=G((2, B, 2)), pred=G((2, B), torch.int64, fill=-1), stats=G((2, 4)))
    # grouped fp32 linear kernels and the fused gradient+Adam at sizes off every tile
    Mo, Bq, Nn, Kk = 3, 5, 100, 132
    X = torch.randn(Mo, Bq, Kk, device=dev, generator=g)
    Wl = G((Mo, Nn, Kk), fill=0.5)
    bl = torch.zeros(Mo, Nn, device=dev)
    ops.linear_fwd(X, Wl, bl, L.ACT_TANH, out=G((Mo, Bq, Nn)))
    dY = torch.randn(Mo, Bq, Nn, device=dev, generator=g)
    ops.linear_bwd_dx(dY, Wl, out=G((Mo, Bq, Kk)))
    ops.linear_bwd_dw(dY, X, dW=G((Mo, Nn, Kk)), db=G((Mo, Nn)))
    mW, vW = G((Mo, Nn, Kk), fill=0.0), G((Mo, Nn, Kk), fill=0.0)
    ops.linear_adam_step(dY, X, Wl, mW, vW, None, None, None, step=1, lr=1e-3)
    ops.adam_step(G((1001,), fill=1.0), torch.ones(1001, device=dev), G((1001,), fill=0.0), G((1001,), fill=0.0), 1)
    torch.cuda.synchronize()
    for c in checks:
        c()


def test_cls_ce_exact_tie_takes_the_first_index(dev):
    """torch.argmax returns the first maximal index (base_train.py:63): with both classifier rows and biases equal every
    logit pair ties EXACTLY, so every prediction must be class 0, accuracy the share of label 0, and the softmax gradient
    (0.5 - onehot) / B -- on the fp32 and on the bf16-activation path, at the reference batch and at a large one."""
    from eeg_multimodal_b200 import ops

    for B, dt in ((8, torch.float32), (8, torch.bfloat16), (601, torch.float32), (4096, torch.bfloat16)):
        g = torch.Generator().manual_seed(B)
        h = torch.tanh(torch.randn(B, 768, generator=g)).to(dev).to(dt)
        row = torch.randn(768, generator=g) * 0.05
        Wc = torch.stack([row, row]).to(dev)
        bc = torch.tensor([0.25, 0.25], device=dev)
        labels = (torch.rand(B, generator=g) < 0.66).long().to(dev)
        res = ops.cls_ce(h, Wc, bc, labels, loss_scale=1.0 / B, grad_scale=1.0 / B, backward=True, dz_dtype=dt)
        logits = res["logits"].cpu()
        assert torch.equal(logits[:, 0], logits[:, 1])                     # the tie is exact, not approximate
        assert int(res["pred"].abs().sum()) == 0                            # first index wins
        st = res["stats"].cpu()
        n0 = int((labels == 0).sum())
        assert float(st[1]) == n0 and abs(float(st[2]) - n0 / B) < 1e-7
        assert abs(float(st[0]) - 0.6931471805599453) < 1e-6                # mean CE of a 50/50 prediction = ln 2
        assert float(res["dbc"].cpu().sum().abs()) < 1e-6                   # softmax rows sum to one: dbc[0] + dbc[1] == 0
        exp_dbc0 = (0.5 * B - n0) / B
        assert abs(float(res["dbc"][0]) - exp_dbc0) < 1e-6
