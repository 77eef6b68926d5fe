"""GPU parity tests of the fp32-parity tensor-core route (pgf_split3 / pgf_gemm_bf16x3, HeadEngine precision='fp32x3',
ConcatModel above tc_min_batch): the reference's fp32 nn.Linear arithmetic (models.py:46-51,80) at batches that are a
real dense contraction, held to north_star's fp32 bar of 1e-5 on logits and gradients."""
import numpy as np
import pytest
import torch

from oracle import head_oracle as ho

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def rel_err(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def planes_sum(p):
    """(hi + mid) + lo in fp32: exact, every partial sum is representable"""
    return (p[0].float() + p[1].float()) + p[2].float()


def test_split3_is_exact_and_applies_the_elementwise_stage(dev):
    from eeg_multimodal_b200 import _lib as L
    from eeg_multimodal_b200 import ops

    g = torch.Generator().manual_seed(0)
    R, C = 37, 264
    x = (torch.randn(R, C, generator=g) * torch.logspace(-6, 6, C)).to(dev)
    x[0, :8] = torch.tensor([0.0, -0.0, 1.0, -1.0, 3.0e-30, 1.0e30, 1.17549435e-38 * 8, 65504.0])
    p = ops.split3(x, planes=torch.empty(3, R, C, device=dev, dtype=torch.bfloat16))
    assert torch.equal(planes_sum(p), x)                       # hi + mid + lo == x bit for bit
    assert torch.equal(p[0], x.to(torch.bfloat16))             # hi is the round-to-nearest bf16 of x
    # bias + ReLU, written both ways; strided source
    big = torch.randn(R, 2 * C, generator=g).to(dev)
    src, bias = big[:, :C], torch.randn(C, generator=g).to(dev)
    out = torch.empty(R, C, device=dev)
    p = ops.split3(src, planes=torch.empty(3, R, C, device=dev, dtype=torch.bfloat16), out=out, bias=bias, act=L.ACT_RELU)
    ref = torch.relu(src + bias)
    assert torch.equal(out, ref) and torch.equal(planes_sum(p), ref)
    # tanhf in place
    t = src.contiguous().clone()
    ops.split3(t, out=t, act=L.ACT_TANH)
    assert float((t - torch.tanh(src)).abs().max()) <= 2.4e-7   # libm tanhf vs torch's: a few ulp
    # ReLU backward mask from the sign of the activation's hi plane
    act = torch.relu(torch.randn(R, C, generator=g)).to(dev)
    act[1, :3] = torch.tensor([0.0, 1e-30, -0.0])
    ap = ops.split3(act, planes=torch.empty(3, R, C, device=dev, dtype=torch.bfloat16))
    dz = ops.split3(src, out=torch.empty(R, C, device=dev), mask_plane=ap[0])
    assert torch.equal(dz, src * (act > 0))


def _operands(M, N, K, seed, dev):
    g = torch.Generator().manual_seed(seed)
    A = torch.rand(M, K, generator=g) + 0.3 * torch.randn(M, K, generator=g)      # perturbed features: [0,1] + noise
    B = (torch.rand(N, K, generator=g) * 2 - 1) / K ** 0.5                           # nn.Linear init
    return A.to(dev), B.to(dev)


@pytest.mark.parametrize("M,N,K", [(1000, 2560, 2560), (300, 768, 2304), (128, 2304, 520), (4096, 2560, 768)])
def test_gemm_bf16x3_matches_fp64_to_fp32_accuracy(dev, M, N, K):
    """C = A . B^T with both operands as plane triples: at least as close to the fp64 product as cuBLAS SGEMM is, and far
    inside the 1e-5 bar."""
    from eeg_multimodal_b200 import _lib as L
    from eeg_multimodal_b200 import ops

    A, B = _operands(M, N, K, M + N + K, dev)
    bias = torch.randn(N, device=dev) * 0.1
    ref = A.double() @ B.double().T + bias.double()
    bf = torch.bfloat16
    Ap, Bp = ops.split3(A, planes=torch.empty(3, M, K, device=dev, dtype=bf)), ops.split3(B, planes=torch.empty(3, N, K, device=dev, dtype=bf))
    C = ops.gemm_bf16x3(Ap, Bp, torch.empty(M, N, device=dev), M=M, N=N, K=K, epi=L.EPI_BIAS_F32, bias=bias)
    torch.backends.cuda.matmul.allow_tf32 = False
    sgemm = A @ B.T + bias
    e3, es = rel_err(C, ref), rel_err(sgemm, ref)
    # the contraction cut into slabs of ~10 k-blocks (what HeadEngine does): the tensor core truncates when it accumulates,
    # so the error grows with the length of one accumulation chain
    ns = max(2, round(((K + 63) // 64) / 10))
    Cs = torch.zeros(M, N, device=dev)
    ops.gemm_bf16x3(Ap, Bp, Cs, M=M, N=N, K=K, epi=L.EPI_ATOMIC_F32, k_slabs=ns)
    e3s = rel_err(Cs + bias, ref)
    print(f"\n[x3 gemm {M}x{N}x{K}] max-abs error / max-abs: x3 one chain {e3:.2e}, x3 {ns} slabs {e3s:.2e}, cuBLAS SGEMM {es:.2e}")
    assert e3 < 5e-6 and e3s < 1.5e-6 and e3s < 1.25 * es + 2e-7, (e3, e3s, es)
    # transposed second operand (dX = dZ . W with W stored [K, N]) through the same planes
    Bt = ops.split3(B.T.contiguous(), planes=torch.empty(3, K, N, device=dev, dtype=bf))
    C2 = ops.gemm_bf16x3(Ap, Bt, torch.empty(M, N, device=dev), M=M, N=N, K=K, b_mn=True, epi=L.EPI_STORE_F32)
    assert rel_err(C2, ref - bias.double()) < 5e-6        # one accumulation chain


@pytest.mark.parametrize("Bsz,N,K", [(1000, 768, 2560), (4096, 2560, 2560), (65, 2304, 2304)])
@pytest.mark.parametrize("slabs", [1, 7, 0])
def test_gemm_bf16x3_weight_gradient_layout(dev, Bsz, N, K, slabs):
    """dW[N,K] = dZ^T . X: both operands stored [batch, *] (MN-major), the contraction over a batch that is not a multiple
    of the 64-row k-block, cut into K slabs combined by fp32 reduce-adds."""
    from eeg_multimodal_b200 import _lib as L
    from eeg_multimodal_b200 import ops

    g = torch.Generator().manual_seed(Bsz)
    dZ = (torch.randn(Bsz, N, generator=g) * (torch.rand(Bsz, N, generator=g) < 0.5)).to(dev)   # half-zero like a ReLU-masked gradient
    X = torch.rand(Bsz, K, generator=g).to(dev)
    bf = torch.bfloat16
    dZp, Xp = ops.split3(dZ, planes=torch.empty(3, Bsz, N, device=dev, dtype=bf)), ops.split3(X, planes=torch.empty(3, Bsz, K, device=dev, dtype=bf))
    dW = torch.zeros(N, K, device=dev)
    if slabs == 0:   # HeadEngine's policy: ~10 k-blocks of 64 batch rows per accumulation chain
        slabs = max(1, round(((Bsz + 63) // 64) / 10))
    ops.gemm_bf16x3(dZp, Xp, dW, M=N, N=K, K=Bsz, a_mn=True, b_mn=True, epi=L.EPI_ATOMIC_F32, k_slabs=slabs)
    ref = dZ.double().T @ X.double()
    e = rel_err(dW, ref)
    print(f"\n[x3 dW batch {Bsz}, {slabs} slab(s)] max-abs error / max-abs {e:.2e}")
    assert e < (1.5e-6 if slabs >= Bsz / 64 / 16 else 1e-5), e


def _engine_and_oracle(dev, B, dims, H, precision, seed=5, lr=1e-3):
    from eeg_multimodal_b200 import HeadEngine

    Dd = sum(dims)
    g = torch.Generator().manual_seed(B)
    blocks = [torch.rand(B, d, generator=g) for d in dims]
    label = (torch.rand(B, 1, generator=g) < 0.66).long()
    eng = HeadEngine(n_models=1, feature_dims=dims, hidden=H, eps=1.0, lr=lr, precision=precision)
    p = ho.make_params(Dd, H, seed=seed, dp=(torch.randn(Dd, generator=g) * 0.1).numpy())
    eng.load_state_dict(0, {"fc_layers.0.weight": p.W1, "fc_layers.0.bias": p.b1, "fc_layers.2.weight": p.W2,
                            "fc_layers.2.bias": p.b2, "classifier.weight": p.Wc, "classifier.bias": p.bc, "DP": p.DP}, strict=True)
    return eng, p, blocks, label


@pytest.mark.parametrize("B,dims", [(1000, (2048, 512)), (1000, (768, 768, 768)), (257, (2048, 512))])
def test_engine_fp32x3_matches_reference_path_to_1e5(dev, B, dims):
    """north_star's fp32 bar on the tensor cores: logits, loss and every gradient of both passes within 1e-5 of the
    reference PyTorch fp32 path (same injected noise), predictions bit-exact."""
    H = 768
    eng, p, blocks, label = _engine_and_oracle(dev, B, dims, H, "fp32x3")
    Dd = sum(dims)
    lap, gum = ho.replay_reference_draws(4, B, Dd)
    db, lab = [b.to(dev) for b in blocks], eng._labels(label.to(dev))
    for mode, hard in (("dp", False), ("model", True)):
        po = p.clone(requires_grad=True)
        pred = ho.head_forward(blocks, po, 1.0, lap, gum, hard)
        loss, acc, pid, _ = ho.cal_loss(pred, label)
        loss.backward()
        eng.inject_noise(lap[None].to(dev), None)
        res = eng._pass(db, lab, hard=hard, mode=mode)
        torch.cuda.synchronize()
        e = rel_err(res["logits"][0], pred)
        print(f"\n[fp32x3 B={B} D={Dd} {mode}] logits rel err {e:.2e}")
        assert e < 1e-5
        assert abs(float(res["stats"][0, 0]) - float(loss.detach())) < 1e-5
        assert torch.equal(res["pred"][0].cpu(), pid)
        if mode == "dp":
            e = rel_err(eng.dDP[0], po.DP.grad.view(-1))
            print(f"  dDP rel err {e:.2e}")
            assert e < 1e-5
        else:
            for name, ref32 in (("W1", po.W1.grad), ("W2", po.W2.grad), ("b1", po.b1.grad), ("b2", po.b2.grad),
                                ("Wc", po.Wc.grad), ("bc", po.bc.grad)):
                e = rel_err(eng.view(name, eng.grad)[0], ref32)
                print(f"  d{name} rel err {e:.2e}")
                assert e < 1e-5, (name, e)


def test_engine_fp32_switches_to_the_tensor_route_and_agrees_with_the_cuda_core_path(dev):
    """precision='fp32' takes the fp32x3 route from tc_min_batch rows on; both routes are the same arithmetic to 1e-5
    (two full train steps with Philox noise: parameters after Adam agree)."""
    B, dims, H, lr = 1024, (768, 768, 768), 768, 1e-6
    eng_a, p, blocks, label = _engine_and_oracle(dev, B, dims, H, "fp32", lr=lr)
    eng_b, _, _, _ = _engine_and_oracle(dev, B, dims, H, "fp32", lr=lr)
    eng_b.tc_min_batch = 1 << 30            # stays on the CUDA-core kernels
    db, lab = [b.to(dev) for b in blocks], label.to(dev)
    from eeg_multimodal_b200 import _lib

    _lib.launch_by_name.clear()
    sa = [eng_a.train_step(db, lab) for _ in range(2)]
    assert _lib.launch_by_name.get("pgf_gemm_bf16x3", 0) > 0 and _lib.launch_by_name.get("pgf_linear_fwd", 0) == 0
    sb = [eng_b.train_step(db, lab) for _ in range(2)]
    for a, b in zip(sa, sb):
        assert abs(float(a["loss"][0]) - float(b["loss"][0])) < 1e-5
        assert float(a["n_correct"][0]) == float(b["n_correct"][0])
    assert rel_err(eng_a.grad, eng_b.grad) < 1e-5
    # dDP comes from pass 1, whose noise differs from pass 2's: of its 2.4 M pre-activations a handful lie within 1e-6 of
    # zero, where two fp32 implementations may pick different ReLU branches; one such flip moves a row of dX by ~3 %, i.e.
    # dDP by ~3 % / sqrt(B) (the bit-level agreement of the two routes is what the injected-noise tests above establish)
    e = rel_err(eng_a.dDP, eng_b.dDP)
    print(f"\n[fp32 vs fp32x3 route, B={B}] dDP rel err {e:.2e}")
    assert e < 2e-3
    # Adam moves every entry by ~lr per step whatever the size of its gradient, so entries whose gradient is ~0 may
    # go opposite ways: all entries within 2 steps x 2 lr, the bulk far tighter
    d = (eng_a.flat - eng_b.flat).abs()
    assert float(d.max()) <= 4.2 * lr and float((d > 0.02 * lr).float().mean()) < 2e-3


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_concat_model_takes_the_tensor_route_above_the_batch_threshold(dev, precision):
    """The module API (model.py:53-64 call shape) reaches the tensor cores: forward + autograd of ConcatModel at a batch
    above tc_min_batch against the reference PyTorch path with the same injected noise."""
    from eeg_multimodal_b200 import ConcatModel, _lib, cal_loss

    B, dims, Dd = 640, (2048, 512), 2560
    g = torch.Generator().manual_seed(11)
    blocks = [torch.rand(B, d, generator=g) for d in dims]
    label = (torch.rand(B, 1, generator=g) < 0.66).long()
    p = ho.make_params(Dd, 768, seed=9, dp=(torch.randn(Dd, generator=g) * 0.1).numpy())
    model = ConcatModel(feature_dims=dims, precision=precision, tc_min_batch=512).to(dev)
    model.load_state_dict({"fc_layers.0.weight": p.W1, "fc_layers.0.bias": p.b1, "fc_layers.2.weight": p.W2, "fc_layers.2.bias": p.b2,
                           "classifier.weight": p.Wc, "classifier.bias": p.bc, "DP": p.DP}, strict=False)
    model.eps = torch.tensor(1.0)
    lap, gum = ho.replay_reference_draws(3, B, Dd)
    po = p.clone(requires_grad=True)
    pred = ho.head_forward(blocks, po, 1.0, lap, gum, True)
    loss, _, pid, _ = ho.cal_loss(pred, label)
    loss.backward()
    _lib.launch_by_name.clear()
    model.inject_noise(lap.to(dev), gum.to(dev))
    out = model(tuple(b.to(dev) for b in blocks), hard=True)
    l2, _, pid2, _ = cal_loss(out, label.to(dev))
    l2.backward()
    torch.cuda.synchronize()
    used = "pgf_gemm_bf16x3" if precision == "fp32" else "pgf_gemm_bf16"
    assert _lib.launch_by_name.get(used, 0) >= 5 and _lib.launch_by_name.get("pgf_linear_fwd", 0) == 1   # only the 768 -> 2 classifier
    tol = 1e-5 if precision == "fp32" else 2e-2
    assert rel_err(out, pred) < tol
    fc0, fc2 = model.fc_layers[0], model.fc_layers[2]
    pairs = (("dW1", fc0.weight.grad, po.W1.grad), ("db1", fc0.bias.grad, po.b1.grad), ("dW2", fc2.weight.grad, po.W2.grad),
             ("db2", fc2.bias.grad, po.b2.grad), ("dWc", model.classifier.weight.grad, po.Wc.grad),
             ("dbc", model.classifier.bias.grad, po.bc.grad), ("dDP", model.DP.grad, po.DP.grad))
    for name, got, ref in pairs:
        e = rel_err(got, ref)
        if precision == "fp32":
            assert e < 1e-5, (name, e)
            assert torch.equal(pid2.cpu(), pid)
        else:   # bf16 operands flip ReLU signs: direction and Frobenius norm, not max-abs
            cos = float(torch.nn.functional.cosine_similarity(got.detach().cpu().double().flatten(), ref.double().flatten(), dim=0))
            assert cos > 0.995, (name, cos)


def test_fp32x3_full_size_sampled_rows_and_freivalds(dev):
    """BASELINE config 4's shape (65,536 x (2048+512)): 64 sampled rows of the logits against the fp32 oracle evaluated
    on the same perturbed rows, and the weight gradients by Freivalds' projection in fp64."""
    from eeg_multimodal_b200 import HeadEngine

    B, dims, Dd, H = 65536, (2048, 512), 2560, 768
    g = torch.Generator(device=dev).manual_seed(1)
    blocks = [torch.rand(B, d, generator=g, device=dev) for d in dims]
    label = (torch.rand(B, generator=g, device=dev) < 0.66).long()
    eng = HeadEngine(n_models=1, feature_dims=dims, hidden=H, eps=1.0, lr=1e-3, precision="fp32x3")
    res = eng._pass(blocks, label, hard=True, mode="model")
    torch.cuda.synchronize()
    rows = torch.randint(0, B, (64,), generator=torch.Generator().manual_seed(2))
    X = eng._bufs[("X", (1, B, Dd), torch.float32)][0]
    sd = {k: v.cpu().double() for k, v in eng.state_dict(0).items()}
    x = X[rows.to(dev)].cpu().double()
    h1 = torch.relu(x @ sd["fc_layers.0.weight"].T + sd["fc_layers.0.bias"])
    h2 = torch.tanh(h1 @ sd["fc_layers.2.weight"].T + sd["fc_layers.2.bias"])
    ref = h2 @ sd["classifier.weight"].T + sd["classifier.bias"]
    e = rel_err(res["logits"][0][rows.to(dev)], ref)
    print(f"\n[fp32x3 full size] 64 sampled logit rows vs fp64: rel err {e:.2e}")
    assert e < 1e-5
    # Freivalds: dW1 . v == dZ1^T . (X . v), dW2 . v == dZ2^T . (H1 . v) with the operands the kernels used (fp64 on the GPU)
    v = torch.randn(Dd, device=dev, dtype=torch.float64, generator=torch.Generator(device=dev).manual_seed(3))
    dZ1 = eng._bufs[("Z1f", (B, Dd), torch.float32)]                      # holds dZ1 after the pass
    lhs = eng.view("W1", eng.grad)[0].double() @ v
    rhs = dZ1.double().T @ (X.double() @ v)
    e1 = rel_err(lhs, rhs)
    H13 = eng._bufs[("H13", (3, B, Dd), torch.bfloat16)]
    dZ2 = eng._bufs[("dZ2f", (B, H), torch.float32)]
    h1v = (H13[0].double() + H13[1].double() + H13[2].double()) @ v
    e2 = rel_err(eng.view("W2", eng.grad)[0].double() @ v, dZ2.double().T @ h1v)
    e3 = rel_err(eng.view("b1", eng.grad)[0], dZ1.double().sum(0))
    print(f"[fp32x3 full size] Freivalds dW1 {e1:.2e}, dW2 {e2:.2e}, db1 {e3:.2e}")
    assert e1 < 1e-5 and e2 < 1e-5 and e3 < 1e-5
