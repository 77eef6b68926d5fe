"""The reference's OLDER PriGumbel head (SURVEY.md section 8 row a-alt; train_val.py:80-123,151-157,203-215).

CPU: the oracle restatement against tests/golden/prigumbel_golden.npz (generated from the UNMODIFIED reference functions,
tests/golden/make_golden_prigumbel.py), and against the live reference where it is present.
GPU: the CUDA path through the C ABI against the oracle on the same injected draws (fp32 bar 1e-5; predictions exact),
the Philox mode against the numpy Philox restatement, and the host class against the reference's golden logits, gradients
and 3-step Adam trajectory.
"""
import os

import numpy as np
import pytest
import torch

from oracle import philox_ref
from oracle import prigumbel_oracle as po
from oracle import ref_shim

HERE = os.path.dirname(os.path.abspath(__file__))
D, H = 2304, 768


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "prigumbel_golden.npz"))


@pytest.fixture(scope="module")
def split():
    g = np.load(os.path.join(HERE, "golden", "testsplit_golden.npz"))
    feat = torch.cat([torch.tensor(g[k].astype(np.float32)) for k in ("eeg", "act", "cm")], dim=1)
    return feat, torch.tensor(g["label"]).view(-1, 1)


def _case(gold, ci):
    eps, tau, train, B, r0 = gold["cases"][ci]
    return float(eps), float(tau), bool(train), int(B), int(r0)


def _proj(gold):
    rng = np.random.default_rng(int(gold["proj_seed"]))
    return {k: rng.standard_normal(n).astype(np.float32).astype(np.float64) for k, n in (("uD", D), ("vD", D), ("vH", H))}


# ------------------------------------------------------------------------------------------------------------------
# CPU: oracle vs golden
# ------------------------------------------------------------------------------------------------------------------
def test_golden_was_pinned_to_reference(gold):
    assert bool(gold["restatement_bitexact_at_generation"]) and len(gold["cases"]) == 7


@pytest.mark.parametrize("ci", range(7))
def test_oracle_matches_reference_golden(gold, split, ci):
    eps, tau, train, B, r0 = _case(gold, ci)
    feat, label = split[0][r0:r0 + B], split[1][r0:r0 + B]
    p = po.make_params(D, H, seed=int(gold["param_seed"])).clone(requires_grad=train)
    gum, lap = po.replay_reference_draws(int(gold["base_seed"]) + ci, B, H, eps)
    k = f"c{ci}_"
    with torch.set_grad_enabled(train):
        pred = po.head_forward(feat, p, tau, not train, gum, lap)
        loss, acc, pred_id, _ = po.loss_function(pred, label, p.w, float(gold["alpha"]), eps)
    # same torch build: bit-equal; other hosts may block the GEMMs differently
    np.testing.assert_allclose(pred.detach().numpy(), gold[k + "logits"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(loss.item(), gold[k + "loss"], rtol=2e-6)
    assert np.array_equal(pred_id.numpy(), gold[k + "pred"]) and abs(acc.item() - gold[k + "acc"]) < 1e-7
    if train:
        loss.backward()
        pr = _proj(gold)
        for n in ("b1", "b2", "Wc", "bc", "w"):
            np.testing.assert_allclose(getattr(p, n).grad.numpy(), gold[k + "d" + n], rtol=2e-4, atol=1e-9)
        np.testing.assert_allclose(p.W1.grad.double().numpy() @ pr["uD"], gold[k + "dW1_u"], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(pr["vH"] @ p.W2.grad.double().numpy(), gold[k + "v_dW2"], rtol=1e-4, atol=1e-9)


def test_oracle_trajectory_matches_reference_golden(gold, split):
    p = po.make_params(D, H, seed=int(gold["param_seed"]))
    tr = po.PriGumbelTrainer(p, 1.0, 0.01, float(gold["alpha"]), 1e-5)
    for s in range(3):
        gum, lap = po.replay_reference_draws(int(gold["base_seed"]) + 100 + s, 8, H, 1.0)
        loss, _, _ = tr.step(split[0][48 + 8 * s:56 + 8 * s], split[1][48 + 8 * s:56 + 8 * s], gum, lap)
        assert abs(loss - gold["traj_loss"][s]) < 2e-6 * abs(gold["traj_loss"][s])
    np.testing.assert_allclose(tr.p.w.detach().numpy(), gold["traj_w"], rtol=0, atol=2e-7)
    np.testing.assert_allclose(tr.p.Wc.detach().numpy(), gold["traj_Wc"], rtol=0, atol=2e-7)


def test_hard_gate_is_a_real_mask_and_soft_gate_saturates():
    """Unlike the main path's identity gate, this one drops columns: hard -> exactly 0 or
    x/(1-w) up to the straight-through rounding; at tau=0.01 the soft gate is within 1e-5 of it on >90 % of the entries."""
    g = torch.Generator().manual_seed(0)
    w = torch.rand(H, generator=g) * 0.9 + 0.05
    x = torch.randn(4, H, generator=g)
    gum = -torch.empty(H, 2).exponential_(generator=g).log()
    hard = po.gumbel_dropout(x, w, gum, 0.01, True)
    keep = (1 - w + gum[:, 1]) > (w + gum[:, 0])
    assert torch.equal(hard[:, ~keep], torch.zeros(4, int((~keep).sum())))
    assert torch.allclose(hard[:, keep], (x / (1 - w))[:, keep], rtol=1e-6)
    soft = po.gumbel_dropout(x, w, gum, 0.01, False)
    assert ((soft - hard).abs() < 1e-5).float().mean() > 0.9


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference checkout not present on this host")
def test_restatement_bitexact_vs_live_reference(split):
    shim = ref_shim.ShimmedPriGumbelHead(0.05, 2.0)
    p = po.make_params(D, H, seed=3)
    shim.load(p)
    blocks = [split[0][:8, i * 768:(i + 1) * 768] for i in range(3)]
    for train in (True, False):
        with torch.no_grad():
            ref = shim.forward(blocks, 1234, train)
        gum, lap = po.replay_reference_draws(1234, 8, H, 2.0)
        assert torch.equal(po.head_forward(split[0][:8], p, 0.05, not train, gum, lap), ref)


# ------------------------------------------------------------------------------------------------------------------
# GPU: CUDA path through the C ABI
# ------------------------------------------------------------------------------------------------------------------
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


@pytest.mark.gpu
@pytest.mark.parametrize("B,Hk,tau,hard", [(8, 768, 0.01, False), (8, 768, 0.01, True), (1, 768, 0.1, False), (37, 768, 1.0, True),
                                            (5, 64, 0.5, False), (300, 2048, 0.1, False), (3000, 768, 0.01, False)])
def test_kernels_match_oracle_with_injected_draws(dev, B, Hk, tau, hard):
    from eeg_multimodal_b200 import ops

    g = torch.Generator().manual_seed(B * 7 + Hk)
    z = torch.randn(B, Hk, generator=g)
    w = torch.rand(Hk, generator=g) * 0.9 + 0.05
    gum = -torch.empty(Hk, 2).exponential_(generator=g).log()
    eps = 0.7
    lap = torch.distributions.Laplace(0.0, 1 / eps).sample([B])
    dout = torch.randn(B, Hk, generator=g) / B
    exp_eps = float(np.float32(np.exp(eps)))
    # oracle: forward, the w-loss term and autograd
    zo, wo = z.clone().requires_grad_(True), w.clone().requires_grad_(True)
    out_o = po.lap_noise(po.gumbel_dropout(zo, wo, gum, tau, hard), lap)
    tmp = (1 - wo) * np.exp(eps) + wo
    loss_w = torch.max(tmp, dim=0)[0]
    ((out_o * dout).sum() + 0.3 * loss_w).backward()
    # CUDA
    zd, wd = z.to(dev), w.to(dev)
    coef, wloss = ops.prigumbel_coef(wd, exp_eps=exp_eps, tau=tau, hard=hard, gumbel=gum.to(dev))
    out, mn, mx = ops.prigumbel_fwd(zd, coef, eps=eps, lap=lap.to(dev), want_minmax=True)
    dz, dw = ops.prigumbel_bwd(zd, coef, dout.to(dev), wloss=wloss, exp_eps=exp_eps, wloss_scale=0.3)
    torch.cuda.synchronize()
    assert rel_err(out, out_o.detach()) < 1e-5
    r = po.gumbel_dropout(z, w, gum, tau, hard)
    assert rel_err(mn, r.min(dim=1)[0]) < 1e-6 and rel_err(mx, r.max(dim=1)[0]) < 1e-6
    assert abs(float(wloss[0]) - loss_w.item()) < 1e-6 * abs(loss_w.item()) and int(wloss[1]) == int(torch.argmax(tmp))
    assert rel_err(dz, zo.grad) < 1e-5
    assert rel_err(dw, wo.grad) < 2e-5


@pytest.mark.gpu
def test_philox_mode_follows_the_counter_convention(dev):
    """Production noise: Gumbel planes from counter (j/4, 0, plane, offset), row Laplace from (0, row0+b, Laplace, offset),
    key = seed -- the same Philox4x32-10 streams as the main path (oracle/philox_ref.py)."""
    from eeg_multimodal_b200 import ops

    B, seed, offset, row0, eps, tau = 19, 0x1234567890ABCDEF, 5, 1000, 2.0, 0.5
    g = torch.Generator().manual_seed(2)
    z = torch.randn(B, H, generator=g)
    w = torch.rand(H, generator=g) * 0.9 + 0.05
    gum = torch.tensor(philox_ref.gumbel(seed, offset, 0, 1, H))[:, 0, :].t().contiguous()           # [H,2]
    lap = torch.tensor(philox_ref.laplace(seed, offset, row0, B, 4)[:, 0]) / eps
    zd, wd = z.to(dev), w.to(dev)
    coef_i, _ = ops.prigumbel_coef(wd, exp_eps=2.0, tau=tau, hard=False, gumbel=gum.to(dev))
    coef_p, _ = ops.prigumbel_coef(wd, exp_eps=2.0, tau=tau, hard=False, seed=seed, offset=offset)
    assert rel_err(coef_p, coef_i) < 2e-5
    out_i = ops.prigumbel_fwd(zd, coef_i, eps=eps, lap=lap.to(dev))
    out_p = ops.prigumbel_fwd(zd, coef_i, eps=eps, seed=seed, offset=offset, row0=row0)
    assert rel_err(out_p, out_i) < 1e-5
    # splitting the batch over GPUs does not change the noise (row0 = global row)
    out_tail = ops.prigumbel_fwd(zd[7:], coef_i, eps=eps, seed=seed, offset=offset, row0=row0 + 7)
    assert torch.equal(out_tail, out_p[7:])


@pytest.mark.gpu
def test_argument_errors_are_reported(dev):
    from eeg_multimodal_b200 import ops

    w = torch.rand(66, device=dev)
    with pytest.raises(RuntimeError, match="multiple of 4"):
        ops.prigumbel_coef(w, exp_eps=2.0, tau=0.1, hard=False)
    with pytest.raises(RuntimeError, match="2048"):
        c, _ = ops.prigumbel_coef(torch.rand(4096, device=dev), exp_eps=2.0, tau=0.1, hard=False)
        ops.prigumbel_fwd(torch.randn(2, 4096, device=dev), c, eps=1.0)
    with pytest.raises(RuntimeError, match="tau"):
        ops.prigumbel_coef(torch.rand(64, device=dev), exp_eps=2.0, tau=0.0, hard=False)


def _head(dev, gold, eps, tau, lr=1e-5):
    from eeg_multimodal_b200.prigumbel import PriGumbelHead

    p = po.make_params(D, H, seed=int(gold["param_seed"]))
    head = PriGumbelHead(D, H, tau=tau, epsilon=eps, alpha=float(gold["alpha"]), lr=lr, device=dev)
    head.load_state_dict({"fc1.weight": p.W1, "fc1.bias": p.b1, "fc2.weight": p.W2, "fc2.bias": p.b2,
                          "classifier.weight": p.Wc, "classifier.bias": p.bc, "w": p.w})
    return head


@pytest.mark.gpu
@pytest.mark.parametrize("ci", range(7))
def test_head_matches_reference_golden(dev, gold, split, ci):
    """The host class on the reference's own draws: logits 1e-5, predictions exact, loss, and (train-mode cases) every
    gradient the reference's autograd produced."""
    eps, tau, train, B, r0 = _case(gold, ci)
    head = _head(dev, gold, eps, tau).train(train)
    feat, label = split[0][r0:r0 + B].to(dev), split[1][r0:r0 + B].to(dev)
    gum, lap = po.replay_reference_draws(int(gold["base_seed"]) + ci, B, H, eps)
    k = f"c{ci}_"
    head.inject_noise(gum, lap)
    if not train:
        res = head.eval_step(feat, label)
    else:
        head.lr = 0.0                                                   # keep the parameters: only the gradients are checked
        res = head.train_step(feat, label)
    torch.cuda.synchronize()
    # tau = 0.01 amplifies the fp32 rounding of exp() in the gate by 1/tau on the few unsaturated columns
    assert rel_err(res["logits"], torch.tensor(gold[k + "logits"])) < (1e-5 if tau >= 0.1 else 5e-5)
    assert np.array_equal(res["pred"].cpu().numpy(), gold[k + "pred"])
    assert abs(res["loss"] - float(gold[k + "loss"])) < 2e-5 * abs(float(gold[k + "loss"]))
    assert abs(res["acc"] - float(gold[k + "acc"])) < 1e-6
    if train:
        pr = _proj(gold)
        tol = 2e-5 if tau >= 0.1 else 2e-4
        for n in ("b1", "b2", "Wc", "bc", "w"):
            assert rel_err(head.view(n, head.grad), torch.tensor(gold[k + "d" + n])) < tol, n
        dW1, dW2 = head.view("W1", head.grad).double().cpu().numpy(), head.view("W2", head.grad).double().cpu().numpy()
        for got, want in ((dW1 @ pr["uD"], gold[k + "dW1_u"]), (pr["vD"] @ dW1, gold[k + "v_dW1"]),
                          (dW2 @ pr["uD"], gold[k + "dW2_u"]), (pr["vH"] @ dW2, gold[k + "v_dW2"])):
            assert np.abs(got - want).max() < tol * np.abs(want).max()


@pytest.mark.gpu
def test_head_trajectory_matches_reference_golden(dev, gold, split):
    """3 reference steps (Adam over every parameter, w included; train_val.py:178,203-215) at the reference's tau=0.01,
    eps=1, lr=1e-5: per-step loss and the final parameters."""
    head = _head(dev, gold, 1.0, 0.01, lr=1e-5).train()
    for s in range(3):
        gum, lap = po.replay_reference_draws(int(gold["base_seed"]) + 100 + s, 8, H, 1.0)
        head.inject_noise(gum, lap)
        res = head.train_step(split[0][48 + 8 * s:56 + 8 * s].to(dev), split[1][48 + 8 * s:56 + 8 * s].to(dev))
        assert abs(res["loss"] - float(gold["traj_loss"][s])) < 5e-5 * abs(float(gold["traj_loss"][s]))
    sd = head.state_dict()
    # Adam's first steps move every parameter by ~lr per step whatever the gradient's size: a wrong SIGN anywhere shows as 2e-5
    for key, name in (("w", "traj_w"), ("classifier.weight", "traj_Wc"), ("classifier.bias", "traj_bc"), ("fc2.bias", "traj_b2"),
                      ("fc1.bias", "traj_b1")):
        diff = (sd[key].cpu().double() - torch.tensor(gold[name]).double()).abs()
        assert float((diff > 5e-6).double().mean()) < 0.01, key         # elements whose gradient sits at the rounding level may flip
    pr = _proj(gold)
    np.testing.assert_allclose(sd["fc1.weight"].double().cpu().numpy() @ pr["uD"], gold["traj_W1_u"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(sd["fc2.weight"].double().cpu().numpy() @ pr["uD"], gold["traj_W2_u"], rtol=0, atol=2e-5)


@pytest.mark.gpu
def test_head_learns_with_philox_noise(dev, split):
    """End to end with the production noise: CE falls over 60 steps on the real test-split features at a workable lr,
    eval mode gives a hard mask, the privacy statistics of train_val.py:222-226 are reported."""
    from eeg_multimodal_b200.prigumbel import PriGumbelHead

    head = PriGumbelHead(D, H, tau=0.1, epsilon=4.0, alpha=5.0, lr=3e-4, device=dev, seed=11)
    feat, label = split[0][:256].to(dev), split[1][:256].to(dev)
    ces = []
    for s in range(60):
        lo = (s * 32) % 256
        ces.append(head.train_step(feat[lo:lo + 32], label[lo:lo + 32])["ce"])
    assert np.mean(ces[-8:]) < np.mean(ces[:8])
    res = head.eval().eval_step(feat[:64], label[:64])
    assert np.isfinite(res["loss"]) and 0.0 <= res["acc"] <= 1.0
    st = head.privacy_stats()
    assert 1.0 <= st["privacy_budget_avg"] <= st["privacy_budget_max"] <= np.exp(4.0) + 1e-3
    assert 0.0 <= st["drop_out_rate_avg"] <= st["drop_out_rate_max"]
