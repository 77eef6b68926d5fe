"""CPU: the oracle restatement against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py), and against the reference itself when it is present."""
import numpy as np
import pytest
import torch

from conftest import case_fields, golden_dp
from oracle import head_oracle as ho
from oracle import philox_ref, ref_shim

D = 2304


def _inputs(golden):
    blocks = [torch.from_numpy(golden[k]) for k in ("eeg", "act", "cm")]
    return blocks, torch.from_numpy(golden["lap"]), torch.from_numpy(golden["gum"]), torch.from_numpy(golden["label"])


def test_golden_was_pinned_to_reference(golden):
    assert bool(golden["restatement_bitexact_at_generation"])
    assert len(golden["cases"]) == 16


def test_replayed_draws_match_golden(golden):
    lap, gum = ho.replay_reference_draws(int(golden["noise_seed"]), 8, D)
    np.testing.assert_allclose(lap.numpy(), golden["lap"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(gum.numpy(), golden["gum"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("idx", range(16))
def test_oracle_matches_reference_golden(golden, idx):
    key = str(golden["cases"][idx])
    kind, eps, hard = case_fields(key)
    blocks, lap, gum, label = _inputs(golden)
    p = ho.make_params(D, seed=int(golden["param_seed"]), dp=golden_dp(golden, kind)).clone(requires_grad=True)
    pred, aux = ho.head_forward(blocks, p, eps, lap, gum, hard, return_aux=True)
    loss, acc, pred_id, _ = ho.cal_loss(pred, label)
    loss.backward()
    # float results: same torch build gives bit-equality here; other hosts may differ in GEMM blocking
    np.testing.assert_allclose(pred.detach().numpy(), golden[key + "_logits"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(float(loss.detach()), float(golden[key + "_loss"]), rtol=1e-5)
    assert float(acc) == float(golden[key + "_acc"])
    # integer results must be bit-exact
    np.testing.assert_array_equal(pred_id.numpy(), golden[key + "_pred"])
    np.testing.assert_array_equal(aux["gate_index"].numpy().astype(np.uint8), golden[f"{kind}_gate_index"])
    scale = np.abs(golden[key + "_dDP"]).max()
    np.testing.assert_allclose(p.DP.grad.numpy(), golden[key + "_dDP"], rtol=1e-4, atol=1e-5 * scale)
    np.testing.assert_allclose(p.b2.grad.numpy(), golden[key + "_db2"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(p.W1.grad[:4].numpy(), golden[key + "_dW1_rows"], rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(aux["perturbed"][:2].detach().numpy(), golden[key + "_perturbed_rows"], rtol=1e-6, atol=1e-6)


def test_hard_gate_is_exact_identity_soft_within_two_ulp(golden):
    """Known-answer property of models.py:77-79."""
    blocks, lap, gum, _ = _inputs(golden)
    p = ho.make_params(D, seed=7, dp=golden_dp(golden, "wvalues"))
    for hard in (True, False):
        _, aux = ho.head_forward(blocks, p, 1.0, lap, gum, hard, return_aux=True)
        if hard:
            assert torch.equal(aux["gated"], aux["perturbed"])
        else:
            ulp = torch.abs(aux["perturbed"]) * 2.0 ** -23 + 1e-45
            assert bool(((aux["gated"] - aux["perturbed"]).abs() <= 2.01 * ulp).all())  # 3 roundings


def test_constant_row_gives_nan_like_reference():
    x = torch.ones(2, 8)
    x[1] = torch.arange(8.0)
    out = ho.minmax_normalise(x)
    assert torch.isnan(out[0]).all() and not torch.isnan(out[1]).any()


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32 10 rounds."""
    h = lambda a: [int(x) for x in a]
    assert h(philox_ref.philox4x32_10(0, 0, 0, 0, 0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert h(philox_ref.philox4x32_10(*[0xFFFFFFFF] * 4, 0xFFFFFFFF, 0xFFFFFFFF)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert h(philox_ref.philox4x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_philox_noise_distributions():
    lap = philox_ref.laplace(980616, 3, 0, 64, 2304).ravel().astype(np.float64)
    gum = philox_ref.gumbel(980616, 3, 0, 64, 2304).ravel().astype(np.float64)
    assert abs(lap.mean()) < 0.02 and abs(lap.var() - 2.0) < 0.05          # Laplace(0,1): var 2
    assert abs(np.abs(lap).mean() - 1.0) < 0.02
    assert abs(gum.mean() - 0.5772) < 0.02 and abs(gum.var() - np.pi ** 2 / 6) < 0.05
    # partition invariance: rows 32..63 generated alone equal the tail of the full draw
    a = philox_ref.laplace(1, 0, 32, 32, 64)
    b = philox_ref.laplace(1, 0, 0, 64, 64)[32:]
    np.testing.assert_array_equal(a, b)


def test_exp_eps_follows_reference_promotion():
    assert ho.exp_eps_f32(1.0) == float(torch.tensor(1.0).exp())
    e = np.around(np.float64(0.616), 3)
    assert ho.exp_eps_f32(e) == float(torch.tensor(e).exp().to(torch.float32))


def test_two_pass_trainer_runs(golden):
    blocks, lap, gum, label = _inputs(golden)
    p = ho.make_params(D, seed=7)
    tr = ho.TwoPassTrainer(p, 1.0, lr=1e-3)
    loss, acc, _ = tr.step(blocks, label, lap, gum, lap.flip(0), gum.flip(1))
    assert np.isfinite(loss) and 0.0 <= acc <= 1.0
    assert not torch.equal(tr.p.DP.detach(), p.DP) and not torch.equal(tr.p.W1.detach(), p.W1)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference checkout not present on this host")
def test_restatement_bitexact_vs_live_reference(golden):
    blocks, lap, gum, label = _inputs(golden)
    shim = ref_shim.ShimmedReferenceHead()
    p = ho.make_params(D, seed=11, dp=golden_dp(golden, "wvalues"))
    shim.load(p)
    for eps, hard in ((1.0, True), (0.1, False)):
        ref = shim.forward(blocks, eps, hard, int(golden["noise_seed"]))
        mine = ho.head_forward(blocks, p, eps, lap, gum, hard)
        assert torch.equal(ref, mine)
