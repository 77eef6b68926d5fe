"""GPU: the fused, graph-replayed sweep step (pgf_sweep_plan_*, csrc/sweep_step.cu) against the one-call-per-kernel
engine path it replaces.  The plan runs the same kernels with the same arithmetic -- it only removes launches and
host work -- so every parameter, optimiser moment and statistic must be BIT-identical, for direct launches, for
CUDA-graph replay (1 and several steps per graph), with and without programmatic dependent launch, with the batches
gathered from a resident dataset through a device permutation."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _engines(dev, M, dims, H, lr=1e-3, eps=None, seeds=None, dp_init=None):
    from eeg_multimodal_b200 import HeadEngine

    eps = eps or [[0.1, 1.0, 3.0, 8.0][i % 4] for i in range(M)]
    seeds = seeds or [980616 + 7 * i for i in range(M)]        # not an arithmetic progression of step 1
    kw = dict(n_models=M, feature_dims=dims, hidden=H, eps=eps, seeds=seeds, lr=lr, precision="fp32", init_seed=11, dp_init=dp_init)
    return HeadEngine(**kw), HeadEngine(**kw)


def _state(e):
    return {"flat": e.flat, "m": e.m, "v": e.v, "DP": e.DP, "DP_m": e.DP_m, "DP_v": e.DP_v}


def _assert_same(a, b, what):
    for k, ta in _state(a).items():
        tb = _state(b)[k]
        assert torch.equal(ta, tb), f"{what}: {k} differs (max abs {float((ta - tb).abs().max()):.3e})"


def _dataset(dev, n, dims, seed=3):
    g = torch.Generator(device=dev).manual_seed(seed)
    blocks = [torch.rand(n, d, device=dev, generator=g) for d in dims]
    labels = (torch.rand(n, device=dev, generator=g) < 0.66).long()
    return blocks, labels


@pytest.mark.parametrize("use_pdl", [False, True])
@pytest.mark.parametrize("dims,H,M,B", [((768, 768, 768), 768, 3, 8), ((128, 64), 32, 4, 8), ((256, 128, 128), 64, 2, 3),
                                        ((128, 64), 32, 24, 8), ((256, 128, 128), 64, 26, 5)])   # >= 24 models: the slab kernels of linear_wide.cu
def test_plan_is_bit_identical_to_the_engine_path(dev, dims, H, M, B, use_pdl):
    from eeg_multimodal_b200.sweep_plan import SweepStepPlan

    ref, fused = _engines(dev, M, dims, H)
    n = 5 * B
    blocks, labels = _dataset(dev, n, dims)
    perm = torch.randperm(n, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
    plan = SweepStepPlan(fused, blocks, labels, B, use_pdl=use_pdl)
    plan.set_rows(perm)

    def ref_step(i):
        idx = perm[(i % 5) * B:(i % 5) * B + B]
        return ref.train_step([b[idx] for b in blocks], labels[idx])

    # direct launches
    for i in range(3):
        st = ref_step(i)
        plan.run(1)
        torch.cuda.synchronize()
        _assert_same(ref, fused, f"direct step {i}")
        assert torch.equal(st["stats"], plan.stats_model), f"direct step {i}: statistics differ"
    # graph replay, one step per graph; the cursor wraps after the fifth batch
    plan.capture(1)
    for i in range(3, 7):
        st = ref_step(i)
        plan.run(1)
        torch.cuda.synchronize()
        _assert_same(ref, fused, f"graph step {i}")
        assert torch.equal(st["stats"], plan.stats_model)
        assert plan.cursor == ((i + 1) % 5) * B
    # several steps per graph
    plan.capture(3)
    for i in range(7, 10):
        st = ref_step(i)
    plan.run(3)
    torch.cuda.synchronize()
    _assert_same(ref, fused, "3-step graph")
    assert torch.equal(st["stats"], plan.stats_model)
    # the engine's own counters followed the plan, so the two paths can be mixed
    assert (fused.noise_offset, fused.t_dp, fused.t_model) == (ref.noise_offset, ref.t_dp, ref.t_model)
    idx = perm[:B]
    ref.train_step([b[idx] for b in blocks], labels[idx])
    fused.train_step([b[idx] for b in blocks], labels[idx])
    plan.set_rows(perm, cursor=B)
    ref_step(1)
    plan.run(1)
    torch.cuda.synchronize()
    _assert_same(ref, fused, "mixed engine / plan steps")
    # predictions and logits of the last pass 2
    res = plan.result()
    assert res["pred"].shape == (M, B) and res["logits"].shape == (M, B, 2)
    assert torch.equal(res["logits"].argmax(-1), res["pred"])


def test_plan_without_the_dp_pass(dev):
    """train.py has the DP pass commented out (train.py:100-105): dp_pass=False."""
    from eeg_multimodal_b200.sweep_plan import SweepStepPlan

    dims, H, M, B = (128, 64), 32, 2, 8
    ref, fused = _engines(dev, M, dims, H)
    blocks, labels = _dataset(dev, 4 * B, dims)
    plan = SweepStepPlan(fused, blocks, labels, B, dp_pass=False)
    plan.set_rows(None)
    plan.run(1)
    plan.capture(1)
    plan.run(3)
    for i in range(4):
        ref.train_step([b[i * B:(i + 1) * B] for b in blocks], labels[i * B:(i + 1) * B], dp_pass=False)
    torch.cuda.synchronize()
    _assert_same(ref, fused, "dp_pass=False")
    assert fused.t_dp == 0 and fused.t_model == 4 and fused.noise_offset == 4


def test_plan_with_trained_dp_and_many_steps(dev):
    """A non-trivial DP (w != 0.5) and enough steps for the Adam bias corrections to matter: the device-side
    coefficients (square-and-multiply in double) must equal the host's."""
    from eeg_multimodal_b200.sweep_plan import SweepStepPlan

    dims, H, M, B = (128, 64), 32, 2, 8
    dp0 = torch.linspace(-2.0, 2.0, sum(dims))
    ref, fused = _engines(dev, M, dims, H, lr=3e-3, dp_init=dp0)
    blocks, labels = _dataset(dev, 8 * B, dims)
    plan = SweepStepPlan(fused, blocks, labels, B)
    plan.set_rows(None)
    plan.run(1)
    plan.capture(4)
    plan.run(36)
    for i in range(37):
        j = i % 8
        ref.train_step([b[j * B:(j + 1) * B] for b in blocks], labels[j * B:(j + 1) * B])
    torch.cuda.synchronize()
    _assert_same(ref, fused, "37 steps")
    assert not torch.equal(fused.DP, dp0.to(dev).expand_as(fused.DP))


def test_single_launch_dx_matches_the_two_launch_form(dev):
    """linear_bwd_dx with the last-CTA reduction (plan) vs partial + finalize launches (ordinary entry point)."""
    from eeg_multimodal_b200 import ops

    g = torch.Generator(device=dev).manual_seed(1)
    M, B, N, K = 3, 8, 768, 2304
    dY = torch.randn(M, B, N, device=dev, generator=g)
    W = torch.randn(M, N, K, device=dev, generator=g)
    H1 = torch.randn(M, B, K, device=dev, generator=g)
    a = ops.linear_bwd_dx(dY, W, mask_src=H1)
    b = ops.linear_bwd_dx(dY, W, mask_src=H1)
    assert torch.equal(a, b)
    ref = torch.einsum("mbn,mnk->mbk", dY.double(), W.double()) * (H1 > 0)
    assert float((a.double() - ref).abs().max() / ref.abs().max()) < 1e-5


def test_plan_rejects_what_it_cannot_run(dev):
    from eeg_multimodal_b200 import HeadEngine
    from eeg_multimodal_b200.sweep_plan import SweepStepPlan

    eng = HeadEngine(n_models=1, feature_dims=(128, 64), hidden=32, precision="fp32")
    blocks, labels = _dataset(dev, 64, (128, 64))
    with pytest.raises(ValueError, match="1..8"):
        SweepStepPlan(eng, blocks, labels, 16)
    with pytest.raises(ValueError, match="widths"):
        SweepStepPlan(eng, [blocks[0]], labels, 8)
    with pytest.raises(ValueError, match="fewer rows"):
        SweepStepPlan(eng, [b[:4] for b in blocks], labels[:4], 8)
