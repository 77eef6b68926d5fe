"""GPU: the train.py-compatible driver end to end on a small sweep:
feature cache -> resident dataset -> fused graph-replayed sweep steps (batch <= 8) / HeadEngine steps -> n_eval repeated
stochastic evaluation as one batched pass -> reference-format results, records and checkpoints."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _separable_cache(path, n, dims, seed):
    """Two-class features whose class is linearly visible after row min-max normalisation."""
    from eeg_multimodal_b200 import feature_cache as fc

    g = torch.Generator().manual_seed(seed)
    labels = (torch.rand(n, generator=g) < 0.66).long()
    blocks = []
    for d in dims:
        x = torch.rand(n, d, generator=g)
        x[:, : d // 2] += labels[:, None].float() * 0.8
        blocks.append(x)
    fc.save_features(path, blocks, labels)


def test_train_driver_sweep_writes_reference_outputs(tmp_path, monkeypatch):
    from eeg_multimodal_b200 import train

    monkeypatch.chdir(tmp_path)
    tr, va = str(tmp_path / "train.npz"), str(tmp_path / "val.npz")
    _separable_cache(tr, 257, (64, 64, 64), 1)     # 257: ends in a 1-sample batch like the reference's 601 % 8
    _separable_cache(va, 65, (64, 64, 64), 2)
    out = train.run(train.build_parser().parse_args(
        ["--exp", "t", "--name", "sweep", "--batch_size", "8", "--n_epochs", "3", "--n_eval", "2", "--metrics", "Accuracy,F1Score",
         "--features", tr, "--val-features", va, "--eps-list", "1,8", "--n-seeds", "2", "--lr", "1e-3",
         "--records-root", str(tmp_path / "model_dict")]))
    base = tmp_path / "experiment" / "t" / "sweep"
    for f in ("debug.log", "info.log", "model.pth", "results.pth", "model_0.pth", "model_1.pth", "model_2.pth", "model_3.pth"):
        assert (base / f).exists(), f
    res = out["results"]
    assert len(out["grid"]) == 4 and sorted(res) == [0, 1, 2, 3] and out["best_acc_all"] == {i: res[i]["best_acc"] for i in res}
    ref0 = res[0]["reference"]
    # the reference's results.pth layout (train.py:131-144), everything a cat over the 3 epochs; 65 validation rows, n_eval 2
    assert ref0["Accuracy"].shape == (6,) and ref0["logits"].shape == (195, 2, 2) and ref0["pred"].shape == (195, 2)
    assert ref0["val_loss"].shape == (195, 2) and ref0["train_loss"].shape == (3 * 257,) and ref0["DP_params"].shape == (3, 192)
    assert torch.equal(ref0["logits"].argmax(-1), ref0["pred"])
    saved = torch.load(base / "results.pth")
    assert torch.equal(saved["logits"], ref0["logits"]) and sorted(saved["sweep_results"]) == [0, 1, 2, 3]
    assert max(r["best_acc"] for r in res.values()) > 0.9             # the separable toy problem is learnt
    for r in res.values():                                            # losses went down from the first to the last epoch
        tl = r["reference"]["train_loss"].view(3, 257).mean(1)
        assert tl[2] <= tl[0] + 1e-6
    out = {"DP_params": torch.cat([res[i]["reference"]["DP_params"][-1:] for i in sorted(res)])}
    sd = torch.load(base / "model.pth")
    assert set(sd) == {"fc_layers.0.weight", "fc_layers.0.bias", "fc_layers.2.weight", "fc_layers.2.bias",
                       "classifier.weight", "classifier.bias", "DP"}
    assert sd["DP"].shape == (1, 192) and sd["fc_layers.2.weight"].shape == (768, 192)
    assert float(out["DP_params"].abs().max()) > 0                    # pass 1 moved DP
    rec_dirs = sorted(os.listdir(tmp_path / "model_dict"))
    assert len(rec_dirs) == 4 and rec_dirs[0].startswith("newfrac_1.0eps")
    whole = open(tmp_path / "model_dict" / rec_dirs[0] / "whole_record.txt").read()
    assert whole.count("Epochs:") == 3 and "| f_1 Score:" in whole


def test_train_driver_no_dp_pass_keeps_DP(tmp_path, monkeypatch):
    """--n_dp 0 reproduces train.py:100-105 (DP pass commented out): DP never moves."""
    from eeg_multimodal_b200 import train

    monkeypatch.chdir(tmp_path)
    for bs in ("16", "8"):          # 16: engine steps; 8: the fused graph-replayed step with dp_pass = 0
        out = train.run(train.build_parser().parse_args(
            ["--synthetic", "64", "--feature-dims", "32,32", "--batch_size", bs, "--n_epochs", "2", "--n_eval", "1", "--n_dp", "0",
             "--eps", "1.0", "--lr", "1e-3"]))
        dp = out["results"][0]["reference"]["DP_params"]
        assert dp.shape == (2, 64) and float(dp.abs().max()) == 0.0


def test_sweep_accuracy_matches_reference_within_seed_noise():
    """North-star acceptance on a problem small enough for the CPU oracle: an eps sweep trained by the engine
    (Philox noise, one grouped launch per kernel for all models) reaches the accuracy the reference's own
    two-pass loop (torch RNG noise, oracle.TwoPassTrainer) reaches at each eps, within seed-to-seed spread, and
    shows the reference's trend: accuracy collapses towards the majority rate as eps -> 0."""
    from eeg_multimodal_b200 import HeadEngine
    from oracle import head_oracle as ho

    dev = torch.device("cuda:0")
    dims, D, Bsz, epochs, lr = (64, 64, 64), 192, 8, 4, 2e-3
    g = torch.Generator().manual_seed(11)

    def make(n):
        y = (torch.rand(n, generator=g) < 0.66).long()
        blocks = []
        for d in dims:
            x = torch.rand(n, d, generator=g)
            x[:, : d // 4] += y[:, None].float() * 0.35        # weakly separable: heavy noise destroys the signal
            blocks.append(x)
        return blocks, y

    tr_b, tr_y = make(384)
    va_b, va_y = make(256)
    eps_list, n_seeds = [0.02, 1.0, 8.0], 3
    p0 = ho.make_params(D, 768, seed=21)
    sd = {"fc_layers.0.weight": p0.W1, "fc_layers.0.bias": p0.b1, "fc_layers.2.weight": p0.W2, "fc_layers.2.bias": p0.b2,
          "classifier.weight": p0.Wc, "classifier.bias": p0.bc, "DP": p0.DP}

    # ---- engine: the whole grid at once
    grid = [(e, 100 + s) for e in eps_list for s in range(n_seeds)]
    eng = HeadEngine(n_models=len(grid), feature_dims=dims, eps=[e for e, _ in grid], seeds=[s for _, s in grid], lr=lr, precision="fp32")
    for i in range(len(grid)):
        eng.load_state_dict(i, sd)
    dtr, dty = [b.to(dev) for b in tr_b], tr_y.to(dev)
    dva, dvy = [b.to(dev) for b in va_b], va_y.to(dev)
    for ep in range(epochs):
        for i in range(0, 384, Bsz):
            eng.train_step([b[i:i + Bsz] for b in dtr], dty[i:i + Bsz])
    acc_eng = torch.stack([eng.eval_step(dva, dvy)["acc"].cpu() for _ in range(3)]).mean(0).view(len(eps_list), n_seeds)

    # ---- reference loop on the CPU (oracle), two seeds per eps
    acc_ref = torch.zeros(len(eps_list), 2)
    for ei, e in enumerate(eps_list):
        for s in range(2):
            t = ho.TwoPassTrainer(p0, e, lr=lr)
            draw = 1000 * (ei + 1) + 50 * s
            for ep in range(epochs):
                for i in range(0, 384, Bsz):
                    n1, g1 = ho.replay_reference_draws(draw, Bsz, D); draw += 1
                    n2, g2 = ho.replay_reference_draws(draw, Bsz, D); draw += 1
                    t.step([b[i:i + Bsz] for b in tr_b], tr_y[i:i + Bsz].view(-1, 1), n1, g1, n2, g2)
            accs = []
            for r in range(3):
                nv, gv = ho.replay_reference_draws(draw, 256, D); draw += 1
                with torch.no_grad():
                    pred = ho.head_forward(va_b, t.p, e, nv, gv, True)
                accs.append(float((pred.argmax(1) == va_y).float().mean()))
            acc_ref[ei, s] = sum(accs) / len(accs)
    me, mr = acc_eng.mean(1), acc_ref.mean(1)
    print("engine acc per eps", me.tolist(), "reference acc per eps", mr.tolist())
    assert float((me - mr).abs().max()) < 0.08, (me, mr)             # within seed noise at every eps
    assert me[2] > 0.9 and mr[2] > 0.9                               # large budget: the problem is learnt
    assert me[0] < me[2] - 0.1 and mr[0] < mr[2] - 0.1               # tiny budget: the noise drowns the signal


def test_train_driver_init_variants(tmp_path, monkeypatch):
    """--variants: the DP initialisations of model_dict/newfrac_1.0eps_{newinit,tt,newinit_k1}; the feature mean
    ('feawei') comes from the normalise kernel and matches the oracle's min-max normalisation."""
    from eeg_multimodal_b200 import feature_cache as fc, train, variants
    from oracle import head_oracle as ho

    monkeypatch.chdir(tmp_path)
    tr = str(tmp_path / "train.npz")
    _separable_cache(tr, 64, (64, 64, 64), 3)
    blocks, _ = fc.load_features(tr)
    fm = variants.feature_mean(blocks)
    want = ho.minmax_normalise(torch.cat(blocks, 1)).mean(0)
    assert float((fm - want).abs().max()) < 1e-6
    out = train.run(train.build_parser().parse_args(
        ["--batch_size", "8", "--n_epochs", "1", "--n_eval", "1", "--features", tr, "--eps", "1.0", "--variants", "newinit,tt,newinit_k1",
         "--lr", "1e-6", "--n_dp", "0"]))
    dp = torch.cat([out["results"][i]["reference"]["DP_params"][-1:] for i in range(3)])
    assert dp.shape == (3, 192)
    assert torch.allclose(dp[0], variants.dp_init("newinit", (64, 64, 64))) and torch.allclose(dp[1], variants.dp_init("tt", (64, 64, 64)))
    assert torch.allclose(dp[2], variants.dp_init("newinit_k1", (64, 64, 64), fm), atol=1e-6)


def test_batched_n_eval_equals_the_loop_of_single_evaluations():
    """train.py:126-131 evaluates every batch n_eval times with fresh noise.  eval_step(n_eval=k) does it as ONE batched
    pass (k Philox offsets inside one launch of every kernel): bit-identical to k consecutive eval_step calls."""
    from eeg_multimodal_b200 import HeadEngine

    dev = torch.device("cuda:0")
    dims, M, B, k = (128, 64, 64), 3, 7, 4
    kw = dict(n_models=M, feature_dims=dims, hidden=96, eps=[0.5, 1.0, 8.0], seeds=[5, 17, 18], precision="fp32", init_seed=3)
    a, b = HeadEngine(**kw), HeadEngine(**kw)
    g = torch.Generator(device=dev).manual_seed(0)
    blocks = [torch.rand(B, d, device=dev, generator=g) for d in dims]
    labels = (torch.rand(B, device=dev, generator=g) < 0.66).long()
    one = a.eval_step(blocks, labels, n_eval=k)
    loop = [b.eval_step(blocks, labels) for _ in range(k)]
    assert one["pred"].shape == (M, k, B) and one["logits"].shape == (M, k, B, 2)
    for e in range(k):
        assert torch.equal(one["logits"][:, e], loop[e]["logits"]) and torch.equal(one["pred"][:, e], loop[e]["pred"])
    assert a.noise_offset == b.noise_offset == k
    assert not torch.equal(one["logits"][:, 0], one["logits"][:, 1])            # fresh noise per repetition
    mean_acc = torch.stack([l["acc"] for l in loop]).mean(0)
    assert torch.allclose(one["acc"], mean_acc, atol=1e-6)


def test_out_of_range_label_poisons_the_loss():
    """F.cross_entropy raises on a label outside {0,1}; the fused kernel reports NaN instead of scoring it silently."""
    from eeg_multimodal_b200 import ops

    dev = torch.device("cuda:0")
    h = torch.tanh(torch.randn(8, 64, device=dev))
    Wc, bc = torch.randn(2, 64, device=dev), torch.zeros(2, device=dev)
    ok = ops.cls_ce(h, Wc, bc, torch.tensor([0, 1, 1, 0, 1, 1, 0, 1], device=dev), loss_scale=1 / 8, grad_scale=1 / 8, backward=False)
    bad = ops.cls_ce(h, Wc, bc, torch.tensor([0, 1, 2, 0, 1, 1, 0, 1], device=dev), loss_scale=1 / 8, grad_scale=1 / 8, backward=False)
    assert torch.isfinite(ok["stats"][0]) and torch.isnan(bad["stats"][0])


def test_engine_label_shapes():
    from eeg_multimodal_b200 import HeadEngine

    eng = HeadEngine(n_models=2, feature_dims=(32, 32), hidden=16, precision="fp32")
    dev = eng.device
    assert eng._labels(torch.zeros(5, 1, dtype=torch.int64, device=dev), 5).shape == (5,)
    assert eng._labels(torch.zeros(5, dtype=torch.int64, device=dev), 5).shape == (5,)
    assert eng._labels(torch.zeros(2, 5, dtype=torch.int64, device=dev), 5).shape == (2, 5)
    assert eng._labels(torch.zeros(2, 1, dtype=torch.int64, device=dev), 1).shape == (2, 1)     # per-model labels of a 1-row batch
    with pytest.raises(ValueError):
        eng._labels(torch.zeros(3, 5, dtype=torch.int64, device=dev), 5)
    with pytest.raises(ValueError):
        eng._labels(torch.zeros(4, dtype=torch.int64, device=dev), 5)
