"""GPU: the train.py-compatible driver end to end on a small sweep:
feature cache -> FeatureLoader -> HeadEngine two-pass steps -> n_eval repeated stochastic evaluation
-> reference-format records and checkpoints."""
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
    for f in ("debug.log", "info.log", "model.pth", "results.pth"):
        assert (base / f).exists(), f
    assert len(out["grid"]) == 4 and len(out["best_acc"]) == 4
    assert out["Accuracy"][0].shape == (4, 2)                         # [models, n_eval]
    assert max(out["best_acc"]) > 0.9                                 # the separable toy problem is learnt
    assert all(b >= a - 1e-6 for a, b in zip(out["train_loss"][-1], out["train_loss"][0]))  # losses went down
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
    out = train.run(train.build_parser().parse_args(
        ["--synthetic", "64", "--feature-dims", "32,32", "--batch_size", "16", "--n_epochs", "1", "--n_eval", "1", "--n_dp", "0",
         "--eps", "1.0", "--lr", "1e-3"]))
    assert float(out["DP_params"].abs().max()) == 0.0
