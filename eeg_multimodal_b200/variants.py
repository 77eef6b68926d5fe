"""Initialisation variants of the privacy-weight logits `DP` used by the reference's ablation runs
(`model_dict/newfrac_1.0eps_{newinit,tt,newinit_1,newinit_k1,newinit_k3,feawei}`, BASELINE config 5).

Reference (past_acc.py):
  :94      DP = zeros(1, 2304)                                              -> 'zeros' (w = 0.5 everywhere)
  :95      DP = cat(full(768, .4), full(768, .5), full(768, .3))            -> 'newinit'; the reversed order is 'tt'
  :98-103  weight = feawei.pkl (mean over samples of the normalised features, past_acc_feawei.py:131-148);
           z = (mean - mean.mean()) / mean.std();  w_init = 1 - sigmoid(k * z);
           DP = blocks(.4,.5,.3) + w_init - 0.5                             -> 'newinit_k<k>' ('newinit_1' == k 1), 'feawei'
`feawei.pkl` is not shipped; `feature_mean()` recomputes it from a feature cache with the normalise kernel.
"""
from __future__ import annotations

import re

import numpy as np
import torch

BLOCK_CONSTANTS = (0.4, 0.5, 0.3)   # EEG / OM / CM blocks, past_acc.py:95


def _blocks(dims, consts):
    if len(dims) != len(consts):
        raise ValueError(f"the block-constant initialisations are defined for {len(consts)} feature blocks, got {len(dims)}")
    return torch.cat([torch.full((d,), float(c)) for d, c in zip(dims, consts)])


def w_init_from_mean(mean_values, k: float = 1.0) -> torch.Tensor:
    """past_acc.py:99-102 (numpy mean/std, population std like np.std)."""
    m = np.asarray(mean_values, dtype=np.float64).reshape(-1)
    z = (m - m.mean()) / m.std()
    return 1.0 - torch.sigmoid(torch.tensor(k * z, dtype=torch.float32))


def dp_init(variant, dims=(768, 768, 768), feature_mean=None) -> torch.Tensor:
    """DP initial value [D] for a run-directory variant name (None / 'zeros' / 'newinit' / 'tt' / 'newinit_1' /
    'newinit_k1' / 'newinit_k3' / 'feawei')."""
    D = int(sum(dims))
    if variant in (None, "", "zeros", "newfrac"):
        return torch.zeros(D)
    if variant == "newinit":
        return _blocks(dims, BLOCK_CONSTANTS)
    if variant == "tt":
        return _blocks(dims, BLOCK_CONSTANTS[::-1])
    m = re.fullmatch(r"newinit_k?(\d+(?:\.\d+)?)|feawei", variant)
    if m:
        if feature_mean is None:
            raise ValueError(f"variant '{variant}' needs the mean normalised feature vector (variants.feature_mean)")
        k = float(m.group(1)) if m.group(1) else 1.0
        w = w_init_from_mean(feature_mean, k)
        if w.numel() != D:
            raise ValueError(f"feature_mean has {w.numel()} entries, expected {D}")
        return _blocks(dims, BLOCK_CONSTANTS) + w - 0.5
    raise ValueError(f"unknown DP initialisation variant '{variant}'")


def feature_mean(blocks, chunk: int = 65536) -> torch.Tensor:
    """Mean over samples of the row-min-max-normalised concatenated features (the content of the reference's
    feawei.pkl, past_acc_feawei.py:131-148), computed with the normalise kernel (PGF_NOISE_NONE) on the GPU."""
    from . import _lib as L, ops

    n = blocks[0].shape[0]
    total = None
    for lo in range(0, n, chunk):
        part = [b[lo:lo + chunk].cuda().contiguous() for b in blocks]
        out, _, _, _ = ops.perturb_gate_fwd(part, None, None, noise_mode=L.NOISE_NONE, out_dtype=torch.float32)
        s = out.double().sum(0)
        total = s if total is None else total + s
    return (total / n).float().cpu()
