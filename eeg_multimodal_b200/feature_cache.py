"""Feature cache: the on-disk input of the hot path.

The reference recomputes BERT / CLIP-projection / cross-attention features every step
(99.8 % of its wall time); with frozen encoders the [N, 768] x 3 blocks can be extracted once --
the authors' own `past_acc_feawei.py:138-148` dumps exactly such a (2402, 2304) matrix.
Format: one `.npz` with float32 arrays `block0..blockK` ([N, d_i]) and int64 `label` ([N]); the
loader mirrors `data.py:37-45` (`DataLoader(batch_size, shuffle=True)` for train AND val, last
partial batch kept) and stages batches through pinned host memory.
"""
from __future__ import annotations

import numpy as np
import torch


def save_features(path: str, blocks, labels):
    arrs = {f"block{i}": np.ascontiguousarray(np.asarray(b, dtype=np.float32)) for i, b in enumerate(blocks)}
    n = arrs["block0"].shape[0]
    lab = np.asarray(labels).reshape(-1).astype(np.int64)
    if any(a.shape[0] != n for a in arrs.values()) or lab.shape[0] != n:
        raise ValueError("all blocks and the labels must have the same number of rows")
    np.savez(path, label=lab, **arrs)


def load_features(path: str):
    z = np.load(path)
    keys = sorted((k for k in z.files if k.startswith("block")), key=lambda k: int(k[5:]))
    return [torch.from_numpy(z[k]) for k in keys], torch.from_numpy(z["label"])


def synthetic_features(n, dims=(2048, 512), seed=980616, p_label1=0.66):
    """SURVEY.md section 8d synthetic inputs: U(0,1) features, Bernoulli(0.66) labels."""
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(n, d, generator=g) for d in dims], (torch.rand(n, generator=g) < p_label1).long()


class FeatureLoader:
    """Batches of (blocks, labels) on `device`; shuffle=True reshuffles every epoch from `seed`."""

    def __init__(self, blocks, labels, batch_size, shuffle=True, seed=980616, device="cuda", drop_last=False):
        self.blocks = [b.contiguous() for b in blocks]
        self.labels = labels.reshape(-1).contiguous()
        self.n, self.bs, self.shuffle, self.drop_last = self.labels.shape[0], int(batch_size), shuffle, drop_last
        self.gen = torch.Generator().manual_seed(seed)
        self.device = torch.device(device)
        pin = self.device.type == "cuda"
        self._stage = [torch.empty(self.bs, b.shape[1]).pin_memory() if pin else torch.empty(self.bs, b.shape[1]) for b in self.blocks]
        self._lstage = torch.empty(self.bs, dtype=torch.int64).pin_memory() if pin else torch.empty(self.bs, dtype=torch.int64)

    def __len__(self):
        return self.n // self.bs if self.drop_last else (self.n + self.bs - 1) // self.bs

    def __iter__(self):
        order = torch.randperm(self.n, generator=self.gen) if self.shuffle else torch.arange(self.n)
        for i in range(len(self)):
            idx = order[i * self.bs:(i + 1) * self.bs]
            k = idx.numel()
            out = []
            for b, st in zip(self.blocks, self._stage):
                torch.index_select(b, 0, idx, out=st[:k])
                out.append(st[:k].to(self.device, non_blocking=True))
            torch.index_select(self.labels, 0, idx, out=self._lstage[:k])
            lab = self._lstage[:k].to(self.device, non_blocking=True)
            if self.device.type == "cuda":
                torch.cuda.current_stream().synchronize()  # staging buffers are reused by the next batch
            yield out, lab
