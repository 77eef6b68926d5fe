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


def read_reference_split(eeg_csv: str, action_pickle: str, eeg_pickle: str):
    """The reference's own dataset files as the tensors `MultiModalDataset_ti.__getitem__` yields (data.py:7-35), for the
    whole split at once: dict(frame_input [N,1,512] f32, vedio_mask [N,1] i64, title_input [N,512] i64,
    text_mask [N,512] i64, label [N] i64).  A null label becomes 0, as at data.py:31-32.  Files:
    feature/{train,test}_EEG.csv, feature/action/*_clip_v2.pickle (ndarray [N,512]), feature/EEG/*_bert.pickle (list of
    tokenizer encodings with 'input_ids' / 'attention_mask')."""
    import csv
    import pickle

    with open(eeg_csv, newline="") as f:
        rows = list(csv.DictReader(f))
    labels = []
    for r in rows:
        v = (r.get("label") or "").strip()
        labels.append(0 if v == "" or v.lower() == "nan" else int(float(v)))
    with open(action_pickle, "rb") as f:
        act = pickle.load(f)
    with open(eeg_pickle, "rb") as f:
        enc = pickle.load(f)
    n = len(rows)
    if len(act) < n or len(enc) < n:
        raise ValueError(f"{eeg_csv} lists {n} samples but the pickles hold {len(act)} / {len(enc)}")
    frame = torch.from_numpy(np.ascontiguousarray(np.asarray(act[:n], dtype=np.float32))).reshape(n, 1, -1)
    ids = torch.tensor([list(e["input_ids"]) for e in enc[:n]], dtype=torch.int64)
    mask = torch.tensor([list(e["attention_mask"]) for e in enc[:n]], dtype=torch.int64)
    return dict(frame_input=frame, vedio_mask=torch.ones(n, 1, dtype=torch.int64), title_input=ids, text_mask=mask,
                label=torch.tensor(labels, dtype=torch.int64))


@torch.no_grad()
def convert_reference_split(eeg_csv: str, action_pickle: str, eeg_pickle: str, encoder, out_path: str, batch_size: int = 64,
                            device="cpu"):
    """Reference dataset files -> feature cache.  `encoder` is the (frozen) stack above the head -- the reference's BERT,
    visual projection and cross-attention (model.py:17-21,34-46), or any callable with the same contract: it takes the
    4-tuple (frame_input, vedio_mask, title_input, text_mask) of data.py:22-35 and returns the feature blocks
    ([B,768] x 3 in the reference) BEFORE the concat + min-max normalisation, which belong to the head.  The encoders are
    out of this repo's scope; this function is the bridge from the reference's files to the head's on-disk input."""
    raw = read_reference_split(eeg_csv, action_pickle, eeg_pickle)
    n = raw["label"].shape[0]
    chunks = None
    for lo in range(0, n, batch_size):
        x = tuple(raw[k][lo:lo + batch_size].to(device) for k in ("frame_input", "vedio_mask", "title_input", "text_mask"))
        blocks = encoder(x)
        blocks = [blocks] if isinstance(blocks, torch.Tensor) else list(blocks)
        blocks = [b.reshape(b.shape[0], -1).float().cpu() for b in blocks]
        if chunks is None:
            chunks = [[] for _ in blocks]
        for c, b in zip(chunks, blocks):
            c.append(b)
    blocks = [torch.cat(c) for c in chunks]
    save_features(out_path, [b.numpy() for b in blocks], raw["label"].numpy())
    return blocks, raw["label"]


class ResidentDataset:
    """A split resident in HBM: feature blocks [N,d_i] fp32 and labels [N] int64 on the device, plus the epoch's row
    order as a DEVICE permutation.  This is what the fused sweep step gathers its batches from (no host work per batch),
    and what the large-batch path index-selects from on the device.  `epoch_order()` draws the permutation with the same
    CPU generator sequence as FeatureLoader / torch's DataLoader(shuffle=True) would (data.py:41-42), so a run is
    reproducible and comparable with the host-staged loader."""

    def __init__(self, blocks, labels, batch_size, shuffle=True, seed=980616, device="cuda"):
        self.device = torch.device(device)
        self.blocks = [b.contiguous().float().to(self.device) for b in blocks]
        self.labels = labels.reshape(-1).to(torch.int64).contiguous().to(self.device)
        self.n, self.bs, self.shuffle = int(self.labels.shape[0]), int(batch_size), bool(shuffle)
        self.gen = torch.Generator().manual_seed(seed)
        self.order = torch.arange(self.n, device=self.device)

    def __len__(self):
        return (self.n + self.bs - 1) // self.bs

    @property
    def n_full(self):
        return self.n // self.bs

    def epoch_order(self):
        """Draw this epoch's row order (one small H2D copy per epoch) into the persistent device tensor `order`."""
        if self.shuffle:
            self.order.copy_(torch.randperm(self.n, generator=self.gen), non_blocking=True)
        return self.order

    def batch(self, i):
        """Batch i of the current order, gathered ON the device (no host synchronisation)."""
        idx = self.order[i * self.bs:(i + 1) * self.bs]
        return [b.index_select(0, idx) for b in self.blocks], self.labels.index_select(0, idx)

    def __iter__(self):
        self.epoch_order()
        for i in range(len(self)):
            yield self.batch(i)


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
