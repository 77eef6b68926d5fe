"""CPU oracle for the OLDER "PriGumbel" head of the reference (SURVEY.md section 8 row a-alt).
TEST INFRASTRUCTURE, not product: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs
may import it; nothing under `eeg_multimodal_b200/` does.

Reference lines restated here (train_val.py at the root of the reference checkout):

* gumbel_dropout ................ train_val.py:95-101   x * gumbel_softmax([w,1-w], tau, hard)[:,1] / (1-w);
                                                       w [H] is SHARED over the batch, one [H,2] draw per forward
* GumbelSoftmaxDropout.forward .. train_val.py:107-112  soft in train mode, hard in eval mode
* Lap_noise ..................... train_val.py:114-123  row min-max, then ONE Laplace(0, 1/eps) scalar per row
* ConcatModel.forward (head) .... train_val.py:151-157  relu(fc1) -> fc2 -> dropout -> Lap_noise -> classifier
* loss_function ................. train_val.py:80-93    alpha * CE + max_j((1-w_j) * e^eps + w_j)
* optimiser ..................... train_val.py:178       ONE Adam over all parameters (w included)

The arithmetic is torch's own, called in the reference's order; the two random draws (`exponential_` on
[H,2] inside F.gumbel_softmax, then `uniform_` on [B,1] inside Laplace.rsample -- in that order, both from the
global CPU generator when the model sits on the CPU) are INJECTED so the CUDA path can be fed the same values.

Parity pin: `tests/golden/prigumbel_golden.npz`, produced by `tests/golden/make_golden_prigumbel.py` from the
UNMODIFIED reference functions `gumbel_dropout`, `Lap_noise`, `loss_function` imported from train_val.py
(opacus / transformers.AdamW stubbed outside the file) with the RNG replayed; this restatement is asserted
bit-identical to those outputs and gradients (`tests/test_oracle.py`).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

from .head_oracle import F32_EPS


def replay_reference_draws(seed: int, B: int, H: int, epsilon: float):
    """The two draws one reference forward makes (train_val.py:155-156), in order: (1) exponential_() on [H,2]
    (F.gumbel_softmax), (2) Laplace(0, 1/eps).sample([B]) = uniform_(eps32-1, 1) on [B,1].  Returns
    (gumbel [H,2] = -log(Exp(1)), lap [B] already scaled by 1/eps as the reference's distribution is)."""
    torch.manual_seed(seed)
    gum = -torch.empty(H, 2).exponential_().log()
    loc, scale = torch.tensor([0.0]), torch.tensor([1 / epsilon])
    u = torch.empty(B, 1).uniform_(F32_EPS - 1, 1)
    lap = loc - scale * u.sign() * torch.log1p(-u.abs())        # torch.distributions.Laplace.rsample
    return gum, lap.view(B)


@dataclass
class PriGumbelParams:
    W1: torch.Tensor  # fc1.weight        [D,D]
    b1: torch.Tensor  # fc1.bias          [D]
    W2: torch.Tensor  # fc2.weight        [H,D]
    b2: torch.Tensor  # fc2.bias          [H]
    Wc: torch.Tensor  # classifier.weight [2,H]
    bc: torch.Tensor  # classifier.bias   [2]
    w: torch.Tensor   # w                 [H]  (torch.rand(768) in the reference, train_val.py:135)

    def tensors(self):
        return [self.W1, self.b1, self.W2, self.b2, self.Wc, self.bc, self.w]

    def clone(self, requires_grad=False):
        return PriGumbelParams(*[t.detach().clone().requires_grad_(requires_grad) for t in self.tensors()])


def make_params(D: int, H: int = 768, seed: int = 0) -> PriGumbelParams:
    """nn.Linear-style U(-1/sqrt(in), 1/sqrt(in)) weights from numpy PCG64 (reproducible without torch's RNG);
    w ~ U(0.05, 0.95) (the reference's torch.rand, kept off the poles of 1/(1-w))."""
    rng = np.random.default_rng(seed)

    def lin(o, i):
        k = 1 / np.sqrt(i)
        return (torch.tensor(rng.uniform(-k, k, (o, i)).astype(np.float32)),
                torch.tensor(rng.uniform(-k, k, (o,)).astype(np.float32)))

    W1, b1 = lin(D, D)
    W2, b2 = lin(H, D)
    Wc, bc = lin(2, H)
    w = torch.tensor(rng.uniform(0.05, 0.95, (H,)).astype(np.float32))
    return PriGumbelParams(W1, b1, W2, b2, Wc, bc, w)


def gumbel_dropout(x: torch.Tensor, w: torch.Tensor, gumbel: torch.Tensor, tau: float, hard: bool):
    """train_val.py:95-101 with F.gumbel_softmax (torch 2.11) unrolled around the injected draw."""
    w_tensor = w.unsqueeze(1)
    w_tensor = torch.cat([w_tensor, 1 - w_tensor], dim=1)                   # [H,2], used as LOGITS by the reference
    y_soft = ((w_tensor + gumbel) / tau).softmax(-1)
    if hard:
        index = y_soft.max(-1, keepdim=True)[1]
        y_hard = torch.zeros_like(w_tensor).scatter_(-1, index, 1.0)
        ret = y_hard - y_soft.detach() + y_soft
    else:
        ret = y_soft
    mask = ret[:, 1]
    return x * mask / (1.0 - w)


def lap_noise(x: torch.Tensor, lap: torch.Tensor):
    """train_val.py:114-123; `lap` [B] are the Laplace(0, 1/eps) draws."""
    mn = torch.min(x, dim=-1, keepdim=True)[0]
    mx = torch.max(x, dim=-1, keepdim=True)[0]
    pooled = (x - mn) / (mx - mn)
    pooled = pooled + lap.view(-1, 1)                                        # the reference's in-place +=
    return pooled


def head_forward(feature_concat: torch.Tensor, p: PriGumbelParams, tau: float, hard: bool, gumbel, lap):
    """train_val.py:151-157 from the concatenated [B,D] features on."""
    x = F.relu(F.linear(feature_concat, p.W1, p.b1))
    x = F.linear(x, p.W2, p.b2)
    res = gumbel_dropout(x, p.w, gumbel, tau, hard)
    res_lap = lap_noise(res, lap)
    return F.linear(res_lap, p.Wc, p.bc)


def loss_function(prediction: torch.Tensor, label: torch.Tensor, w: torch.Tensor, alpha: float, epsilon: float):
    """train_val.py:80-93.  label is [B,1] int64."""
    label = label.squeeze(dim=1)
    cross_entropy_loss = F.cross_entropy(prediction, label)
    with torch.no_grad():
        pred_label_id = torch.argmax(prediction, dim=1)
        accuracy = (label == pred_label_id).float().sum() / label.shape[0]
    tmp = (1 - w) * np.exp(epsilon) + w
    loss_w, _ = torch.max(tmp, dim=0)
    total_loss = alpha * cross_entropy_loss + loss_w
    return total_loss, accuracy, pred_label_id, label


class PriGumbelTrainer:
    """train_val.py:178,203-215: one Adam over every parameter, soft gate (train mode), fresh draws per step."""

    def __init__(self, p: PriGumbelParams, epsilon: float, tau: float, alpha: float, lr: float):
        self.p = p.clone(requires_grad=True)
        self.epsilon, self.tau, self.alpha = epsilon, tau, alpha
        self.opt = torch.optim.Adam(self.p.tensors(), lr=lr)

    def step(self, feature_concat, label, gumbel, lap):
        self.opt.zero_grad()                                                 # train_val.py:206
        pred = head_forward(feature_concat, self.p, self.tau, False, gumbel, lap)             # :208
        loss, acc, pred_id, _ = loss_function(pred, label, self.p.w, self.alpha, self.epsilon)  # :210
        loss.backward()                                                      # :214
        self.opt.step()                                                      # :215
        return float(loss.detach()), float(acc), pred_id
