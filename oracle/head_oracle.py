"""CPU oracle for the privatised fusion head (TEST INFRASTRUCTURE, not product).

This file is a dimension-parametrised restatement of the reference's hot path so the
CUDA kernels can be checked against it on the same seeded inputs.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it; nothing under `eeg_multimodal_b200/` does.

Reference lines restated here (paths relative to the reference checkout, CRLF files):

* head forward ........ python/src/custom_models/models.py:69-82  (== past_acc.py:120-138)
* loss / accuracy ..... python/src/custom_models/base_train.py:59-65 (== past_acc.py:71-77)
* two-pass step ....... past_acc.py:155-160,198-212 (== base_train.py:183-210)
* unfixed eps_hat ..... model.py:57 (commented-out variant, `new_*eps` runs)
* Laplace sampling .... torch.distributions.Laplace.rsample (torch 2.11): u~U(eps32-1,1),
                        -sign(u)*log1p(-|u|)
* Gumbel gate ......... torch.nn.functional.gumbel_softmax (torch 2.11)

The arithmetic is torch's own (installed here, torch 2.11.0), so the restatement calls
torch CPU ops in the reference's order; the only change is that the two random draws are
INJECTED (`lap_noise`, `gumbel`) instead of drawn inside the function, so the same
tensors can be fed to the CUDA path.

Parity pin: `tests/golden/head_golden.npz` was produced by running the UNMODIFIED
reference class `TICA_LapDropout.forward` (encoders stubbed, RNG replayed, see
`oracle/ref_shim.py` + `tests/golden/make_golden.py`) and this restatement is asserted
bit-identical to it at D=2304 (`tests/test_oracle.py`).  The reference has no tests of its
own for this path, so those generated vectors are the pin.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

F32_EPS = float(torch.finfo(torch.float32).eps)


# --------------------------------------------------------------------------------------
# random draws, restated so they can be replayed / injected
# --------------------------------------------------------------------------------------
def laplace_from_uniform(u: torch.Tensor) -> torch.Tensor:
    """torch.distributions.Laplace(0,1).rsample given its uniform draw `u` in (eps-1, 1).

    Reference call site: models.py:74 `self.noiser.sample(feature.shape)`.
    """
    loc = torch.zeros((), dtype=u.dtype)
    scale = torch.ones((), dtype=u.dtype)
    return loc - scale * u.sign() * torch.log1p(-u.abs())


def replay_reference_draws(seed: int, B: int, D: int):
    """Replay the two global-RNG draws one reference forward makes on CPU, in order
    (SURVEY.md section 8c): (1) uniform_(eps-1, 1) on [B,D,1] for the Laplace noise,
    (2) exponential_() on [2,B,D] for the Gumbel noise.  Returns (lap_noise[B,D],
    gumbel[2,B,D]) where gumbel = -log(Exp(1))."""
    torch.manual_seed(seed)
    u = torch.empty(B, D, 1).uniform_(F32_EPS - 1, 1)
    lap = laplace_from_uniform(u).view(B, D)
    e = torch.empty(2, B, D).exponential_()
    gum = -e.log()
    return lap, gum


# --------------------------------------------------------------------------------------
# the head, models.py:69-82
# --------------------------------------------------------------------------------------
@dataclass
class HeadParams:
    W1: torch.Tensor  # fc_layers.0.weight [D,D]
    b1: torch.Tensor  # fc_layers.0.bias   [D]
    W2: torch.Tensor  # fc_layers.2.weight [H,D]
    b2: torch.Tensor  # fc_layers.2.bias   [H]
    Wc: torch.Tensor  # classifier.weight  [C,H]
    bc: torch.Tensor  # classifier.bias    [C]
    DP: torch.Tensor  # DP                 [1,D]

    def tensors(self):
        return [self.W1, self.b1, self.W2, self.b2, self.Wc, self.bc, self.DP]

    def clone(self, requires_grad=False):
        return HeadParams(*[t.detach().clone().requires_grad_(requires_grad) for t in self.tensors()])


def make_params(D: int, H: int = 768, C: int = 2, seed: int = 0, dp: np.ndarray | None = None,
                dtype=torch.float32) -> HeadParams:
    """Deterministic head parameters from numpy's PCG64 (stable across numpy versions), with
    nn.Linear's U(-1/sqrt(fan_in), 1/sqrt(fan_in)) scale.  Used by tests and golden vectors so
    the 7M-parameter head never has to be stored in a fixture."""
    rng = np.random.default_rng(seed)

    def lin(out_f, in_f):
        k = 1.0 / math.sqrt(in_f)
        w = rng.uniform(-k, k, size=(out_f, in_f)).astype(np.float32)
        b = rng.uniform(-k, k, size=(out_f,)).astype(np.float32)
        return torch.from_numpy(w).to(dtype), torch.from_numpy(b).to(dtype)

    W1, b1 = lin(D, D)
    W2, b2 = lin(H, D)
    Wc, bc = lin(C, H)
    if dp is None:
        DP = torch.zeros(1, D, dtype=dtype)
    else:
        DP = torch.as_tensor(np.asarray(dp, dtype=np.float32)).reshape(1, D).to(dtype)
    return HeadParams(W1, b1, W2, b2, Wc, bc, DP)


def eps_tensor(epsilon) -> torch.Tensor:
    """models.py:58 `eps = torch.tensor(epsilon)`: python float -> fp32 0-dim tensor,
    np.float64 -> fp64 0-dim tensor (eps_hat conditioning)."""
    return torch.tensor(epsilon)


def exp_eps_f32(epsilon) -> float:
    """e^eps as the fp32 value that enters `(eps.exp() - w)` at models.py:75.  A 0-dim
    fp64 tensor is demoted to fp32 by type promotion against the fp32 `w`."""
    return float(eps_tensor(epsilon).exp().to(torch.float32))


def minmax_normalise(feature_concat: torch.Tensor) -> torch.Tensor:
    """models.py:70-72.  No epsilon guard: a constant row gives NaN, like the reference."""
    feature_min = torch.min(feature_concat, dim=-1, keepdims=True)[0]
    feature_max = torch.max(feature_concat, dim=-1, keepdims=True)[0]
    return (feature_concat - feature_min) / (feature_max - feature_min)


def eps_hat_of(w: torch.Tensor, eps: torch.Tensor, fixed: bool = True) -> torch.Tensor:
    """models.py:75 (fixed, `newfrac_*` runs) or model.py:57 (unfixed, `new_*` runs)."""
    if fixed:
        return 1 / (((eps.exp() - w) / (1 - w)).log())
    return ((eps.exp() - w) / (1 - w)).log()


def gumbel_mask(w: torch.Tensor, B: int, gumbel: torch.Tensor, hard: bool, tau: float = 1.0):
    """F.gumbel_softmax(torch.stack((w, 1-w)).repeat(1,B,1), hard=hard, dim=0) with the
    Gumbel draw injected (models.py:77-78).  Returns (mask[2,B,D], index[B,D])."""
    logits = torch.stack((w, 1 - w)).repeat(1, B, 1)  # [2,B,D]
    gumbels = (logits + gumbel) / tau
    y_soft = gumbels.softmax(0)
    index = y_soft.max(0, keepdim=True)[1]
    if hard:
        y_hard = torch.zeros_like(logits).scatter_(0, index, 1.0)
        ret = y_hard - y_soft.detach() + y_soft
    else:
        ret = y_soft
    return ret, index[0]


def head_forward(blocks, p: HeadParams, epsilon, lap_noise: torch.Tensor, gumbel: torch.Tensor,
                 hard: bool, tau: float = 1.0, fixed: bool = True, return_aux: bool = False):
    """models.py:69-82 on injected feature blocks.  `blocks` is a list of [B,Di] tensors
    (reference: eeg pooled, action projected, cross-attention; synthetic: eeg, action)."""
    eps = epsilon if isinstance(epsilon, torch.Tensor) else eps_tensor(epsilon)
    feature_concat = torch.cat(tuple(blocks), dim=1)                      # :69
    feature = minmax_normalise(feature_concat)                            # :70-72
    w = F.sigmoid(p.DP)                                                   # :73
    noise = lap_noise.view(*feature.shape)                                # :74
    eps_hat = eps_hat_of(w, eps, fixed)                                   # :75
    feature_p = feature + noise * eps_hat                                 # :76
    mask, index = gumbel_mask(w, feature_p.shape[0], gumbel, hard, tau)   # :77-78
    gated = (feature_p * mask).sum(0)                                     # :79
    h = F.linear(gated, p.W1, p.b1).relu()                                # :80 fc_layers.0/1
    h2 = F.linear(h, p.W2, p.b2).tanh()                                   # :80 fc_layers.2/3
    prediction = F.linear(h2, p.Wc, p.bc)                                 # :81
    if return_aux:
        return prediction, dict(feature=feature, perturbed=feature_p, gated=gated,
                                gate_index=index, eps_hat=eps_hat, w=w, h1=h, h2=h2)
    return prediction


def head_forward_nonprivate(blocks, p: HeadParams):
    """model.py:53-64 (privacy block commented out): normalise -> fc_layers -> classifier."""
    feature = minmax_normalise(torch.cat(tuple(blocks), dim=1))
    h = F.linear(feature, p.W1, p.b1).relu()
    h2 = F.linear(h, p.W2, p.b2).tanh()
    return F.linear(h2, p.Wc, p.bc)


def cal_loss(prediction: torch.Tensor, label: torch.Tensor):
    """base_train.py:59-65.  label is [B,1] int64."""
    label = label.squeeze(dim=1)
    loss = F.cross_entropy(prediction, label)
    with torch.no_grad():
        pred_label_id = torch.argmax(prediction, dim=1)
        accuracy = (label == pred_label_id).float().sum() / label.shape[0]
    return loss, accuracy, pred_label_id, label


# --------------------------------------------------------------------------------------
# two-pass training step, past_acc.py:155-160,198-212
# --------------------------------------------------------------------------------------
class TwoPassTrainer:
    """Adam(DP, lr) on a hard=False pass, then Adam(rest, lr) on a hard=True pass, fresh noise
    each pass, torch.optim.Adam defaults (betas .9/.999, eps 1e-8)."""

    def __init__(self, p: HeadParams, epsilon, lr: float = 1e-6, fixed: bool = True, tau: float = 1.0):
        self.p = p.clone(requires_grad=True)
        self.epsilon, self.fixed, self.tau = epsilon, fixed, tau
        self.dp_opt = torch.optim.Adam([self.p.DP], lr=lr)
        self.model_opt = torch.optim.Adam([self.p.W1, self.p.b1, self.p.W2, self.p.b2, self.p.Wc, self.p.bc], lr=lr)

    def step(self, blocks, label, noise1, gum1, noise2, gum2):
        self.dp_opt.zero_grad()                                            # past_acc.py:198
        pred = head_forward(blocks, self.p, self.epsilon, noise1, gum1, hard=False, tau=self.tau, fixed=self.fixed)
        loss, acc, _, _ = cal_loss(pred, label)                            # :201
        loss.backward()                                                    # :202
        self.dp_opt.step()                                                 # :203
        self.model_opt.zero_grad()                                         # :206
        pred = head_forward(blocks, self.p, self.epsilon, noise2, gum2, hard=True, tau=self.tau, fixed=self.fixed)
        loss, acc, pred_id, lab = cal_loss(pred, label)                    # :208
        loss.backward()                                                    # :211
        self.model_opt.step()                                              # :212
        return float(loss.detach()), float(acc), pred_id


def reference_step_cpu(blocks, label, p: HeadParams, epsilon, hard_noise_on_host: bool = True):
    """One reference-style two-pass fwd+bwd with the reference's OWN noise calls (host
    Laplace.sample + F.gumbel_softmax), used as the timed CPU baseline."""
    noiser = torch.distributions.laplace.Laplace(torch.tensor([0.0]), torch.tensor([1.0]))
    eps = eps_tensor(epsilon)
    out = None
    for hard in (False, True):
        for t in p.tensors():
            t.grad = None
        feature = minmax_normalise(torch.cat(tuple(blocks), dim=1))
        w = F.sigmoid(p.DP)
        noise = noiser.sample(feature.shape).view(*feature.shape)
        eps_hat = 1 / (((eps.exp() - w) / (1 - w)).log())
        feature = feature + noise * eps_hat
        mask = F.gumbel_softmax(torch.stack((w, 1 - w)).repeat(1, feature.shape[0], 1), hard=hard, dim=0)
        feature = (feature * mask).sum(0)
        h2 = F.linear(F.linear(feature, p.W1, p.b1).relu(), p.W2, p.b2).tanh()
        pred = F.linear(h2, p.Wc, p.bc)
        loss, acc, _, _ = cal_loss(pred, label)
        loss.backward()
        out = (float(loss.detach()), float(acc))
    return out


# --------------------------------------------------------------------------------------
# mixed-precision restatement: the same path with the casts the tensor-core kernels make
# --------------------------------------------------------------------------------------
def _q(t: torch.Tensor) -> torch.Tensor:
    """round-to-nearest-even to bf16 and back (what a bf16 store + load does)"""
    return t.to(torch.bfloat16).to(torch.float32)


def head_fwd_bwd_bf16sim(blocks, p: HeadParams, epsilon, lap_noise, fixed: bool = True, grad_scale=None,
                         h2_bf16: bool = False):
    """models.py:69-82 + cal_loss + autograd, restated with bf16 rounding at exactly the points where
    the B200 bf16 path stores bf16 (perturbed features X, weights W1/W2, ReLU output H1, dZ2, dZ1);
    everything else (accumulation, Tanh output H2 unless h2_bf16, logits, loss, dX, all weight gradients) is fp32.
    The Gumbel gate is the identity (hard) / within 1 ulp (soft), see gumbel_mask, and is skipped.
    Needed because a ReLU network's gradient is discontinuous in its inputs: rounding X and W to
    bf16 flips the sign of a fraction p of pre-activations, which moves dZ1 by ~sqrt(p) in Frobenius
    norm however exact the GEMMs are; against THIS oracle the kernels must agree to the bf16 bar.
    Returns dict(logits, loss, grads...)."""
    eps = epsilon if isinstance(epsilon, torch.Tensor) else eps_tensor(epsilon)
    label = None
    feature = minmax_normalise(torch.cat(tuple(blocks), dim=1))
    w = F.sigmoid(p.DP.detach())
    eps_hat = eps_hat_of(w, eps, fixed)
    X = _q(feature + lap_noise * eps_hat)
    W1q, W2q = _q(p.W1.detach()), _q(p.W2.detach())
    H1 = _q(F.linear(X, W1q, p.b1.detach()).relu())
    H2 = F.linear(H1, W2q, p.b2.detach()).tanh()
    if h2_bf16:                      # HeadEngine.h2_bf16: the Tanh output is stored as bf16 too
        H2 = _q(H2)
    logits = F.linear(H2, p.Wc.detach(), p.bc.detach())

    def backward(label):
        B = logits.shape[0]
        lab = label.view(-1)
        gs = (1.0 / B) if grad_scale is None else grad_scale
        dlogits = (torch.softmax(logits, 1) - F.one_hot(lab, 2).float()) * gs
        loss = F.cross_entropy(logits, lab)
        dWc, dbc = dlogits.t() @ H2, dlogits.sum(0)
        dZ2 = _q((dlogits @ p.Wc.detach()) * (1 - H2 * H2))
        dW2, db2 = dZ2.t() @ H1, dZ2.sum(0)
        dZ1 = _q((dZ2 @ W2q) * (H1 > 0))
        dW1, db1 = dZ1.t() @ X, dZ1.sum(0)
        dX = dZ1 @ W1q
        E = eps.exp().to(torch.float32)
        L = ((E - w) / (1 - w)).log()
        coef = -(E - 1) * w / (L * L * (E - w)) if fixed else (E - 1) * w / (E - w)
        dDP = (dX * lap_noise).sum(0, keepdim=True) * coef
        return dict(loss=loss, dW1=dW1, db1=db1, dW2=dW2, db2=db2, dWc=dWc, dbc=dbc, dDP=dDP)

    return logits, backward
