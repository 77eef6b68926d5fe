"""The reference's OLDER "PriGumbel" head (SURVEY.md section 8 row a-alt), host side.

Mirrors the head part of `train_val.ConcatModel` (train_val.py:125-157) and its training step (train_val.py:178,
203-215) on the CUDA kernels of libpgfuse.so:

    x   = relu(fc1(feature_concat))                       pgf_linear_fwd (ReLU)
    x   = fc2(x)                                          pgf_linear_fwd
    res = x * gumbel_softmax([w,1-w], tau, hard)[:,1]/(1-w)   pgf_prigumbel_coef + pgf_prigumbel_fwd
    res = minmax_row(res) + Laplace(0, 1/eps) per row         (same launch)
    prediction = classifier(res)                          pgf_cls_ce (no tanh in front of this classifier)
    loss = alpha * CE + max_j((1-w_j) e^eps + w_j)         pgf_cls_ce + pgf_prigumbel_coef (train_val.py:80-93)

`state_dict` keys are the reference's (`fc1.*`, `fc2.*`, `classifier.*`, `w`); train mode uses the soft gate, eval mode
the hard gate (train_val.py:108-111); ONE Adam over every parameter, `w` included.  There is no CPU fallback.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib as L
from . import ops

REFERENCE_TAU, REFERENCE_LR = 0.01, 1e-5           # train_val.py:524,529


class PriGumbelHead:
    _KEYS = {"fc1.weight": "W1", "fc1.bias": "b1", "fc2.weight": "W2", "fc2.bias": "b2",
             "classifier.weight": "Wc", "classifier.bias": "bc", "w": "w"}

    def __init__(self, feature_dim=2304, hidden=768, tau=REFERENCE_TAU, epsilon=1.0, alpha=1.0, lr=REFERENCE_LR,
                 device="cuda", seed=980616, init_seed=980616, betas=(0.9, 0.999), adam_eps=1e-8):
        if not torch.cuda.is_available():
            raise RuntimeError("PriGumbelHead needs a CUDA device: there is no CPU fallback")
        L.load()
        self.D, self.H, self.C = int(feature_dim), int(hidden), 2
        self.tau, self.epsilon, self.alpha, self.lr = float(tau), float(epsilon), float(alpha), float(lr)
        self.exp_eps = float(np.float32(math.exp(float(epsilon))))   # np.exp(epsilon) in fp64, rounded once by the fp32 multiply (train_val.py:88)
        self.betas, self.adam_eps = betas, adam_eps
        self.seed, self.noise_offset, self.t = int(seed), 0, 0
        self.training = True
        self.device = torch.device(device)
        D, H, C = self.D, self.H, self.C
        self.layout, off = {}, 0
        for name, shape in (("W1", (D, D)), ("b1", (D,)), ("W2", (H, D)), ("b2", (H,)), ("Wc", (C, H)), ("bc", (C,)), ("w", (H,))):
            self.layout[name] = (off, shape)
            off += (math.prod(shape) + 7) // 8 * 8
        self.P = off
        dev = self.device
        self.flat, self.grad = torch.zeros(self.P, device=dev), torch.zeros(self.P, device=dev)
        self.m, self.v = torch.zeros(self.P, device=dev), torch.zeros(self.P, device=dev)
        self.coef = torch.empty(4, H, device=dev)
        self.wloss = torch.empty(2, device=dev)
        self._injected = None
        g = torch.Generator().manual_seed(int(init_seed))
        for wn, bn in (("W1", "b1"), ("W2", "b2"), ("Wc", "bc")):           # nn.Linear default init
            shape = self.layout[wn][1]
            k = 1.0 / math.sqrt(shape[1])
            self.view(wn).copy_((torch.rand(shape, generator=g) * 2 - 1) * k)
            self.view(bn).copy_((torch.rand(shape[0], generator=g) * 2 - 1) * k)
        self.view("w").copy_(torch.rand(H, generator=g))                     # train_val.py:135

    # ---- parameters ----------------------------------------------------------------------------
    def view(self, name, src=None):
        off, shape = self.layout[name]
        return (self.flat if src is None else src)[off:off + math.prod(shape)].view(*shape)

    def state_dict(self):
        return {k: self.view(n).detach().clone() for k, n in self._KEYS.items()}

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self._KEYS if k not in sd]
        if strict and missing:
            raise KeyError(f"missing keys {missing}")
        for k, n in self._KEYS.items():
            if k in sd:
                self.view(n).copy_(sd[k].reshape(self.layout[n][1]))
        return missing

    def train(self, mode=True):
        self.training = bool(mode)
        return self

    def eval(self):
        return self.train(False)

    def inject_noise(self, gumbel, lap):
        """Draws for the NEXT forward: gumbel [H,2], lap [B] = Laplace(0,1/eps) (parity tests)."""
        self._injected = (gumbel.contiguous().to(self.device), lap.contiguous().to(self.device))

    def privacy_stats(self):
        """The per-epoch figures of train_val.py:222-226: privacy_budget_max/avg, drop_out_rate_max/avg."""
        w = self.view("w")
        tmp = (1 - w) * self.exp_eps + w
        return dict(privacy_budget_max=float(tmp.max()), privacy_budget_avg=float(tmp.mean()),
                    drop_out_rate_max=float(w.max()), drop_out_rate_avg=float(w.mean()))

    # ---- forward / loss / step -------------------------------------------------------------------
    def _features(self, x):
        if isinstance(x, (list, tuple)):
            x = torch.cat(tuple(x), dim=1)                                     # train_val.py:150
        if x.dtype != torch.float32 or x.dim() != 2 or x.shape[1] != self.D:
            raise ValueError(f"expected fp32 features [B,{self.D}]")
        return x.contiguous()

    def _forward(self, x, labels, backward):
        v = self.view
        B = x.shape[0]
        gum = lap = None
        if self._injected is not None:
            gum, lap = self._injected
            self._injected = None
        self.noise_offset += 1
        h1 = ops.linear_fwd(x, v("W1"), v("b1"), act=L.ACT_RELU)
        z = ops.linear_fwd(h1, v("W2"), v("b2"), act=L.ACT_NONE)
        ops.prigumbel_coef(v("w"), exp_eps=self.exp_eps, tau=self.tau, hard=not self.training, gumbel=gum, seed=self.seed,
                           offset=self.noise_offset, coef=self.coef, wloss=self.wloss)
        res = ops.prigumbel_fwd(z, self.coef, eps=self.epsilon, lap=lap, seed=self.seed, offset=self.noise_offset)
        g = self.grad
        ce = ops.cls_ce(res, v("Wc"), v("bc"), labels, loss_scale=1.0 / B, grad_scale=self.alpha / B, backward=backward,
                        through_tanh=False, dWc=v("Wc", g) if backward else None, dbc=v("bc", g) if backward else None)
        return h1, z, ce

    def forward(self, x):
        """prediction [B,2] of train_val.py:157 (soft gate in train mode, hard in eval mode)."""
        with ops.stream_scope():
            x = self._features(x)
            return self._forward(x, None, False)[2]["logits"]

    __call__ = forward

    def _result(self, ce):
        st = ce["stats"].tolist()                                              # {mean CE, n_correct, ., B}
        total = self.alpha * st[0] + float(self.wloss[0])                       # train_val.py:90
        return dict(loss=total, ce=st[0], acc=st[1] / st[3], pred=ce["pred"], logits=ce["logits"])

    def eval_step(self, x, labels):
        """One evaluation batch of train_val.py:229-241 (loss_function included)."""
        with ops.stream_scope():
            x = self._features(x)
            return self._result(self._forward(x, labels.reshape(-1).contiguous(), False)[2])

    def train_step(self, x, labels):
        """train_val.py:206-215: zero_grad, forward, loss_function, backward, Adam.step over every parameter."""
        with ops.stream_scope():
            x = self._features(x)
            v, g = self.view, self.grad
            h1, z, ce = self._forward(x, labels.reshape(-1).contiguous(), True)
            dz, _ = ops.prigumbel_bwd(z, self.coef, ce["dz"], wloss=self.wloss, exp_eps=self.exp_eps, wloss_scale=1.0, dw=v("w", g))
            ops.linear_bwd_dw(dz, h1, dW=v("W2", g), db=v("b2", g))
            dh1 = ops.linear_bwd_dx(dz, v("W2"), mask_src=h1, mask_mode=L.ACT_RELU)
            ops.linear_bwd_dw(dh1, x, dW=v("W1", g), db=v("b1", g))
            self.t += 1
            ops.adam_step(self.flat, g, self.m, self.v, self.t, lr=self.lr, betas=self.betas, eps=self.adam_eps)
            return self._result(ce)
