"""Drop-in host-side mirror of the reference's `model.py` for the privatised fusion head.

Same names, call shapes and `state_dict` keys as the reference:

    get_model(cfg)                           reference model.py:8-12
    ConcatModel.feature(x) -> [B,D]          reference model.py:34-51
    ConcatModel.forward(x, hard=True)        reference model.py:53-64, with the privacy block that
                                             is commented out there and live at past_acc.py:130-136
                                             == python/src/custom_models/models.py:73-79
    state_dict keys: fc_layers.0.{weight,bias}, fc_layers.2.{weight,bias},
                     classifier.{weight,bias}, DP        

All arithmetic runs in the sm_100a kernels behind libpgfuse.so; this file only owns the
parameters, the autograd glue and the noise bookkeeping.  The encoders above the head (BERT,
visual projection, cross-attention: reference model.py:17-21) are out of scope: `x` is the tuple
of pre-extracted feature blocks ([B,768] x3 in the reference, [B,2048]+[B,512] synthetic), or
the reference's 4-tuple if an `encoder` module producing those blocks is plugged in.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from . import ops

REFERENCE_SEED = 980616  # past_acc.py:33, train.py:25


def exp_eps_of(eps) -> float:
    """e^eps as the fp32 scalar the reference feeds into `(eps.exp() - w)` (models.py:58,75)."""
    t = eps if isinstance(eps, torch.Tensor) else torch.tensor(eps)
    return float(t.detach().cpu().exp().to(torch.float32))


class _NoiseCfg:
    __slots__ = ("mode", "lap", "gum", "seed", "offset", "row0", "tau", "hard", "fixed", "want_gate")


class _PerturbGateFn(torch.autograd.Function):
    """models.py:69-79 forward; backward = dDP (models.py:75-76) and, if the blocks need grad, the
    gradient through the min-max normalisation (models.py:70-72)."""

    @staticmethod
    def forward(ctx, DP, exp_eps, cfg, *blocks):
        w, eps_hat, deps = ops.dp_coeffs(DP.detach().reshape(-1).contiguous(), exp_eps, cfg.fixed)
        out, gate_idx, _, _ = ops.perturb_gate_fwd(
            [b.detach() for b in blocks], w, eps_hat, noise_mode=cfg.mode, lap=cfg.lap, gum=cfg.gum, seed=cfg.seed,
            offset=cfg.offset, row0=cfg.row0, tau=cfg.tau, hard=cfg.hard, want_gate=cfg.want_gate,
            want_gate_idx=cfg.want_gate)
        ctx.cfg, ctx.deps, ctx.dp_shape = cfg, deps, DP.shape
        ctx.blocks = blocks if any(b.requires_grad for b in blocks) else None
        ctx.mark_non_differentiable(*([gate_idx] if gate_idx is not None else []))
        if gate_idx is not None:
            return out, gate_idx
        return out

    @staticmethod
    def backward(ctx, dout, *unused):
        cfg = ctx.cfg
        dout = dout.contiguous()
        dDP = None
        if ctx.needs_input_grad[0]:
            dDP = ops.perturb_gate_bwd_dp(dout, ctx.deps, noise_mode=cfg.mode, lap=cfg.lap, seed=cfg.seed,
                                          offset=cfg.offset, row0=cfg.row0).view(ctx.dp_shape)
        dblocks = [None] * (len(ctx.needs_input_grad) - 3)
        if ctx.blocks is not None:
            dblocks = ops.minmax_norm_bwd([b.detach() for b in ctx.blocks], dout)
        return (dDP, None, None, *dblocks)


class _NormaliseFn(torch.autograd.Function):
    """Non-private path: concat + row min-max normalise (reference model.py:47-50)."""

    @staticmethod
    def forward(ctx, *blocks):
        out, _, _, _ = ops.perturb_gate_fwd([b.detach() for b in blocks], None, None, noise_mode=L.NOISE_NONE)
        ctx.blocks = blocks if any(b.requires_grad for b in blocks) else None
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.blocks is None:
            return tuple([None] * len(ctx.needs_input_grad))
        return tuple(ops.minmax_norm_bwd([b.detach() for b in ctx.blocks], dout.contiguous()))


class _HeadMLPFn(torch.autograd.Function):
    """fc_layers + classifier (models.py:80-81) on the fp32 CUDA-core kernels, with autograd."""

    @staticmethod
    def forward(ctx, X, W1, b1, W2, b2, Wc, bc):
        X = X.contiguous()
        H1 = ops.linear_fwd(X, W1.detach(), b1.detach(), L.ACT_RELU)
        H2 = ops.linear_fwd(H1, W2.detach(), b2.detach(), L.ACT_TANH)
        logits = ops.linear_fwd(H2, Wc.detach(), bc.detach(), L.ACT_NONE)
        ctx.save_for_backward(X, H1, H2, W1, W2, Wc)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        X, H1, H2, W1, W2, Wc = ctx.saved_tensors
        dlogits = dlogits.contiguous()
        need = ctx.needs_input_grad
        dWc, dbc = ops.linear_bwd_dw(dlogits, H2)
        dZ2 = ops.linear_bwd_dx(dlogits, Wc.detach(), mask_src=H2, mask_mode=L.ACT_TANH)
        dW2, db2 = ops.linear_bwd_dw(dZ2, H1)
        dZ1 = ops.linear_bwd_dx(dZ2, W2.detach(), mask_src=H1, mask_mode=L.ACT_RELU)
        dW1, db1 = ops.linear_bwd_dw(dZ1, X)
        dX = ops.linear_bwd_dx(dZ1, W1.detach()) if need[0] else None
        return dX, dW1, db1, dW2, db2, dWc, dbc


X3_CHAIN_KB = 10   # k-blocks of 64 per TMEM accumulation chain of the fp32x3 GEMMs (see HeadEngine.x3_chain_kb)


def _gemm_x3(A3, B3, C, *, K, zeroed=False, **kw):
    """C = A . B^T on the fp32x3 route, the contraction cut into K slabs of ~X3_CHAIN_KB k-blocks (tensor-core
    accumulation truncates; slabs are summed by round-to-nearest fp32 adds)."""
    ns = max(1, round(((K + 63) // 64) / X3_CHAIN_KB))
    if ns > 1 or zeroed:
        if not zeroed:
            ops.fill_zero(C)
        return ops.gemm_bf16x3(A3, B3, C, K=K, epi=L.EPI_ATOMIC_F32, k_slabs=max(ns, 2) if ns > 1 else 1, **kw)
    return ops.gemm_bf16x3(A3, B3, C, K=K, epi=L.EPI_STORE_F32, **kw)


class _HeadMLPTensorFn(torch.autograd.Function):
    """fc_layers + classifier (models.py:80-81) with the two dense contractions on the tensor cores, for batches that are a
    real GEMM.  precision 'fp32x3': fp32 arithmetic (hi/mid/lo bf16 planes of both operands, six plane-pair products, fp32
    accumulate: the 1e-5 bar); 'bf16': bf16 operands, fp32 accumulate, fused epilogues (the 2e-2 bar).  The 768 -> 2
    classifier stays on the CUDA-core kernels in fp32 either way."""

    @staticmethod
    def forward(ctx, X, W1, b1, W2, b2, Wc, bc, precision):
        X = X.contiguous()
        B, D = X.shape
        H = W2.shape[0]
        dev, bf = X.device, torch.bfloat16
        exact = precision == "fp32x3"
        H2 = torch.empty(B, H, device=dev)
        if exact:
            mk = lambda t: ops.split3(t.detach().contiguous(), planes=torch.empty(3, *t.shape, device=dev, dtype=bf))
            Xp, W1p, W2p = mk(X), mk(W1), mk(W2)
            Z = torch.empty(B, D, device=dev)
            _gemm_x3(Xp, W1p, Z, M=B, N=D, K=D)
            H1p = ops.split3(Z, planes=torch.empty(3, B, D, device=dev, dtype=bf), bias=b1.detach(), act=L.ACT_RELU)
            _gemm_x3(H1p, W2p, H2, M=B, N=H, K=D)
            ops.split3(H2, out=H2, bias=b2.detach(), act=L.ACT_TANH)
            aux = Z                          # storage reused for dH1 / dZ1 in backward
        else:
            mk = lambda t: ops.cast_bf16(t.detach().contiguous())
            Xp, W1p, W2p = mk(X), mk(W1), mk(W2)
            H1p = torch.empty(B, D, device=dev, dtype=bf)
            aux = torch.empty(B, D // 32, device=dev, dtype=torch.int32)      # ReLU sign bits for the backward mask
            ops.gemm_bf16(Xp, W1p, H1p, M=B, N=D, K=D, epi=L.EPI_BIAS_RELU_BF16, bias=b1.detach(), aux=aux)
            ops.gemm_bf16(H1p, W2p, H2, M=B, N=H, K=D, epi=L.EPI_BIAS_TANH_F32, bias=b2.detach())
        logits = ops.linear_fwd(H2, Wc.detach(), bc.detach(), L.ACT_NONE)
        ctx.save_for_backward(Xp, H1p, H2, W1p, W2p, Wc, aux)
        ctx.exact = exact
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        Xp, H1p, H2, W1p, W2p, Wc, aux = ctx.saved_tensors
        exact, need = ctx.exact, ctx.needs_input_grad
        dlogits = dlogits.contiguous()
        B, H = H2.shape
        D = Xp.shape[-1]
        dev, bf = H2.device, torch.bfloat16
        dWc, dbc = ops.linear_bwd_dw(dlogits, H2)
        dZ2 = ops.linear_bwd_dx(dlogits, Wc.detach(), mask_src=H2, mask_mode=L.ACT_TANH)
        db2 = ops.colsum(dZ2)
        dW2, dW1 = torch.zeros(H, D, device=dev), torch.zeros(D, D, device=dev)
        dX = torch.empty(B, D, device=dev) if need[0] else None
        if exact:
            dZ2p = ops.split3(dZ2, planes=torch.empty(3, B, H, device=dev, dtype=bf))
            dH1 = aux
            _gemm_x3(dZ2p, W2p, dH1, M=B, N=D, K=H, b_mn=True)
            dZ1p = ops.split3(dH1, out=dH1, planes=torch.empty(3, B, D, device=dev, dtype=bf), mask_plane=H1p[0])
            db1 = ops.colsum(dH1)
            _gemm_x3(dZ2p, H1p, dW2, M=H, N=D, K=B, a_mn=True, b_mn=True, zeroed=True)
            _gemm_x3(dZ1p, Xp, dW1, M=D, N=D, K=B, a_mn=True, b_mn=True, zeroed=True)
            if dX is not None:
                _gemm_x3(dZ1p, W1p, dX, M=B, N=D, K=D, b_mn=True)
        else:
            dZ2p = ops.cast_bf16(dZ2)
            dZ1p, db1 = torch.empty(B, D, device=dev, dtype=bf), torch.empty(D, device=dev)
            ops.gemm_bf16(dZ2p, W2p, dZ1p, M=B, N=D, K=H, b_mn=True, epi=L.EPI_BITMASK_BF16, aux=aux, colsum_out=db1)
            ops.gemm_bf16(dZ2p, H1p, dW2, M=H, N=D, K=B, a_mn=True, b_mn=True, epi=L.EPI_ATOMIC_F32, stream_k=True)
            ops.gemm_bf16(dZ1p, Xp, dW1, M=D, N=D, K=B, a_mn=True, b_mn=True, epi=L.EPI_ATOMIC_F32, stream_k=True)
            if dX is not None:
                ops.gemm_bf16(dZ1p, W1p, dX, M=B, N=D, K=D, b_mn=True, epi=L.EPI_STORE_F32)
        return dX, dW1, db1, dW2, db2, dWc, dbc, None


class ConcatModel(nn.Module):
    """The reference's `ConcatModel` head.  Parameters and their names match the reference so
    `load_state_dict(torch.load('model_dict/<run>/best_f1.pickle'), strict=False)` fills the head."""

    def __init__(self, feature_dims=(768, 768, 768), hidden=768, n_class=2, encoder: nn.Module | None = None,
                 private: bool = True, fixed_formula: bool = True, seed: int = REFERENCE_SEED, tau: float = 1.0,
                 precision: str = "fp32", tc_min_batch: int = 1024):
        super().__init__()
        if n_class != 2:
            raise NotImplementedError("the reference classifier is nn.Linear(768, 2)")
        self.feature_dims = tuple(int(d) for d in feature_dims)
        D = sum(self.feature_dims)
        self.encoder = encoder
        self.classifier = nn.Linear(hidden, n_class)                    # model.py:24
        self.DP = nn.parameter.Parameter(torch.zeros(1, D))             # model.py:25
        self.fc_layers = nn.Sequential(nn.Linear(D, D), nn.ReLU(),      # model.py:27-32
                                       nn.Linear(D, hidden), nn.Tanh())
        self.eps = torch.tensor(1.0)                                    # set by get_model (model.py:11)
        self.private = private            # False = the reference's model.py with the privacy block commented out
        self.fixed_formula = fixed_formula  # True: past_acc.py:132 ("# fix"); False: model.py:57
        self.tau = tau
        # arithmetic of the two dense layers: 'fp32' (reference arithmetic: CUDA-core kernels below tc_min_batch rows, the
        # fp32x3 tensor-core route from there on), 'fp32x3' (always the tensor-core route) or 'bf16' (bf16 GEMM operands)
        assert precision in ("fp32", "fp32x3", "bf16")
        self.precision, self.tc_min_batch = precision, int(tc_min_batch)
        self.seed = int(seed)
        self.noise_offset = 0             # one Philox offset per forward: fresh noise every pass
        self.return_gate_index = False
        self._injected = None
        self.last_gate_index = None

    # -- noise control -----------------------------------------------------------------------
    def inject_noise(self, lap: torch.Tensor, gumbel: torch.Tensor | None):
        """Use these Laplace [B,D] / Gumbel [2,B,D] tensors for the next forward (parity tests:
        the same tensors are fed to the reference path)."""
        self._injected = (lap.contiguous(), None if gumbel is None else gumbel.contiguous())

    def _blocks(self, x):
        if self.encoder is not None:
            x = self.encoder(x)
        if isinstance(x, torch.Tensor):
            x = (x,)
        blocks = [b if b.dim() == 2 else b.reshape(b.shape[0], -1) for b in x]
        got = tuple(b.shape[1] for b in blocks)
        if sum(got) != sum(self.feature_dims):
            raise ValueError(f"feature blocks {got} do not add up to the head width {sum(self.feature_dims)}")
        return [b.contiguous().float() for b in blocks]

    # -- reference API -----------------------------------------------------------------------
    def feature(self, x):
        """reference model.py:34-51: concat + per-row min-max normalise -> [B,D] in [0,1]."""
        return _NormaliseFn.apply(*self._blocks(x))

    def forward(self, x, hard=True, row0: int = 0):
        blocks = self._blocks(x)
        if not self.private:
            gated = _NormaliseFn.apply(*blocks)
        else:
            cfg = _NoiseCfg()
            cfg.tau, cfg.hard, cfg.fixed, cfg.row0 = float(self.tau), bool(hard), self.fixed_formula, int(row0)
            cfg.want_gate = self.return_gate_index
            if self._injected is not None:
                cfg.mode, (cfg.lap, cfg.gum) = L.NOISE_INJECTED, self._injected
                cfg.seed = cfg.offset = 0
                cfg.want_gate = cfg.gum is not None
                self._injected = None
            else:
                cfg.mode, cfg.lap, cfg.gum = L.NOISE_PHILOX, None, None
                cfg.seed, cfg.offset = self.seed, self.noise_offset
                self.noise_offset += 1
            res = _PerturbGateFn.apply(self.DP, exp_eps_of(self.eps), cfg, *blocks)
            if isinstance(res, tuple):
                gated, self.last_gate_index = res
            else:
                gated = res
        fc0, fc2 = self.fc_layers[0], self.fc_layers[2]
        args = (gated, fc0.weight, fc0.bias, fc2.weight, fc2.bias, self.classifier.weight, self.classifier.bias)
        D, H = fc0.weight.shape[1], fc2.weight.shape[0]
        big = gated.shape[0] >= self.tc_min_batch
        if self.precision == "bf16" and big and D % 128 == 0 and H % 8 == 0:
            return _HeadMLPTensorFn.apply(*args, "bf16")
        if D % 8 == 0 and H % 8 == 0 and (self.precision == "fp32x3" or (self.precision == "fp32" and big)):
            return _HeadMLPTensorFn.apply(*args, "fp32x3")
        return _HeadMLPFn.apply(*args)


def get_model(cfg):
    """reference model.py:8-12.  `cfg.data_name == 'EEG'` builds ConcatModel; `cfg.eps` is the
    privacy budget.  Optional extras: cfg.feature_dims, cfg.private, cfg.fixed_formula, cfg.precision, cfg.tc_min_batch."""
    if not torch.cuda.is_available():
        raise RuntimeError("get_model needs a CUDA device (the reference calls .cuda() too, model.py:11-12)")
    if cfg.data_name == 'EEG':
        model = ConcatModel(feature_dims=getattr(cfg, "feature_dims", (768, 768, 768)),
                            private=getattr(cfg, "private", True),
                            fixed_formula=getattr(cfg, "fixed_formula", True),
                            precision=getattr(cfg, "precision", "fp32"),
                            tc_min_batch=getattr(cfg, "tc_min_batch", 1024))
    else:
        raise ValueError(f"unknown data_name {cfg.data_name!r} (the reference only defines 'EEG')")
    model.eps = torch.tensor(cfg.eps).cuda()
    return model.cuda()


def cal_loss(prediction, label):
    """reference past_acc.py:71-77 == base_train.py:59-65, on [B,2] logits already on the device.
    (The fused classifier+CE kernel is used by the training engine; this helper exists so reference
    scripts that call cal_loss on the model output keep working.)"""
    label = label.squeeze(dim=1)
    loss = torch.nn.functional.cross_entropy(prediction, label)
    with torch.no_grad():
        pred_label_id = torch.argmax(prediction, dim=1)
        accuracy = (label == pred_label_id).float().sum() / label.shape[0]
    return loss, accuracy, pred_label_id, label
