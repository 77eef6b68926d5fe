"""Training / evaluation engine for one or many independent heads (the eps x seed sweep).

Implements the reference's step structure without autograd, on preallocated device buffers:

    pass 1  hard=False -> loss -> backward -> Adam(DP)            past_acc.py:198-203
    pass 2  hard=True  -> loss -> backward -> Adam(everything else) past_acc.py:206-212
    eval    hard=True, noise still on, no grad                     past_acc.py:218-228

Pass 1 only needs the dX chain and dDP, pass 2 only the weight gradients: the gradients the
reference computes and then discards (`zero_grad` at past_acc.py:206 / :198) are never formed.

Two arithmetic modes:
  precision='fp32'  all models advance together through the GROUPED fp32 CUDA-core kernels
                    (reference batch size, weight-streaming bound; parity mode, 1e-5)
  precision='bf16'  each model runs the tcgen05 GEMM path (large batch; bf16 operands, fp32
                    accumulate, fp32 master weights + Adam; 2e-2 bar)
  precision='fp32x3' the same tensor-core GEMMs held to the fp32 bar (1e-5) at large batch: every
                    fp32 operand enters as three bf16 planes and six plane-pair products are accumulated
                    (ops.split3 / ops.gemm_bf16x3); precision='fp32' switches to it by itself from
                    `tc_min_batch` rows on, where the batch is a real dense contraction

Parameter layout per model: one flat fp32 buffer [W1 | b1 | W2 | b2 | Wc | bc | pad] so the
optimiser is a single launch; `state_dict(i)` / `load_state_dict(i, sd)` use the reference's
key names (fc_layers.0.weight ... DP).
"""
from __future__ import annotations

import math

import torch

from . import _lib as L
from . import ops
from .model import REFERENCE_SEED, exp_eps_of


class HeadEngine:
    def __init__(self, n_models=1, feature_dims=(768, 768, 768), hidden=768, eps=1.0, seeds=None, lr=1e-6,
                 precision="fp32", fixed_formula=True, tau=1.0, device="cuda", init_seed=REFERENCE_SEED,
                 noise="philox", betas=(0.9, 0.999), adam_eps=1e-8, dp_init=None):
        if not torch.cuda.is_available():
            raise RuntimeError("HeadEngine needs a CUDA device: there is no CPU fallback")
        assert precision in ("fp32", "bf16", "fp32x3")
        if precision == "fp32x3" and (sum(int(d) for d in feature_dims) % 8 or int(hidden) % 8):
            raise ValueError("the fp32x3 tensor-core path needs widths that are multiples of 8")
        if precision == "bf16" and sum(int(d) for d in feature_dims) % 128:
            raise ValueError("the bf16 tensor-core path packs the ReLU sign bits in 128-bit rows: the fused width must be a "
                             "multiple of 128 (2304 and 2560 are); use precision='fp32' for other widths")
        self.M = int(n_models)
        self.dims = tuple(int(d) for d in feature_dims)
        self.D, self.H, self.C = sum(self.dims), int(hidden), 2
        self.device = torch.device(device)
        self.precision, self.fixed, self.tau, self.lr = precision, bool(fixed_formula), float(tau), float(lr)
        self.betas, self.adam_eps = betas, adam_eps
        self.noise = noise
        eps_list = list(eps) if isinstance(eps, (list, tuple)) else [eps] * self.M
        assert len(eps_list) == self.M
        self.eps = eps_list
        self.exp_eps = [exp_eps_of(e) for e in eps_list]
        self.seeds = [int(x) for x in seeds] if seeds is not None else [REFERENCE_SEED + i for i in range(self.M)]
        assert len(self.seeds) == self.M
        steps = {self.seeds[i + 1] - self.seeds[i] for i in range(self.M - 1)}
        # seeds in arithmetic progression -> the whole ensemble's noise is one grouped launch
        self.seed_step = (steps.pop() if steps else 0) if len(steps) <= 1 else None
        if self.seed_step is not None and self.seed_step < 0:
            self.seed_step = None
        D, H, C = self.D, self.H, self.C
        sizes = [("W1", (D, D)), ("b1", (D,)), ("W2", (H, D)), ("b2", (H,)), ("Wc", (C, H)), ("bc", (C,))]
        self.layout, off = {}, 0
        for name, shape in sizes:
            n = math.prod(shape)
            self.layout[name] = (off, shape)
            off += (n + 7) // 8 * 8  # 16-byte aligned segments in fp32 AND in the bf16 shadow (TMA)
        self.P = off
        self._views = {}
        dev = self.device
        self.flat = torch.zeros(self.M, self.P, device=dev)
        self.grad = torch.zeros(self.M, self.P, device=dev)
        self.m = torch.zeros(self.M, self.P, device=dev)
        self.v = torch.zeros(self.M, self.P, device=dev)
        self.DP = torch.zeros(self.M, D, device=dev)
        self.dDP = torch.zeros(self.M, D, device=dev)
        self.DP_m = torch.zeros(self.M, D, device=dev)
        self.DP_v = torch.zeros(self.M, D, device=dev)
        self.shadow = torch.zeros(self.M, self.P, device=dev, dtype=torch.bfloat16) if precision == "bf16" else None
        # fp32x3: hi/mid/lo bf16 planes of the two GEMM weights (allocated on first use: precision='fp32' only needs them
        # once a batch of tc_min_batch rows arrives)
        self.W1p = self.W2p = None
        self._planes_key = None
        self.tc_min_batch = 1024      # precision='fp32': batches from this size on take the fp32x3 tensor-core route
        # fp32x3: 64-wide k-blocks of the contraction per TMEM accumulation chain.  The tensor core truncates when it
        # accumulates (measured: error grows linearly with the chain, 2.5e-6 rms at 40 k-blocks, below cuBLAS SGEMM's
        # 9e-7 from ~12 down), so longer contractions are cut into K slabs summed by round-to-nearest fp32 adds.
        self.x3_chain_kb = 10
        self.t_model = 0
        self.t_dp = 0
        # bf16 path: True stores the Tanh output H2 [B,768] as bf16 as well.  Measured at B=65,536 (6 models): the
        # fc_layers.2 GEMM gains 3 % (0.241 -> 0.233 ms), cls_ce pass 2 -- issue-bound, not byte-bound -- loses 9 %
        # (0.465 -> 0.508 ms): +0.2 % on the step, so H2 stays fp32 (exact classifier / loss arithmetic) by default.
        self.h2_bf16 = False
        self.fuse_adam = True   # fp32 small-batch path: weight gradients recomputed inside the Adam kernel
        self.fast_replay = False  # launch-bound regimes: replay recorded C-ABI call plans (see _train_step_planned)
        self._plans = {}
        self.noise_offset = 0
        self._bufs = {}
        self._injected = None
        self._bucket_hook = None      # set for the duration of a data-parallel step whose grad_hook exchanges buckets
        self._coef_key = None
        self.exp_eps_dev = torch.tensor(self.exp_eps, dtype=torch.float32, device=dev)
        # arbitrary per-model seeds: a device array read by the grouped kernels (one launch for the whole sweep)
        self.seeds_dev = torch.tensor([s if s < 2 ** 63 else s - 2 ** 64 for s in self.seeds], dtype=torch.int64, device=dev)
        self._init_params(init_seed, dp_init)

    # ---- parameters ----------------------------------------------------------------------------
    def view(self, name, src=None):
        """[M, *shape] view of one parameter inside a flat per-model buffer (views are cached: the buffers never move)."""
        src = self.flat if src is None else src
        key = (name, src.data_ptr())
        v = self._views.get(key)
        if v is None:
            off, shape = self.layout[name]
            v = self._views[key] = src[:, off:off + math.prod(shape)].view(self.M, *shape)
        return v

    def _init_params(self, init_seed, dp_init):
        """nn.Linear default init (kaiming-uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))),
        one CPU generator per model so a model's init does not depend on the ensemble size."""
        for i in range(self.M):
            g = torch.Generator().manual_seed(int(init_seed) + i)
            for wn, bn in (("W1", "b1"), ("W2", "b2"), ("Wc", "bc")):
                _, shape = self.layout[wn]
                k = 1.0 / math.sqrt(shape[1])
                self.view(wn)[i].copy_((torch.rand(shape, generator=g) * 2 - 1) * k)
                self.view(bn)[i].copy_((torch.rand(shape[0], generator=g) * 2 - 1) * k)
            if dp_init is not None:   # one [D] vector for every model, or one per model ([M, D] / list of M vectors)
                init = dp_init[i] if (isinstance(dp_init, (list, tuple)) or (torch.is_tensor(dp_init) and dp_init.dim() == 2
                                                                             and dp_init.shape[0] == self.M and self.M > 1)) else dp_init
                self.DP[i].copy_(torch.as_tensor(init, dtype=torch.float32).reshape(-1))
        self.sync_shadow()

    def sync_shadow(self):
        if self.shadow is not None:
            ops.cast_bf16(self.flat, self.shadow)
        self._planes_key = None

    def _weight_planes(self):
        """hi/mid/lo planes of W1 / W2 for the fp32x3 path, refreshed when the weights moved (Adam step, load_state_dict)."""
        key = self.flat._version          # in-place torch writes; the Adam kernels reset _planes_key themselves
        if self.W1p is None:
            bf = torch.bfloat16
            self.W1p = torch.empty(self.M, 3, self.D, self.D, device=self.device, dtype=bf)
            self.W2p = torch.empty(self.M, 3, self.H, self.D, device=self.device, dtype=bf)
        if self._planes_key != key:
            W1, W2 = self.view("W1"), self.view("W2")
            for i in range(self.M):
                ops.split3(W1[i], planes=self.W1p[i])
                ops.split3(W2[i], planes=self.W2p[i])
            self._planes_key = key
        return self.W1p, self.W2p

    _KEYS = {"fc_layers.0.weight": "W1", "fc_layers.0.bias": "b1", "fc_layers.2.weight": "W2",
             "fc_layers.2.bias": "b2", "classifier.weight": "Wc", "classifier.bias": "bc"}

    def state_dict(self, i=0):
        sd = {k: self.view(n)[i].detach().clone() for k, n in self._KEYS.items()}
        sd["DP"] = self.DP[i].detach().clone().view(1, self.D)
        return sd

    def load_state_dict(self, i, sd, strict=False):
        missing = []
        for k, n in self._KEYS.items():
            if k in sd:
                self.view(n)[i].copy_(sd[k])
            else:
                missing.append(k)
        if "DP" in sd:
            self.DP[i].copy_(sd["DP"].reshape(-1))
        else:
            missing.append("DP")
        if strict and missing:
            raise KeyError(f"missing keys {missing}")
        self.sync_shadow()
        return missing

    # ---- helpers -------------------------------------------------------------------------------
    def _buf(self, name, shape, dtype):
        """Persistent scratch tensor per (name, shape, dtype): a differently-shaped batch (the tail of an epoch) gets
        its own buffers instead of reallocating, so pointers held by recorded call plans stay valid."""
        key = (name, tuple(shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            t = self._bufs[key] = torch.empty(shape, dtype=dtype, device=self.device)
        return t

    def _coef_state(self):
        """What the cached (w, eps_hat, d eps_hat/d DP) rows depend on: DP as torch sees it (`_version` counts in-place
        writes: load_state_dict, user code), DP as the Adam kernel moves it (t_dp), and e^eps / the formula switch."""
        return (self.DP._version, self.t_dp, self.exp_eps_dev.data_ptr(), self.exp_eps_dev._version, self.fixed)

    def inject_noise(self, lap, gum=None):
        """Per-model injected noise for the NEXT pass: lap [M,B,D], gum [M,2,B,D] (parity tests)."""
        self._injected = (lap.contiguous(), None if gum is None else gum.contiguous())

    def _ensure_coef(self):
        """The cached (w, eps_hat, d eps_hat/d DP) rows [3,M,D], recomputed only when DP moved."""
        coef = self._buf("coef", (3, self.M, self.D), torch.float32)
        key = self._coef_state()
        if self._coef_key != key:      # DP only moves in the DP pass (past_acc.py:203): one dp_coeffs per step, not two
            ops.dp_coeffs(self.DP, self.exp_eps_dev, self.fixed, out=coef)
            self._coef_key = key
        return coef

    def _perturb(self, blocks, hard, out, row0, n_rep=1):
        """kernel (a) for every model; returns the per-model coefficient rows and the noise spec.  n_rep > 1 (evaluation
        only): the batch is perturbed n_rep times with consecutive Philox offsets in one launch (train.py:126-131)."""
        M, D = self.M, self.D
        coef = self._ensure_coef()
        inj = self._injected
        self._injected = None
        offset = self.noise_offset
        self.noise_offset += n_rep
        if n_rep > 1:
            if inj is not None:
                raise ValueError("repeated evaluation draws its noise in-kernel (no injected tensors)")
            kw = dict(seed=self.seeds[0], seed_step=self.seed_step) if self.seed_step is not None else dict(model_seeds=self.seeds_dev)
            ops.perturb_gate_fwd(blocks, coef[0], coef[1], noise_mode=L.NOISE_PHILOX, offset=offset, row0=row0, tau=self.tau, hard=hard,
                                 out=out, n_models=M, n_rep=n_rep, **kw)
            return coef, (None, offset)
        if inj is not None:
            ops.perturb_gate_fwd(blocks, coef[0], coef[1], noise_mode=L.NOISE_INJECTED, lap=inj[0], gum=inj[1], tau=self.tau,
                                 hard=hard, want_gate=inj[1] is not None, out=out, n_models=M)
        elif self.seed_step is not None:   # one grouped launch for the whole ensemble
            ops.perturb_gate_fwd(blocks, coef[0], coef[1], noise_mode=L.NOISE_PHILOX, seed=self.seeds[0],
                                 seed_step=self.seed_step, offset=offset, row0=row0, tau=self.tau, hard=hard, out=out, n_models=M)
        elif blocks[0].shape[-2] * self.D < (1 << 22) or all(b.dim() == 2 for b in blocks):
            # one grouped launch, seeds from the device array: small batches (B=8 sweep), or a large batch SHARED by
            # the sweep (each row fetched and normalised once, perturbed once per model)
            ops.perturb_gate_fwd(blocks, coef[0], coef[1], noise_mode=L.NOISE_PHILOX, model_seeds=self.seeds_dev, offset=offset,
                                 row0=row0, tau=self.tau, hard=hard, out=out, n_models=M)
        else:                                            # large per-model batches: one TMA-ring launch per model
            for i in range(M):
                bl = [b[i] if b.dim() == 3 else b for b in blocks]
                ops.perturb_gate_fwd(bl, coef[0, i], coef[1, i], noise_mode=L.NOISE_PHILOX, seed=self.seeds[i], offset=offset,
                                     row0=row0, tau=self.tau, hard=hard, out=out[i])
        return coef, (inj, offset)

    def _dDP_one(self, i, dXi, coef, noise_spec, row0):
        inj, offset = noise_spec
        if inj is not None:
            ops.perturb_gate_bwd_dp(dXi, coef[2, i], noise_mode=L.NOISE_INJECTED, lap=inj[0][i], out=self.dDP[i])
        else:
            ops.perturb_gate_bwd_dp(dXi, coef[2, i], noise_mode=L.NOISE_PHILOX, seed=self.seeds[i], offset=offset,
                                    row0=row0, out=self.dDP[i])

    def _dDP(self, dX, coef, noise_spec, row0):
        inj, offset = noise_spec
        if inj is not None:
            ops.perturb_gate_bwd_dp(dX, coef[2], noise_mode=L.NOISE_INJECTED, lap=inj[0], out=self.dDP)
        elif self.seed_step is not None:
            ops.perturb_gate_bwd_dp(dX, coef[2], noise_mode=L.NOISE_PHILOX, seed=self.seeds[0], seed_step=self.seed_step,
                                    offset=offset, row0=row0, out=self.dDP)
        else:
            ops.perturb_gate_bwd_dp(dX, coef[2], noise_mode=L.NOISE_PHILOX, model_seeds=self.seeds_dev, offset=offset, row0=row0,
                                    out=self.dDP)

    def _ce_out(self, mode, B):
        """Persistent output buffers of the classifier/CE kernel (one set per pass kind, so the statistics of pass 2
        survive the next step's pass 1 and recorded call plans see stable pointers).  Evaluation keeps fresh tensors."""
        if mode == "eval":
            return {}
        M = self.M
        return dict(logits=self._buf("logits_" + mode, (M, B, 2), torch.float32), pred=self._buf("pred_" + mode, (M, B), torch.int64),
                    stats=self._buf("stats_" + mode, (M, 4), torch.float32))

    def _labels(self, labels, B=None):
        """int64 labels as the kernels take them: [B] shared by every model, or [M,B] per model.  Accepted inputs: [B], the
        reference's [B,1] column (past_acc.py:71 squeezes it), or [M,B]; with `B` (the batch size of the feature blocks) given
        the shape is checked against it, so that a [M,1] per-model tensor of a 1-row tail batch is not mistaken for a column."""
        if labels.dim() == 2:
            if B is not None and labels.shape == (self.M, B) and not (labels.shape[1] == 1 and labels.shape[0] == B):
                pass                                        # per-model labels
            elif labels.shape[1] == 1 and (B is None or labels.shape[0] == B):
                labels = labels[:, 0]                       # the reference's [B,1] column
            elif B is not None:
                raise ValueError(f"labels of shape {tuple(labels.shape)} match neither [B]=[{B}], [B,1] nor [M,B]=[{self.M},{B}]")
        elif labels.dim() != 1 or (B is not None and labels.shape[0] != B):
            raise ValueError(f"labels of shape {tuple(labels.shape)} do not match the batch of {B} rows")
        return labels.contiguous()

    # ---- one forward(+backward) pass -----------------------------------------------------------
    def _pass(self, blocks, labels, hard, mode, row0=0, global_batch=None, fuse_adam=False, n_rep=1):
        """mode: 'dp' (pass 1), 'model' (pass 2) or 'eval'.  Returns the cls_ce result dict.
        fuse_adam (fp32 path, batch <= 8, no gradient all-reduce): the Adam step of pass 2 is applied inside the
        pass, with the weight gradients recomputed in the optimiser kernel instead of written to HBM;
        res['adam_done'] tells the caller."""
        B = blocks[0].shape[-2] * n_rep        # n_rep > 1 (eval): the repetitions are further batch rows, repetition-major
        M, D, H = self.M, self.D, self.H
        gb = float(global_batch or B)
        if n_rep > 1:
            assert mode == "eval"
            labels = labels.repeat(n_rep) if labels.dim() == 1 else labels.repeat(1, n_rep)
        W1, b1, W2, b2, Wc, bc = (self.view(n) for n in ("W1", "b1", "W2", "b2", "Wc", "bc"))
        backward = mode != "eval"
        if self.precision == "fp32x3" or (self.precision == "fp32" and B >= self.tc_min_batch and D % 8 == 0 and H % 8 == 0):
            return self._pass_x3(blocks, labels, hard, mode, row0, gb, n_rep, B)
        if self.precision == "fp32":
            X = self._buf("X", (M, B, D), torch.float32)
            coef, nspec = self._perturb(blocks, hard, X, row0, n_rep)
            H1 = ops.linear_fwd(X, W1, b1, L.ACT_RELU, out=self._buf("H1", (M, B, D), torch.float32))
            H2 = ops.linear_fwd(H1, W2, b2, L.ACT_TANH, out=self._buf("H2", (M, B, H), torch.float32))
            res = ops.cls_ce(H2, Wc, bc, labels, loss_scale=1.0 / B, grad_scale=1.0 / gb, backward=backward, **self._ce_out(mode, B),
                             dz=self._buf("dZ2", (M, B, H), torch.float32) if backward else None,
                             dWc=self.view("Wc", self.grad) if mode == "model" else None,
                             dbc=self.view("bc", self.grad) if mode == "model" else None, want_dw=mode == "model")
            if mode == "eval":
                return res
            dZ2 = res["dz"]
            dZ1 = ops.linear_bwd_dx(dZ2, W2, mask_src=H1, mask_mode=L.ACT_RELU, out=self._buf("dZ1", (M, B, D), torch.float32))
            if mode == "dp":
                dX = ops.linear_bwd_dx(dZ1, W1, out=self._buf("dX", (M, B, D), torch.float32))
                self._dDP(dX, coef, nspec, row0)
            elif fuse_adam and B <= 8:
                kw = dict(step=self.t_model, lr=self.lr, betas=self.betas, eps=self.adam_eps)
                for wn, bn, dY, inp in (("W2", "b2", dZ2, H1), ("W1", "b1", dZ1, X)):
                    ops.linear_adam_step(dY, inp, self.view(wn), self.view(wn, self.m), self.view(wn, self.v),
                                         self.view(bn), self.view(bn, self.m), self.view(bn, self.v), **kw)
                off = self.layout["Wc"][0]   # [Wc | bc] of every model: gradients came from cls_ce
                ops.adam_step_strided(self.flat[:, off:], self.grad[:, off:], self.m[:, off:], self.v[:, off:], **kw)
                res["adam_done"] = True
            else:
                ops.linear_bwd_dw(dZ2, H1, dW=self.view("W2", self.grad), db=self.view("b2", self.grad))
                ops.linear_bwd_dw(dZ1, X, dW=self.view("W1", self.grad), db=self.view("b1", self.grad))
            return res
        # ---- bf16 tensor-core path, one model at a time
        bf = torch.bfloat16
        if mode == "model":   # the split-K weight-gradient GEMMs accumulate into a zeroed buffer
            ops.fill_zero(self.grad)
        X = self._buf("Xh", (M, B, D), bf)
        coef, nspec = self._perturb(blocks, hard, X, row0, n_rep)
        H1 = self._buf("H1h", (M, B, D), bf)
        h2b = bool(self.h2_bf16)
        H2 = self._buf("H2h", (M, B, H), bf) if h2b else self._buf("H2f", (M, B, H), torch.float32)
        W1h, W2h = self.view("W1", self.shadow), self.view("W2", self.shadow)
        # ReLU sign bits (1 bit per activation) for the backward mask
        bits = self._buf("relu_bits", (M, B, D // 32), torch.int32) if backward else None
        for i in range(M):
            ops.gemm_bf16(X[i], W1h[i], H1[i], M=B, N=D, K=D, epi=L.EPI_BIAS_RELU_BF16, bias=b1[i],
                          aux=None if bits is None else bits[i])
            ops.gemm_bf16(H1[i], W2h[i], H2[i], M=B, N=H, K=D, epi=L.EPI_BIAS_TANH_BF16 if h2b else L.EPI_BIAS_TANH_F32, bias=b2[i])
        res = ops.cls_ce(H2, Wc, bc, labels, loss_scale=1.0 / B, grad_scale=1.0 / gb, backward=backward, **self._ce_out(mode, B),
                         dz=self._buf("dZ2h", (M, B, H), bf) if backward else None, dz_dtype=bf,
                         dWc=self.view("Wc", self.grad) if mode == "model" else None,
                         dbc=self.view("bc", self.grad) if mode == "model" else None, want_dw=mode == "model",
                         dz_colsum=self.view("b2", self.grad) if mode == "model" else None)   # db2 rides along
        if mode == "eval":
            return res
        dZ2 = res["dz"]
        dZ1 = self._buf("dZ1h", (M, B, D), bf)
        gb1 = self.view("b1", self.grad)
        for i in range(M):
            # dZ1 = (dZ2 . W2) * relu'(H1):   B operand = W2 stored [K=H, N=D]  -> MN-major.  In pass 2 the bias
            # gradient db1 = colsum(dZ1) is reduced inside the epilogue.
            ops.gemm_bf16(dZ2[i], W2h[i], dZ1[i], M=B, N=D, K=H, b_mn=True, epi=L.EPI_BITMASK_BF16, aux=bits[i],
                          colsum_out=gb1[i] if mode == "model" else None)
        if mode == "dp":
            inj, offset = nspec
            for i in range(M):
                if inj is None:
                    # dX = dZ1 . W1 only feeds dDP = deps * colsum(dX * noise): fused, dX never leaves TMEM
                    ops.gemm_bf16_ddp(dZ1[i], W1h[i], M=B, N=D, K=D, b_mn=True, seed=self.seeds[i], offset=offset, row0=row0,
                                      deps_dDP=coef[2, i], out=self.dDP[i])
                else:   # injected-noise parity mode: materialise dX (fp32) and reduce against the supplied tensor
                    dX = self._buf("dXf", (B, D), torch.float32)
                    ops.gemm_bf16(dZ1[i], W1h[i], dX, M=B, N=D, K=D, b_mn=True, epi=L.EPI_STORE_F32)
                    self._dDP_one(i, dX, coef, nspec, row0)
            return res
        gW1, gW2 = self.view("W1", self.grad), self.view("W2", self.grad)
        bucket = self._bucket_hook
        off_tail = self.layout["b1"][0]
        for i in range(M):
            # dW = dZ^T . act: both operands are stored [K=B, *] -> MN-major, K = batch, split-K
            ops.gemm_bf16(dZ2[i], H1[i], gW2[i], M=H, N=D, K=B, a_mn=True, b_mn=True, epi=L.EPI_ATOMIC_F32, stream_k=True)
            if bucket is None:
                ops.gemm_bf16(dZ1[i], X[i], gW1[i], M=D, N=D, K=B, a_mn=True, b_mn=True, epi=L.EPI_ATOMIC_F32, stream_k=True)
                continue
            # data-parallel mode: everything but fc_layers.0.weight is final (b1 from the dZ1 epilogue, b2 / Wc / bc from
            # cls_ce, W2 just queued) and travels while the fc_layers.0 weight gradient is computed in row blocks, each
            # block's exchange running under the next block's GEMM (parallel.OverlappedAllReduce)
            self._bucket(bucket, i * self.P + off_tail, self.P - off_tail)
            nchunk = max(1, min(int(bucket.w1_chunks), D // 256))
            rows = -(-D // nchunk // 8) * 8
            for r0 in range(0, D, rows):
                r1 = min(D, r0 + rows)
                ops.gemm_bf16(dZ1[i][:, r0:r1], X[i], gW1[i][r0:r1], M=r1 - r0, N=D, K=B, a_mn=True, b_mn=True,
                              epi=L.EPI_ATOMIC_F32, stream_k=True)
                self._bucket(bucket, i * self.P + self.layout["W1"][0] + r0 * D, (r1 - r0) * D)
        return res

    def _pass_x3(self, blocks, labels, hard, mode, row0, gb, n_rep, B):
        """One pass on the fp32-parity tensor-core route (models.py:46-51,80 in fp32 arithmetic at large batch): each GEMM
        contracts hi/mid/lo bf16 planes of BOTH operands (six plane pairs, fp32 accumulate), the elementwise stages between
        them (bias, ReLU, tanhf, ReLU mask) run in fp32 and emit the next GEMM's planes.  One model at a time: the plane
        buffers are shared by the models of the engine."""
        M, D, H = self.M, self.D, self.H
        bf, f32 = torch.bfloat16, torch.float32
        b1, b2, Wc, bc = (self.view(n) for n in ("b1", "b2", "Wc", "bc"))
        W1p, W2p = self._weight_planes()
        backward = mode != "eval"
        if mode == "model":
            ops.fill_zero(self.grad)      # the split-K weight-gradient GEMMs accumulate
        X = self._buf("X", (M, B, D), f32)
        coef, nspec = self._perturb(blocks, hard, X, row0, n_rep)
        X3, H13 = self._buf("X3", (3, B, D), bf), self._buf("H13", (3, B, D), bf)
        Z1 = self._buf("Z1f", (B, D), f32)
        H2 = self._buf("H2", (M, B, H), f32)
        out = self._ce_out(mode, B)
        if not out:
            out = dict(logits=torch.empty(M, B, 2, device=self.device), pred=torch.empty(M, B, dtype=torch.int64, device=self.device),
                       stats=torch.empty(M, 4, device=self.device))
        chain = max(1, int(self.x3_chain_kb))

        def slabs_of(K):
            return max(1, round(((K + 63) // 64) / chain))

        def gemm(A3, B3, C, bias=None, **kw):   # forward / input-gradient GEMMs; returns the bias still to be added
            ns = slabs_of(kw["K"])
            if ns > 1:
                ops.fill_zero(C)
                ops.gemm_bf16x3(A3, B3, C, epi=L.EPI_ATOMIC_F32, k_slabs=ns, **kw)
                return bias
            ops.gemm_bf16x3(A3, B3, C, epi=L.EPI_STORE_F32 if bias is None else L.EPI_BIAS_F32, bias=bias, **kw)
            return None

        for i in range(M):
            ops.split3(X[i], planes=X3)
            rb = gemm(X3, W1p[i], Z1, bias=b1[i], M=B, N=D, K=D)
            ops.split3(Z1, planes=H13, bias=rb, act=L.ACT_RELU)
            rb = gemm(H13, W2p[i], H2[i], bias=b2[i], M=B, N=H, K=D)
            ops.split3(H2[i], out=H2[i], bias=rb, act=L.ACT_TANH)
            lab = labels if labels.dim() == 1 else labels[i]
            want_dw = mode == "model"
            res = ops.cls_ce(H2[i], Wc[i], bc[i], lab, loss_scale=1.0 / B, grad_scale=1.0 / gb, backward=backward,
                             logits=out["logits"][i], pred=out["pred"][i], stats=out["stats"][i],
                             dz=self._buf("dZ2f", (B, H), f32) if backward else None,
                             dWc=self.view("Wc", self.grad)[i] if want_dw else None,
                             dbc=self.view("bc", self.grad)[i] if want_dw else None, want_dw=want_dw,
                             dz_colsum=self.view("b2", self.grad)[i] if want_dw else None)
            if not backward:
                continue
            dZ23, dZ13 = self._buf("dZ23", (3, B, H), bf), self._buf("dZ13", (3, B, D), bf)
            ops.split3(res["dz"], planes=dZ23)
            # dH1 = dZ2 . W2 (B operand stored [K=H, N=D]: MN-major), then the ReLU mask from the sign of H1's hi plane
            gemm(dZ23, W2p[i], Z1, M=B, N=D, K=H, b_mn=True)
            ops.split3(Z1, out=Z1, planes=dZ13, mask_plane=H13[0])        # Z1 now holds dZ1 (fp32), dZ13 its planes
            if mode == "dp":
                dX = self._buf("dXf", (B, D), f32)
                gemm(dZ13, W1p[i], dX, M=B, N=D, K=D, b_mn=True)
                self._dDP_one(i, dX, coef, nspec, row0)
                continue
            ops.colsum(Z1, out=self.view("b1", self.grad)[i])
            slabs = slabs_of(B)
            # dW = dZ^T . act: both operands stored [K=B, *] (MN-major), K = batch cut into slabs of x3_chain_kb k-blocks
            ops.gemm_bf16x3(dZ23, H13, self.view("W2", self.grad)[i], M=H, N=D, K=B, a_mn=True, b_mn=True,
                            epi=L.EPI_ATOMIC_F32, k_slabs=slabs)
            ops.gemm_bf16x3(dZ13, X3, self.view("W1", self.grad)[i], M=D, N=D, K=B, a_mn=True, b_mn=True,
                            epi=L.EPI_ATOMIC_F32, k_slabs=slabs)
        return dict(logits=out["logits"], pred=out["pred"], stats=out["stats"])

    def _bucket(self, hook, off, n):
        """A finished slice [off, off+n) of the flat gradient buffer goes to the data-parallel exchange."""
        if ops.RECORD is not None:
            ops.RECORD.append((("hook",), None, ("bucket", off, n)))
        hook.bucket(self.grad.view(-1)[off:off + n])

    # ---- public steps ---------------------------------------------------------------------------
    def train_step(self, blocks, labels, row0=0, global_batch=None, grad_hook=None, dp_pass=True):
        """One reference step (past_acc.py:198-212) for every model.  `blocks`: list of [B,Di]
        (shared by all models) or [M,B,Di]; labels int64 [B] / [B,1] (or [M,B]).
        `grad_hook(tensor)` is called on each gradient buffer before its Adam step (data-parallel
        all-reduce).  Returns per-model stats of pass 2: dict(loss[M], acc[M])."""
        with ops.stream_scope():
            if self.fast_replay and self._injected is None:
                return self._train_step_planned(blocks, labels, row0, global_batch, grad_hook, dp_pass)
            return self._train_step(blocks, labels, row0, global_batch, grad_hook, dp_pass)

    def _state(self):
        return (self.noise_offset, self.t_dp, self.t_model)

    def _train_step_planned(self, blocks, labels, row0, global_batch, grad_hook, dp_pass):
        """Launch-bound regimes: record the C-ABI calls of two steps with this input signature, then replay them
        (ops.CallPlan) -- same functions, buffers and order, none of the Python between the launches."""
        labels = self._labels(labels, blocks[0].shape[-2])
        blocks = [b.contiguous() for b in blocks]
        key = (tuple((tuple(b.shape), tuple(b.stride()), b.dtype) for b in blocks), tuple(labels.shape), tuple(labels.stride()),
               global_batch, bool(dp_pass), grad_hook is not None, ops._stream())   # a plan replays on the stream it was recorded on
        ent = self._plans.get(key)
        inputs = {("block", i): b.data_ptr() for i, b in enumerate(blocks)}
        inputs["labels"] = labels.data_ptr()
        extents = [(t.data_ptr(), t.data_ptr() + t.numel() * t.element_size()) for t in (*blocks, labels)]
        if ent is not None and "plan" in ent:
            n, rem = divmod(self.t_model - ent["state"][2], 1)
            d = ent["dstate"]
            ok = (self.noise_offset == ent["state"][0] + n * d[0] and self.t_dp == ent["state"][1] + n * d[1]
                  and row0 == ent["row0"] + n * ent["drow0"] and ent["plan"].ws_generation == ops.WS_GENERATION
                  and self._coef_key == self._coef_state())   # the plan skips pass 1's dp_coeffs: the cached rows must be current
            if ok:
                hook = None
                if grad_hook is not None:
                    def hook(which, off=0, n=0):
                        if which == "bucket":
                            grad_hook.bucket(self.grad.view(-1)[off:off + n])
                        elif which == "finish":
                            grad_hook.finish()
                        else:
                            grad_hook(self.dDP if which == "dDP" else self.grad)
                ent["plan"].replay(n, inputs, hook)
                self.noise_offset += d[0]
                self.t_dp += d[1]
                self.t_model += d[2]
                self._coef_key = self._coef_state()      # the replayed pass 2 recomputed the rows after the DP update
                self._planes_key = None                  # the replayed Adam moved the weights after the step's own split
                return self._result(ent["stats"])
            self._plans.pop(key)          # the step state no longer advances the way it did while recording
            ent = None
        # record this step through the ordinary wrapper path
        state, rec = self._state(), []
        ops.RECORD = rec
        try:
            result = self._train_step(blocks, labels, row0, global_batch, grad_hook, dp_pass)
        finally:
            ops.RECORD = None
        ptrs = {v: k for k, v in inputs.items()}
        if ent is None:
            self._plans[key] = dict(rec=rec, state=state, row0=row0, ptrs=ptrs, extents=extents)
        else:
            apart = state[2] - ent["state"][2]
            try:
                if apart < 1:
                    raise RuntimeError("recorded steps out of order")
                plan = ops.CallPlan(ent["rec"], rec, ent["ptrs"], ptrs, steps_apart=apart, input_extents=ent["extents"] + extents)
                dstate = tuple((b - a) // apart for a, b in zip(ent["state"], state))
                self._plans[key] = dict(plan=plan, state=ent["state"], dstate=dstate, row0=ent["row0"],
                                        drow0=(row0 - ent["row0"]) // apart, stats=self._buf("stats_model", (self.M, 4), torch.float32))
            except RuntimeError:
                self._plans[key] = dict(rec=rec, state=state, row0=row0, ptrs=ptrs, extents=extents)   # start over from this step
        return result

    def _train_step(self, blocks, labels, row0, global_batch, grad_hook, dp_pass):
        labels = self._labels(labels, blocks[0].shape[-2])
        blocks = [b.contiguous() for b in blocks]
        if dp_pass:   # dp_pass=False reproduces train.py, where the DP pass is commented out (train.py:100-105)
            self._pass(blocks, labels, hard=False, mode="dp", row0=row0, global_batch=global_batch)
            if grad_hook is not None:
                if ops.RECORD is not None:
                    ops.RECORD.append((("hook",), None, ("dDP",)))
                grad_hook(self.dDP)
            self.t_dp += 1
            ops.adam_step(self.DP, self.dDP, self.DP_m, self.DP_v, self.t_dp, self.lr, self.betas, self.adam_eps)
        self.t_model += 1
        bucketed = grad_hook is not None and getattr(grad_hook, "bucketed", False) and self.precision == "bf16"
        self._bucket_hook = grad_hook if bucketed else None
        try:
            res = self._pass(blocks, labels, hard=True, mode="model", row0=row0, global_batch=global_batch,
                             fuse_adam=grad_hook is None and self.fuse_adam)
        finally:
            self._bucket_hook = None
        if not res.get("adam_done"):
            if bucketed:          # every slice is already travelling: wait for the last one
                if ops.RECORD is not None:
                    ops.RECORD.append((("hook",), None, ("finish",)))
                grad_hook.finish()
            elif grad_hook is not None:
                if ops.RECORD is not None:
                    ops.RECORD.append((("hook",), None, ("grad",)))
                grad_hook(self.grad)
            ops.adam_step(self.flat, self.grad, self.m, self.v, self.t_model, self.lr, self.betas, self.adam_eps,
                          bf16_shadow=self.shadow)
        self._planes_key = None           # the weights moved: the fp32x3 planes are refreshed on their next use
        return self._result(res["stats"])

    def _result(self, stats):
        """Per-model statistics of pass 2 as fresh tensors (the kernels write into persistent buffers that the next
        step overwrites)."""
        st = stats.view(self.M, 4).clone()
        return dict(loss=st[:, 0], acc=st[:, 2], n_correct=st[:, 1], stats=st)   # stats [M,4] = {loss, n_correct, acc, B}

    @torch.no_grad()
    def eval_step(self, blocks, labels, row0=0, n_eval=1):
        """past_acc.py:218-228: hard=True, noise still sampled.  Returns dict(loss, acc, pred, logits).
        n_eval > 1: the n_eval repeated stochastic evaluations of train.py:126-131 as ONE batched pass (n_eval Philox
        offsets inside one launch of every kernel); pred / logits are then [M, n_eval, B(, 2)], loss / acc the means over
        all repetitions."""
        labels = self._labels(labels, blocks[0].shape[-2])
        B = blocks[0].shape[-2]
        with ops.stream_scope():
            res = self._pass([b.contiguous() for b in blocks], labels, hard=True, mode="eval", row0=row0, n_rep=int(n_eval))
        st = res["stats"].view(self.M, 4)
        pred, logits = res["pred"], res["logits"]
        if n_eval > 1:
            pred, logits = pred.view(self.M, n_eval, B), logits.view(self.M, n_eval, B, 2)
        return dict(loss=st[:, 0], acc=st[:, 2], n_correct=st[:, 1], pred=pred, logits=logits)
