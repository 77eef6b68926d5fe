"""Tensor-level wrappers over the C ABI: validate torch tensors, pass raw device pointers and
the current CUDA stream.  No arithmetic happens here and nothing falls back to torch ops."""
from __future__ import annotations

import torch

from . import _lib as L


# torch.cuda.current_stream() costs more host time than the ctypes call it feeds; a caller that issues many
# kernels back to back (HeadEngine) pins the handle for the duration of a step with `with ops.stream_scope():`
_PINNED_STREAM = None


class stream_scope:
    def __enter__(self):
        global _PINNED_STREAM
        self.prev = _PINNED_STREAM
        _PINNED_STREAM = torch.cuda.current_stream().cuda_stream
        return self

    def __exit__(self, *exc):
        global _PINNED_STREAM
        _PINNED_STREAM = self.prev
        return False


def _stream() -> int:
    return _PINNED_STREAM if _PINNED_STREAM is not None else torch.cuda.current_stream().cuda_stream


_QUERY_CACHE = {}


def _query(name, *args) -> int:
    """Workspace-size queries are pure functions of their arguments: cache them."""
    key = (name, args)
    v = _QUERY_CACHE.get(key)
    if v is None:
        v = _QUERY_CACHE[key] = L.query(name, *args)
    return v


def _ptr(t):
    return None if t is None else t.data_ptr()


def _dt(t) -> int:
    if t.dtype == torch.float32:
        return L.DT_F32
    if t.dtype == torch.bfloat16:
        return L.DT_BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _chk(t, dtype=None, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the pgfuse kernels have no CPU fallback")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if t.dim() >= 1 and t.stride(-1) != 1 and t.shape[-1] != 1:
        raise ValueError(f"{name} must be contiguous in its last dimension")
    return t


WS_GENERATION = 0   # bumped whenever a scratch buffer is (re)allocated: recorded call plans hold raw pointers into them


class Workspace:
    """Grow-only scratch buffer (caller-owned scratch of the C ABI)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes: int, device) -> torch.Tensor:
        global WS_GENERATION
        n = max(4, (nbytes + 3) // 4)
        if self.buf is None or self.buf.numel() < n or self.buf.device != device:
            self.buf = torch.empty(n, dtype=torch.float32, device=device)
            WS_GENERATION += 1
        return self.buf


_ws = {}

# optional per-launch timing (bench.py roofline): list of (tag, start_event, end_event); tag[0] names the op
TIMING = None


# when a list, every C-ABI call is appended as (tag, name, args): HeadEngine records two consecutive steps and turns
# them into a CallPlan (below)
RECORD = None


def _call(tag, name, *args):
    """L.call, bracketed by CUDA events on the launching stream when bench.py asks for it."""
    if RECORD is not None:
        RECORD.append((tag, name, args))
    if TIMING is None:
        return L.call(name, *args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.call(name, *args)
    e1.record()
    TIMING.append((tag, e0, e1))


class CallPlan:
    """The C-ABI calls of one engine step, replayable without the Python wrappers around them.

    In the launch-bound regimes (few models per GPU at the reference batch size; small per-GPU shards of the
    data-parallel mode) a step is ~20 kernels of a few microseconds each and the tensor-level wrappers (validation,
    stride arithmetic, workspace look-ups) cost more host time than the GPU needs.  Two consecutive recorded steps
    give, per call, the argument tuple and -- by difference -- the arguments that advance with the step (Philox
    offsets, Adam step counts) and their increments; caller-supplied input pointers are substituted by position.
    Replaying is then one ctypes call per kernel with the same C functions, same buffers, same order: results are
    bit-identical to the wrapper path (tests/test_gpu_model.py)."""

    def __init__(self, rec_a, rec_b, input_ptrs, input_ptrs_b=None, steps_apart=1, input_extents=()):
        """input_ptrs / input_ptrs_b: {data pointer: key} of the caller's input tensors in the two recorded steps;
        input_extents: [lo, hi) address ranges of those tensors' storage -- a pointer argument that lies INSIDE one of them
        without being its base (a slice such as blocks[i] of a [M,B,D] input) cannot be substituted on replay and would keep
        reading the recorded step's memory, so such a sequence is refused (the caller then stays on the wrapper path)."""
        input_ptrs_b = input_ptrs if input_ptrs_b is None else input_ptrs_b

        def derived(ptr):
            return ptr is not None and any(lo < ptr < hi for lo, hi in input_extents)
        if len(rec_a) != len(rec_b) or any(a[1] != b[1] or len(a[2]) != len(b[2]) for a, b in zip(rec_a, rec_b)):
            raise RuntimeError("the two recorded steps issued different call sequences")
        lib = L.load()
        self.calls = []
        for (tag, name, a), (_, _, b) in zip(rec_a, rec_b):
            if name is None:                      # marker: gradient hook (all-reduce) of the data-parallel mode
                self.calls.append((tag, None, None, list(a), (), (), 0))
                continue
            types = L.SIGNATURES[name][1]
            dyn, sub = [], []
            for i, (x, y) in enumerate(zip(a, b)):
                if types[i] is L.P:
                    if x in input_ptrs:
                        if input_ptrs_b.get(y) != input_ptrs[x]:
                            raise RuntimeError(f"{name}: argument {i} is an input pointer in one recorded step but not in the other")
                        sub.append((i, input_ptrs[x]))
                    elif x != y:
                        raise RuntimeError(f"{name}: pointer argument {i} changed between the recorded steps (buffers must be stable)")
                    elif derived(x):
                        raise RuntimeError(f"{name}: pointer argument {i} points into a caller input tensor (a slice of it): "
                                           "a replay with other inputs would read stale memory")
                elif x != y:
                    if not (isinstance(x, int) and isinstance(y, int)) or (y - x) % steps_apart:
                        raise RuntimeError(f"{name}: argument {i} changed between steps in a way a plan cannot replay ({x} -> {y})")
                    dyn.append((i, (y - x) // steps_apart, 0xFFFFFFFF if types[i] is L.U32 else 0xFFFFFFFFFFFFFFFF))
            self.calls.append((tag, name, getattr(lib, name), list(a), dyn, sub, L.launches_of(name, a)))
        self.n_launches = sum(c[6] for c in self.calls)
        self.ws_generation = WS_GENERATION   # the plan is void once any scratch buffer it may point into has moved

    def replay(self, n, inputs, hook=None):
        """Issue the step that lies n steps after the first recorded one; `inputs` maps the substitution keys to the
        data pointers of this step's input tensors; `hook(which)` is called at the recorded gradient-hook points."""
        lib = L.load()
        for tag, name, fn, base, dyn, sub, _ in self.calls:
            if name is None:
                hook(*base)
                continue
            args = base
            if dyn or sub:
                args = list(base)
                for i, d, mask in dyn:
                    args[i] = (base[i] + n * d) & mask
                for i, key in sub:
                    args[i] = inputs[key]
            if TIMING is None:
                rc = fn(*args)
            else:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = fn(*args)
                e1.record()
                TIMING.append((tag, e0, e1))
            if rc != 0:
                raise RuntimeError(f"{name} failed ({rc}): {lib.pgf_last_error().decode()}")
        L.launch_count += self.n_launches


def workspace(tag: str) -> Workspace:
    """Scratch buffer of one op ON THE LAUNCHING STREAM: work issued on different streams (independent engines whose
    kernels overlap) must not share scratch, work on one stream is ordered and may."""
    return _ws.setdefault((tag, _stream()), Workspace())


# --------------------------------------------------------------------------------------------
def _exp_eps_tensor(exp_eps, n_models, device):
    if isinstance(exp_eps, torch.Tensor):
        t = exp_eps.to(device=device, dtype=torch.float32).reshape(-1)
    else:
        vals = list(exp_eps) if isinstance(exp_eps, (list, tuple)) else [exp_eps] * n_models
        t = torch.tensor([float(v) for v in vals], dtype=torch.float32, device=device)
    assert t.numel() == n_models
    return t.contiguous()


def _seeds_ptr(model_seeds, n_models):
    """Device int64 tensor [n_models] of per-model Philox seeds (bit pattern of the uint64 the kernels read)."""
    if model_seeds is None:
        return None
    _chk(model_seeds, torch.int64, "model_seeds")
    assert model_seeds.is_contiguous() and model_seeds.numel() == n_models
    return model_seeds.data_ptr()


def dp_coeffs(DP: torch.Tensor, exp_eps, fixed: bool = True, out=None):
    """(w, eps_hat, deps_dDP), each shaped like DP ([D] or [n_models, D]).  models.py:73,75.
    `exp_eps`: float (one model), list, or device tensor [n_models] of e^eps values."""
    _chk(DP, torch.float32, "DP")
    assert DP.is_contiguous()
    D = DP.shape[-1]
    n_models = DP.numel() // D
    ee = _exp_eps_tensor(exp_eps, n_models, DP.device)
    if out is None:
        out = torch.empty(3, *DP.shape, dtype=torch.float32, device=DP.device)
    _call(("dp_coeffs", D, n_models), "pgf_dp_coeffs", DP.data_ptr(), ee.data_ptr(), int(fixed), D, n_models, out[0].data_ptr(), out[1].data_ptr(),
           out[2].data_ptr(), _stream())
    return out[0], out[1], out[2]


def perturb_gate_fwd(blocks, w, eps_hat, *, noise_mode, lap=None, gum=None, seed=0, offset=0, row0=0, tau=1.0,
                     hard=True, want_gate=False, out_dtype=torch.float32, out=None, want_gate_idx=False,
                     want_minmax=False, n_models=1, seed_step=0, model_seeds=None, n_rep=1, src_rows=None, cursor_state=None):
    """models.py:69-79 in one kernel.  Returns (out, gate_idx|None, row_min|None, row_max|None).
    Single model: blocks [B,Di], out [B,D].  Grouped (n_models > 1): blocks [B,Di] (shared batch) or
    [M,B,Di]; w/eps_hat [M,D]; out [M,B,D]; lap [M,B,D]; gum [M,2,B,D]; model m uses seed + m*seed_step.
    n_rep > 1: the batch is perturbed n_rep times with offsets offset..offset+n_rep-1 in one launch (the n_eval
    repetitions of train.py:126-131); out then has n_rep*B rows, repetition-major.  src_rows (device int64): the
    batch is rows src_rows[0:B] of the (resident) blocks."""
    blocks = [_chk(b, torch.float32, "feature block") for b in blocks]
    if not 1 <= len(blocks) <= 3:
        raise ValueError("1 to 3 feature blocks expected")
    n_rep = int(n_rep)
    gather = src_rows is not None
    B = int(src_rows.numel()) if gather else blocks[0].shape[-2]
    if gather:
        _chk(src_rows, torch.int64, "src_rows")
        assert src_rows.is_contiguous()
    dims = [b.shape[-1] for b in blocks]
    D = sum(dims)
    dev = blocks[0].device
    M = int(n_models)
    lead = (M,) if M > 1 or (out is not None and out.dim() == 3) else ()
    if out is None:
        out = torch.empty(*lead, n_rep * B, D, dtype=out_dtype, device=dev)
    elif out.shape[-2] != n_rep * B:
        raise ValueError(f"out must have n_rep*B = {n_rep * B} rows, got {out.shape[-2]}")
    gate_idx = torch.empty(*lead, n_rep * B, D, dtype=torch.uint8, device=dev) if (want_gate and want_gate_idx) else None
    rmin = torch.empty(*lead, n_rep * B, dtype=torch.float32, device=dev) if want_minmax else None
    rmax = torch.empty(*lead, n_rep * B, dtype=torch.float32, device=dev) if want_minmax else None
    bl = blocks + [None] * (3 - len(blocks))
    args, sx = [], []
    for b in bl:
        args += [_ptr(b), 0 if b is None else b.shape[-1], 0 if b is None else b.stride(-2)]
        sx.append(0 if b is None or b.dim() == 2 else b.stride(0))
    if lap is not None:
        _chk(lap, torch.float32, "lap")
        assert lap.is_contiguous() and lap.numel() == M * n_rep * B * D
    if gum is not None:
        _chk(gum, torch.float32, "gum")
        assert gum.is_contiguous() and gum.numel() == M * 2 * n_rep * B * D
    s_coef = 0 if w is None or w.dim() == 1 else w.stride(0)
    s_out = out.stride(0) if out.dim() == 3 else 0
    common = (*args, _ptr(w), _ptr(eps_hat), B, noise_mode, _ptr(lap), _ptr(gum), int(seed),
              int(offset) & 0xFFFFFFFF, int(row0), float(tau), int(bool(hard)), int(bool(want_gate)), out.data_ptr(), _dt(out),
              out.stride(-2), _ptr(gate_idx), _ptr(rmin), _ptr(rmax), M, sx[0], sx[1], sx[2], s_coef, s_out, int(seed_step),
              _seeds_ptr(model_seeds, M))
    tag = ("perturb_fwd", n_rep * B, D, M, _dt(out), noise_mode, int(bool(want_gate)), int(M > 1 and not any(sx)))
    if n_rep > 1 or gather or cursor_state is not None:
        _call(tag, "pgf_perturb_gate_fwd_ex", *common, _ptr(cursor_state), _ptr(src_rows), int(gather), n_rep, _stream())
    else:
        _call(tag, "pgf_perturb_gate_fwd", *common, _stream())
    return out, gate_idx, rmin, rmax


def perturb_gate_bwd_dp(dF, deps_dDP, *, noise_mode, lap=None, seed=0, offset=0, row0=0, out=None, accumulate=False,
                        seed_step=0, model_seeds=None):
    """dDP = deps_dDP * sum_b dF * noise.  Autograd of models.py:75-76.  dF [B,D] or [M,B,D]."""
    _chk(dF, None, "dF")
    B, D = dF.shape[-2:]
    M = 1 if dF.dim() == 2 else dF.shape[0]
    if out is None:
        out = torch.empty((D,) if dF.dim() == 2 else (M, D), dtype=torch.float32, device=dF.device)
    nbytes = _query("pgf_perturb_gate_bwd_dp_workspace", B, D, M)
    ws = workspace("perturb_bwd").get(nbytes, dF.device)
    s_dF = 0 if dF.dim() == 2 else dF.stride(0)
    s_coef = 0 if deps_dDP.dim() == 1 else deps_dDP.stride(0)
    s_out = 0 if out.dim() == 1 else out.stride(0)
    _call(("perturb_bwd_dp", B, D, M, _dt(dF)), "pgf_perturb_gate_bwd_dp", dF.data_ptr(), _dt(dF), dF.stride(-2), s_dF, B, D, M, noise_mode, _ptr(lap), int(seed),
           int(seed_step), _seeds_ptr(model_seeds, M), int(offset) & 0xFFFFFFFF, int(row0), deps_dDP.data_ptr(), s_coef,
           ws.data_ptr(), ws.numel() * 4,
           out.data_ptr(), s_out, int(accumulate), _stream())
    return out


def minmax_norm_bwd(blocks, dn):
    """Gradient wrt the raw feature blocks through models.py:70-72."""
    blocks = [_chk(b, torch.float32, "feature block") for b in blocks]
    _chk(dn, None, "dn")
    B = blocks[0].shape[0]
    dxs = [torch.empty_like(b) for b in blocks]
    bl = blocks + [None] * (3 - len(blocks))
    dl = dxs + [None] * (3 - len(dxs))
    args = []
    for b in bl:
        args += [_ptr(b), 0 if b is None else b.shape[1], 0 if b is None else b.stride(0)]
    dargs = []
    for d in dl:
        dargs += [_ptr(d), 0 if d is None else d.stride(0)]
    L.call("pgf_minmax_norm_bwd", *args, dn.data_ptr(), _dt(dn), dn.stride(0), B, *dargs, _stream())
    return dxs


# --------------------------------------------------------------------------------------------
def _grouped(t, nd):
    """(tensor, n_models, model_stride) for a [*, ...] or [n_models, *, ...] tensor."""
    if t.dim() == nd:
        return 1, 0
    assert t.dim() == nd + 1, f"expected {nd} or {nd + 1} dims, got {t.dim()}"
    return t.shape[0], t.stride(0)


def linear_fwd(X, W, bias, act=L.ACT_NONE, out=None):
    """Y = act(X W^T + bias); X [B,K] or [m,B,K], W [N,K] or [m,N,K] (fp32, CUDA-core path)."""
    _chk(X, torch.float32, "X"); _chk(W, torch.float32, "W")
    nm, sW = _grouped(W, 2)
    nx, sX = _grouped(X, 2)
    n_models = max(nm, nx)
    B, K = X.shape[-2:]
    N = W.shape[-2]
    assert W.shape[-1] == K and W.stride(-2) == K, "W must be dense [N,K]"
    if out is None:
        out = torch.empty((*X.shape[:-1], N) if nx > 1 or nm == 1 else (nm, B, N), dtype=torch.float32, device=X.device)
    sY = out.stride(0) if out.dim() == 3 else 0
    sb = 0 if bias is None or bias.dim() == 1 else bias.stride(0)
    _call(("linear_fwd", B, N, K, n_models), "pgf_linear_fwd", X.data_ptr(), X.stride(-2), sX, W.data_ptr(), sW, _ptr(bias), sb, out.data_ptr(),
           out.stride(-2), sY, B, N, K, act, n_models, _stream())
    return out


def linear_bwd_dx(dY, W, mask_src=None, mask_mode=L.ACT_RELU, out=None):
    """dX = dY W, times act'(mask_src) when given (RELU: src>0, TANH: 1-src^2)."""
    _chk(dY, torch.float32, "dY"); _chk(W, torch.float32, "W")
    nm, sW = _grouped(W, 2)
    ny, sdY = _grouped(dY, 2)
    n_models = max(nm, ny)
    B, N = dY.shape[-2:]
    K = W.shape[-1]
    if out is None:
        out = torch.empty((*dY.shape[:-1], K), dtype=torch.float32, device=dY.device)
    sdX = out.stride(0) if out.dim() == 3 else 0
    nbytes = _query("pgf_linear_bwd_dx_workspace", B, N, K, n_models)
    ws = workspace("linear_dx").get(nbytes, dY.device)
    ld_mask = 0 if mask_src is None else mask_src.stride(-2)
    s_mask = 0 if mask_src is None or mask_src.dim() == 2 else mask_src.stride(0)
    _call(("linear_bwd_dx", B, N, K, n_models), "pgf_linear_bwd_dx", dY.data_ptr(), dY.stride(-2), sdY, W.data_ptr(), sW, _ptr(mask_src), mask_mode, ld_mask, s_mask,
           out.data_ptr(), out.stride(-2), sdX, B, N, K, n_models, ws.data_ptr(), ws.numel() * 4, _stream())
    return out


def linear_bwd_dw(dY, X, dW=None, db=None, want_db=True, accumulate=False):
    """dW = dY^T X, db = colsum(dY)."""
    _chk(dY, torch.float32, "dY"); _chk(X, torch.float32, "X")
    ny, sdY = _grouped(dY, 2)
    nx, sX = _grouped(X, 2)
    n_models = max(nx, ny)
    B, N = dY.shape[-2:]
    K = X.shape[-1]
    lead = (n_models,) if (dY.dim() == 3 or X.dim() == 3) else ()
    if dW is None:
        dW = torch.empty((*lead, N, K), dtype=torch.float32, device=dY.device)
    if db is None and want_db:
        db = torch.empty((*lead, N), dtype=torch.float32, device=dY.device)
    sdW = dW.stride(0) if dW.dim() == 3 else 0
    sdb = 0 if db is None or db.dim() == 1 else db.stride(0)
    nbytes = _query("pgf_linear_bwd_dw_workspace", B, N, K, n_models)     # > 0 only for a narrow layer at a large batch
    ws = workspace("linear_dw").get(nbytes, dY.device) if nbytes else None
    _call(("linear_bwd_dw", B, N, K, n_models), "pgf_linear_bwd_dw_ex", dY.data_ptr(), dY.stride(-2), sdY, X.data_ptr(), X.stride(-2), sX, dW.data_ptr(), sdW,
           _ptr(db), sdb, B, N, K, int(accumulate), n_models, _ptr(ws), 0 if ws is None else ws.numel() * 4, _stream())
    return dW, db


def gemm_bf16(A, B, C, *, M, N, K, a_mn=False, b_mn=False, epi=L.EPI_STORE_BF16, bias=None, aux=None, stream_k=False,
              colsum_out=None):
    """C[M,N] = A . B^T on tcgen05 (bf16 in, fp32 accumulate in TMEM) with a fused epilogue.
    colsum_out [N] (bf16-output epilogues): also the column sums of the fp32 epilogue values, i.e. the
    bias gradient when C is a pre-activation gradient -- reduced per 128-row slab inside the epilogue.
    aux: int32 [M, N/32] ReLU sign bits -- EPI_BIAS_RELU_BF16 writes them (optional), EPI_BITMASK_BF16 applies them."""
    _chk(A, torch.bfloat16, "A"); _chk(B, torch.bfloat16, "B"); _chk(C, None, "C")
    if aux is not None:
        _chk(aux, torch.int32, "aux")   # packed ReLU sign bits
    part = None
    if colsum_out is not None:
        rows = _query("pgf_gemm_partial_rows", M)
        part = workspace("gemm_partial").get(rows * N * 4, C.device)
    _call(("gemm", M, N, K, int(a_mn), int(b_mn), epi), "pgf_gemm_bf16", A.data_ptr(), A.stride(0), int(a_mn), B.data_ptr(),
          B.stride(0), int(b_mn), C.data_ptr(), C.stride(0), M, N, K, epi, _ptr(bias), _ptr(aux),
          0 if aux is None else aux.stride(0), int(stream_k), _ptr(part), _stream())
    if colsum_out is not None:
        _chk(colsum_out, torch.float32, "colsum_out")
        _call(("reduce_partials", rows, N), "pgf_reduce_partials", part.data_ptr(), rows, N, None, colsum_out.data_ptr(), 0, _stream())
    return C


def split3(src, *, planes=None, out=None, bias=None, act=L.ACT_NONE, mask_plane=None):
    """Elementwise stage of the fp32-parity tensor path: v = act(src + bias) [* (mask_plane > 0)], written as fp32 (`out`,
    may alias src) and/or as the three bf16 planes [3, R, C] with hi + mid + lo == v exactly (pgf_split3)."""
    _chk(src, torch.float32, "src")
    R, Cn = src.shape
    if planes is not None:
        _chk(planes, torch.bfloat16, "planes")
        assert planes.shape == (3, R, Cn) and planes.stride(2) == 1
    if out is not None:
        _chk(out, torch.float32, "out")
    if mask_plane is not None:
        _chk(mask_plane, torch.bfloat16, "mask_plane")
    _call(("split3", R, Cn, int(act), int(planes is not None), int(out is not None), int(mask_plane is not None)), "pgf_split3",
          src.data_ptr(), src.stride(0), R, Cn, _ptr(bias), int(act), _ptr(mask_plane),
          0 if mask_plane is None else mask_plane.stride(0), _ptr(out), 0 if out is None else out.stride(0), _ptr(planes),
          0 if planes is None else planes.stride(1), 0 if planes is None else planes.stride(0), _stream())
    return planes if planes is not None else out


def gemm_bf16x3(A3, B3, C, *, M, N, K, a_mn=False, b_mn=False, epi=L.EPI_STORE_F32, bias=None, k_slabs=0):
    """C[M,N] (fp32) = A . B^T with both operands given as hi/mid/lo bf16 plane triples [3, rows, cols] (ops.split3): the
    fp32-parity route onto tcgen05 (six plane-pair products in one fp32 accumulation).  k_slabs (EPI_ATOMIC_F32, C zeroed
    by the caller): 1 = automatic split over K, > 1 = exactly that many slabs."""
    _chk(A3, torch.bfloat16, "A3"); _chk(B3, torch.bfloat16, "B3"); _chk(C, torch.float32, "C")
    assert A3.dim() == 3 and B3.dim() == 3 and A3.shape[0] == 3 and B3.shape[0] == 3
    _call(("gemm_x3", M, N, K, int(a_mn), int(b_mn), epi), "pgf_gemm_bf16x3", A3.data_ptr(), A3.stride(1), A3.stride(0), int(a_mn),
          B3.data_ptr(), B3.stride(1), B3.stride(0), int(b_mn), C.data_ptr(), C.stride(0), M, N, K, epi, _ptr(bias), int(k_slabs),
          _stream())
    return C


def gemm_bf16_ddp(A, B, *, M, N, K, b_mn=True, seed, offset, row0, deps_dDP, out, accumulate=False):
    """dDP = deps_dDP * colsum((A . B^T) * Laplace noise): the input-gradient GEMM of fc_layers.0 with the
    dL/dDP reduction fused into its epilogue (the [M,N] product is never written)."""
    _chk(A, torch.bfloat16, "A"); _chk(B, torch.bfloat16, "B")
    _chk(deps_dDP, torch.float32, "deps_dDP"); _chk(out, torch.float32, "dDP")
    rows = _query("pgf_gemm_partial_rows", M)
    ws = workspace("gemm_partial").get(rows * N * 4, A.device)
    _call(("gemm", M, N, K, 0, int(b_mn), L.EPI_DDP_PARTIAL), "pgf_gemm_bf16_ddp", A.data_ptr(), A.stride(0), B.data_ptr(),
          B.stride(0), int(b_mn), M, N, K, int(seed), int(offset) & 0xFFFFFFFF, int(row0), deps_dDP.data_ptr(), ws.data_ptr(),
          ws.numel() * 4, out.data_ptr(), int(accumulate), _stream())
    return out


def cls_ce(h, Wc, bc, labels, *, loss_scale, grad_scale, backward, through_tanh=True, want_logits=True, want_pred=True,
           dz_dtype=None, dz=None, dWc=None, dbc=None, want_dw=True, dz_colsum=None, logits=None, pred=None, stats=None):
    """classifier + mean-CE + accuracy (+ backward).  Returns dict(logits, pred, stats, dz, dWc, dbc).
    want_dw=False (pass 1 of the reference step) forms dz only; dz_colsum [H] / [M,H] receives the
    column sums of dz (the bias gradient of fc_layers.2)."""
    _chk(h, None, "h"); _chk(Wc, torch.float32, "Wc"); _chk(bc, torch.float32, "bc")
    nh, sh = _grouped(h, 2)
    nw, sWc = _grouped(Wc, 2)
    n_models = max(nh, nw)
    B, H = h.shape[-2:]
    dev = h.device
    lead = (n_models,) if (h.dim() == 3 or Wc.dim() == 3) else ()
    if Wc.shape[-2] != 2:
        raise NotImplementedError("the reference head has 2 classes (nn.Linear(768, 2), models.py:52)")
    sbc = 0 if bc.dim() == 1 else bc.stride(0)
    slab = 0
    if labels is not None:
        _chk(labels, torch.int64, "labels")
        slab = labels.stride(0) if labels.dim() == 2 else 0
    if logits is None and want_logits:
        logits = torch.empty((*lead, B, 2), dtype=torch.float32, device=dev)
    if pred is None and want_pred:
        pred = torch.empty((*lead, B), dtype=torch.int64, device=dev)
    if stats is None:
        stats = torch.empty((*lead, 4), dtype=torch.float32, device=dev)
    if backward:
        if dz is None:
            dz = torch.empty(h.shape, dtype=dz_dtype or h.dtype, device=dev)
        if dWc is None and want_dw:
            dWc = torch.empty((*lead, 2, H), dtype=torch.float32, device=dev)
        if dbc is None and want_dw:
            dbc = torch.empty((*lead, 2), dtype=torch.float32, device=dev)
    nbytes = _query("pgf_cls_ce_workspace", B, H, n_models)
    ws = workspace("cls_ce").get(nbytes, dev)
    g3 = lambda t: 0 if t is None or t.dim() < len(lead) + 1 or not lead else t.stride(0)
    _call(("cls_ce", B, H, n_models, _dt(h), (2 if (dWc is not None or dz_colsum is not None) else 1) if backward else 0), "pgf_cls_ce", h.data_ptr(), _dt(h), h.stride(-2), sh, Wc.data_ptr(), sWc, bc.data_ptr(), sbc, _ptr(labels), slab,
           B, H, n_models, float(loss_scale), float(grad_scale), int(bool(backward)), int(bool(through_tanh)),
           _ptr(logits), g3(logits), _ptr(pred), g3(pred), stats.data_ptr(), _ptr(dz), 0 if dz is None else _dt(dz),
           0 if dz is None else dz.stride(-2), 0 if dz is None or dz.dim() == 2 else dz.stride(0), _ptr(dWc), g3(dWc),
           _ptr(dbc), g3(dbc), _ptr(dz_colsum), g3(dz_colsum), ws.data_ptr(), ws.numel() * 4, _stream())
    return dict(logits=logits, pred=pred, stats=stats, dz=dz, dWc=dWc if backward else None, dbc=dbc if backward else None)


def adam_step(p, g, m, v, step, lr=1e-6, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0, bf16_shadow=None):
    """torch.optim.Adam single step over flat fp32 buffers (past_acc.py:157-160,203,212)."""
    for t, n in ((p, "p"), (g, "g"), (m, "m"), (v, "v")):
        _chk(t, torch.float32, n)
        assert t.is_contiguous()
    _call(("adam", p.numel(), int(bf16_shadow is not None)), "pgf_adam_step", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _ptr(bf16_shadow), p.numel(), int(step),
           float(lr), float(betas[0]), float(betas[1]), float(eps), float(grad_scale), _stream())


def adam_step_strided(p, g, m, v, step, lr=1e-6, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
    """Adam over a [n_models, n] strided view (row stride = model stride) of the flat buffers: one launch."""
    for t in (p, g, m, v):
        _chk(t, torch.float32, "adam buffer")
        assert t.dim() == 2 and t.shape == p.shape and t.stride() == p.stride()
    _call(("adam", p.numel(), 0), "pgf_adam_step_strided", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), None, p.shape[1],
          p.stride(0), p.shape[0], int(step), float(lr), float(betas[0]), float(betas[1]), float(eps), float(grad_scale), _stream())


def linear_adam_step(dY, X, W, mW, vW, bias, mb, vb, step, lr=1e-6, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
    """Adam step of one nn.Linear with the weight gradient dY^T X (and bias gradient colsum(dY)) recomputed in
    the optimiser kernel instead of materialised (batch <= 8).  Grouped: dY [M,B,N], X [M,B,K], W/mW/vW [M,N,K]
    views of per-model flat buffers with a common model stride; bias/mb/vb [M,N] views of the same buffers."""
    for t, n in ((dY, "dY"), (X, "X"), (W, "W"), (mW, "mW"), (vW, "vW")):
        _chk(t, torch.float32, n)
    nm, sP = _grouped(W, 2)
    if nm == 1:
        sP = 0   # the stride of a size-1 dimension is meaningless (torch may report anything)
    B, N = dY.shape[-2:]
    K = X.shape[-1]
    assert W.shape[-2:] == (N, K) and W.stride(-2) == K, "W must be dense [N,K]"
    for t in (mW, vW):
        assert t.shape == W.shape and (nm == 1 or t.stride() == W.stride())
    if bias is not None:
        for t in (bias, mb, vb):
            _chk(t, torch.float32, "bias")
            assert nm == 1 or (0 if t.dim() == 1 else t.stride(0)) == sP
    _call(("linear_adam", B, N, K, nm), "pgf_linear_adam_step", dY.data_ptr(), dY.stride(-2), 0 if dY.dim() == 2 else dY.stride(0),
          X.data_ptr(), X.stride(-2), 0 if X.dim() == 2 else X.stride(0), B, N, K, W.data_ptr(), mW.data_ptr(), vW.data_ptr(),
          _ptr(bias), _ptr(mb), _ptr(vb), sP, int(step), float(lr), float(betas[0]), float(betas[1]), float(eps),
          float(grad_scale), nm, _stream())


def fill_zero(t):
    """cudaMemsetAsync on the current stream (a C-ABI call, so that it is part of recorded call plans)."""
    _chk(t, None, "tensor")
    assert t.is_contiguous()
    _call(("fill_zero", t.numel() * t.element_size()), "pgf_fill_zero", t.data_ptr(), t.numel() * t.element_size(), _stream())


def cast_bf16(src, dst=None):
    _chk(src, torch.float32, "src")
    assert src.is_contiguous()
    if dst is None:
        dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    _call(("cast_bf16", src.numel()), "pgf_cast_f32_to_bf16", src.data_ptr(), dst.data_ptr(), src.numel(), _stream())
    return dst


def colsum(x, out=None):
    _chk(x, None, "x")
    B, N = x.shape
    if out is None:
        out = torch.empty(N, dtype=torch.float32, device=x.device)
    nbytes = _query("pgf_colsum_workspace", B, N)
    ws = workspace("colsum").get(nbytes, x.device)
    _call(("colsum", B, N, _dt(x)), "pgf_colsum", x.data_ptr(), _dt(x), x.stride(0), B, N, out.data_ptr(), ws.data_ptr(), ws.numel() * 4, _stream())
    return out


# --------------------------------------------------------------------------------------------
# the older PriGumbel head tail (SURVEY 8 row a-alt; train_val.py:80-123)
def prigumbel_coef(w, *, exp_eps, tau, hard, gumbel=None, seed=0, offset=0, coef=None, wloss=None):
    """Per-forward column coefficients of gumbel_dropout (train_val.py:95-101) and the w term of loss_function
    (train_val.py:88-89).  gumbel [H,2] injects the draw; None = Philox(seed, offset).  Returns (coef [4,H], wloss [2])."""
    _chk(w, torch.float32, "w")
    H = w.shape[0]
    if gumbel is not None:
        _chk(gumbel, torch.float32, "gumbel")
        if tuple(gumbel.shape) != (H, 2) or not gumbel.is_contiguous():
            raise ValueError("gumbel must be a contiguous [H,2] tensor (the shape F.gumbel_softmax draws, train_val.py:99)")
    if coef is None:
        coef = torch.empty((4, H), dtype=torch.float32, device=w.device)
    if wloss is None:
        wloss = torch.empty(2, dtype=torch.float32, device=w.device)
    _call(("prigumbel_coef", H), "pgf_prigumbel_coef", w.data_ptr(), _ptr(gumbel), H, float(exp_eps), float(tau), int(bool(hard)),
          int(seed), int(offset) & 0xFFFFFFFF, coef.data_ptr(), wloss.data_ptr(), _stream())
    return coef, wloss


def prigumbel_fwd(z, coef, *, eps, lap=None, seed=0, offset=0, row0=0, out=None, want_minmax=False):
    """gumbel_dropout + Lap_noise (train_val.py:95-101,114-123) on fc2's output z [B,H].  lap [B] injects the
    Laplace(0,1/eps) row draws; None = Philox(seed, offset, row0 + b)."""
    _chk(z, torch.float32, "z")
    B, H = z.shape
    if lap is not None:
        _chk(lap, torch.float32, "lap")
        if lap.numel() != B or not lap.is_contiguous():
            raise ValueError("lap must hold one Laplace(0,1/eps) draw per row (train_val.py:121)")
    if out is None:
        out = torch.empty((B, H), dtype=torch.float32, device=z.device)
    mn = torch.empty(B, dtype=torch.float32, device=z.device) if want_minmax else None
    mx = torch.empty(B, dtype=torch.float32, device=z.device) if want_minmax else None
    _call(("prigumbel_fwd", B, H), "pgf_prigumbel_fwd", z.data_ptr(), z.stride(0), coef.data_ptr(), _ptr(lap), float(eps), int(seed),
          int(offset) & 0xFFFFFFFF, int(row0), out.data_ptr(), out.stride(0), _ptr(mn), _ptr(mx), B, H, _stream())
    return (out, mn, mx) if want_minmax else out


def prigumbel_bwd(z, coef, dout, *, wloss=None, exp_eps=1.0, wloss_scale=1.0, dz=None, dw=None, want_dw=True, accumulate=False):
    """Autograd of prigumbel_fwd (+ the w term of loss_function): returns (dz [B,H], dw [H] or None)."""
    _chk(z, torch.float32, "z"); _chk(dout, torch.float32, "dout")
    B, H = z.shape
    if dz is None:
        dz = torch.empty((B, H), dtype=torch.float32, device=z.device)
    if dw is None and want_dw:
        dw = torch.zeros(H, dtype=torch.float32, device=z.device) if accumulate else torch.empty(H, dtype=torch.float32, device=z.device)
    nbytes = _query("pgf_prigumbel_bwd_workspace", B, H)
    ws = workspace("prigumbel").get(nbytes, z.device)
    _call(("prigumbel_bwd", B, H), "pgf_prigumbel_bwd", z.data_ptr(), z.stride(0), coef.data_ptr(), dout.data_ptr(), dout.stride(0),
          _ptr(wloss), float(exp_eps), float(wloss_scale), dz.data_ptr(), dz.stride(0), _ptr(dw), int(bool(accumulate)), B, H,
          ws.data_ptr(), ws.numel() * 4, _stream())
    return dz, dw
