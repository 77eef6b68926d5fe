"""ctypes binding of libpgfuse.so (the C ABI in include/pgfuse.h).

This is the only place the product touches native code.  There is NO fallback: if the
shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpgfuse.so")

DT_F32, DT_BF16 = 0, 1
NOISE_INJECTED, NOISE_PHILOX, NOISE_NONE = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2
EPI_STORE_BF16, EPI_BIAS_RELU_BF16, EPI_BIAS_TANH_BF16, _EPI_RETIRED_3, EPI_ATOMIC_F32, EPI_STORE_F32, EPI_BIAS_F32, EPI_BIAS_TANH_F32, EPI_DDP_PARTIAL, EPI_BITMASK_BF16 = range(10)

P, I, LL, F, U32, U64, SZ = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_uint, C.c_ulonglong, C.c_size_t

# name -> (restype, argtypes); mirrors include/pgfuse.h one to one
SIGNATURES = {
    "pgf_version": (I, []),
    "pgf_last_error": (C.c_char_p, []),
    "pgf_num_sms": (I, []),
    "pgf_dp_coeffs": (I, [P, P, I, I, I, P, P, P, P]),
    "pgf_perturb_gate_fwd": (I, [P, I, LL, P, I, LL, P, I, LL, P, P, I, I, P, P, U64, U32, U64, F, I, I, P, I, LL, P, P, P, I, LL, LL, LL, LL, LL, U64, P, P]),
    "pgf_perturb_gate_fwd_ex": (I, [P, I, LL, P, I, LL, P, I, LL, P, P, I, I, P, P, U64, U32, U64, F, I, I, P, I, LL, P, P, P, I, LL, LL, LL, LL, LL, U64, P, P, P, I, I, P]),
    "pgf_perturb_gate_bwd_dp_workspace": (SZ, [I, I, I]),
    "pgf_perturb_gate_bwd_dp": (I, [P, I, LL, LL, I, I, I, I, P, U64, U64, P, U32, U64, P, LL, P, SZ, P, LL, I, P]),
    "pgf_minmax_norm_bwd": (I, [P, I, LL, P, I, LL, P, I, LL, P, I, LL, I, P, LL, P, LL, P, LL, P]),
    "pgf_linear_fwd": (I, [P, LL, LL, P, LL, P, LL, P, LL, LL, I, I, I, I, I, P]),
    "pgf_linear_bwd_dx_workspace": (SZ, [I, I, I, I]),
    "pgf_linear_bwd_dx": (I, [P, LL, LL, P, LL, P, I, LL, LL, P, LL, LL, I, I, I, I, P, SZ, P]),
    "pgf_linear_bwd_dw": (I, [P, LL, LL, P, LL, LL, P, LL, P, LL, I, I, I, I, I, P]),
    "pgf_linear_bwd_dw_workspace": (SZ, [I, I, I, I]),
    "pgf_linear_bwd_dw_ex": (I, [P, LL, LL, P, LL, LL, P, LL, P, LL, I, I, I, I, I, P, SZ, P]),
    "pgf_gemm_bf16": (I, [P, LL, I, P, LL, I, P, LL, I, I, I, I, P, P, LL, I, P, P]),
    "pgf_split3": (I, [P, LL, I, I, P, I, P, LL, P, LL, P, LL, LL, P]),
    "pgf_gemm_bf16x3": (I, [P, LL, LL, I, P, LL, LL, I, P, LL, I, I, I, I, P, I, P]),
    "pgf_gemm_partial_rows": (I, [I]),
    "pgf_set_sm_reserve": (I, [I]),
    "pgf_reduce_partials": (I, [P, I, I, P, P, I, P]),
    "pgf_gemm_bf16_ddp": (I, [P, LL, P, LL, I, I, I, I, U64, U32, U64, P, P, SZ, P, I, P]),
    "pgf_cls_ce_workspace": (SZ, [I, I, I]),
    "pgf_cls_ce": (I, [P, I, LL, LL, P, LL, P, LL, P, LL, I, I, I, F, F, I, I, P, LL, P, LL, P, P, I, LL, LL, P, LL, P, LL, P, LL, P, SZ, P]),
    "pgf_adam_step": (I, [P, P, P, P, P, LL, I, F, F, F, F, F, P]),
    "pgf_adam_step_strided": (I, [P, P, P, P, P, LL, LL, I, I, F, F, F, F, F, P]),
    "pgf_linear_adam_step": (I, [P, LL, LL, P, LL, LL, I, I, I, P, P, P, P, P, P, LL, I, F, F, F, F, F, I, P]),
    "pgf_fill_zero": (I, [P, SZ, P]),
    "pgf_memcpy_peer_async": (I, [P, I, P, I, SZ, P]),
    "pgf_cast_f32_to_bf16": (I, [P, P, LL, P]),
    "pgf_colsum_workspace": (SZ, [I, I]),
    "pgf_colsum": (I, [P, I, LL, I, I, P, P, SZ, P]),
    "pgf_prigumbel_coef": (I, [P, P, I, F, F, I, U64, U32, P, P, P]),
    "pgf_prigumbel_fwd": (I, [P, LL, P, P, F, U64, U32, U64, P, LL, P, P, I, I, P]),
    "pgf_prigumbel_bwd_workspace": (SZ, [I, I]),
    "pgf_prigumbel_bwd": (I, [P, LL, P, P, LL, P, F, F, P, LL, P, I, I, I, P, SZ, P]),
}



class SweepDesc(C.Structure):
    """pgf_sweep_desc (include/pgfuse.h), field for field."""
    _fields_ = [(n, I) for n in ("n_models", "B", "d0", "d1", "d2", "H", "dp_pass", "fixed_formula", "use_pdl", "reserved_")] + \
               [(n, F) for n in ("tau", "lr", "beta1", "beta2", "adam_eps", "reserved_f_")] + \
               [("x0", P), ("ld0", LL), ("x1", P), ("ld1", LL), ("x2", P), ("ld2", LL), ("labels", P), ("src_rows", P), ("n_rows", LL),
                ("params", P), ("adam_m", P), ("adam_v", P), ("grads", P), ("P", LL)] + \
               [(n, LL) for n in ("off_W1", "off_b1", "off_W2", "off_b2", "off_Wc", "off_bc")] + \
               [("DP", P), ("DP_m", P), ("DP_v", P), ("dDP", P), ("coef", P), ("exp_eps", P), ("seeds", P), ("row0", U64),
                ("stats_dp", P), ("stats_model", P), ("logits", P), ("pred", P), ("state", P), ("workspace", P), ("workspace_bytes", SZ)]


SIGNATURES.update({
    "pgf_step_state_set": (I, [P, LL, LL, LL, LL, F, F, F, P]),
    "pgf_sweep_plan_workspace": (SZ, [I, I, I, I]),
    "pgf_sweep_plan_create": (I, [C.POINTER(SweepDesc), C.POINTER(P)]),
    "pgf_sweep_plan_reset": (I, [P, P]),
    "pgf_sweep_plan_capture": (I, [P, P, I]),
    "pgf_sweep_plan_run": (I, [P, P, I]),
    "pgf_sweep_plan_launches_per_step": (I, [P]),
    "pgf_sweep_plan_destroy": (I, [P]),
})

_lib = None


def load() -> C.CDLL:
    """Load libpgfuse.so; raises if it has not been built (python -m eeg_multimodal_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
                "Build it with `python -m eeg_multimodal_b200.build` or `__graft_entry__.build()`.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


# kernels launched per successful call (for the bench's `gpu_launches` count)
WIDE_MODELS = 24     # PGF_WIDE_MODELS in csrc/pgf_kernels.cuh
LAUNCHES_PER_CALL = {"pgf_memcpy_peer_async": 0, "pgf_perturb_gate_bwd_dp": 2, "pgf_gemm_bf16_ddp": 2, "pgf_cls_ce": 2, "pgf_colsum": 2, "pgf_prigumbel_bwd": 2}
launch_count = 0
launch_by_name = {}


def launches_of(name: str, args) -> int:
    """Kernels one successful call launches (pgf_cls_ce / pgf_perturb_gate_bwd_dp: the finalize launch exists only when a
    model spans several CTAs / slabs, i.e. B > 8 / B > 32)."""
    if name == "pgf_cls_ce" and args[10] <= 8:
        return 1
    if name == "pgf_perturb_gate_bwd_dp" and args[4] <= 32:      # B <= 32: one slab, no finalize launch
        return 1
    if name == "pgf_linear_bwd_dw_ex":   # narrow layer at a large batch: slab kernel + the reductions of dW (and db) per model
        B, N, n_models, has_db = args[10], args[11], args[14], args[8] is not None
        return 1 + n_models * (2 if has_db else 1) if (N <= 8 and B >= 512 and args[16] > 0) else 1
    if name == "pgf_linear_bwd_dx" and args[12] <= 8 and args[15] >= WIDE_MODELS:   # slab kernel + finalize (linear_wide.cu)
        return 2
    return LAUNCHES_PER_CALL.get(name, 1)


def call(name: str, *args):
    """Call an int-returning entry point and raise on a non-zero status."""
    global launch_count
    rc = getattr(load(), name)(*args)
    n = launches_of(name, args)
    launch_count += n
    launch_by_name[name] = launch_by_name.get(name, 0) + n
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {load().pgf_last_error().decode()}")


def query(name: str, *args) -> int:
    return int(getattr(load(), name)(*args))
