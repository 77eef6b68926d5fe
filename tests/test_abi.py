"""CPU: the C-ABI shared library builds, loads, and exports every symbol include/pgfuse.h
declares, with the signature table of the ctypes binding in sync.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pgfuse.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pgf_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from eeg_multimodal_b200 import _lib, build

    build.build()
    return _lib.load()


def test_header_declares_the_hot_path(lib):
    syms = _declared_symbols()
    for must in ("pgf_perturb_gate_fwd", "pgf_perturb_gate_bwd_dp", "pgf_linear_fwd", "pgf_gemm_bf16", "pgf_cls_ce",
                 "pgf_adam_step", "pgf_last_error", "pgf_version"):
        assert must in syms


def test_every_declared_symbol_is_exported(lib):
    raw = ctypes.CDLL(os.path.join(ROOT, "eeg_multimodal_b200", "libpgfuse.so"))
    for s in _declared_symbols():
        assert hasattr(raw, s), f"{s} declared in pgfuse.h but not exported"


def test_binding_table_matches_header(lib):
    from eeg_multimodal_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    # argument counts agree with the prototypes
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "pgfuse.h")).read(), flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, text, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), f"{name}: header has {n} parameters, binding {len(args)}"


def test_version_and_error_string(lib):
    assert lib.pgf_version() >= 100
    assert isinstance(lib.pgf_last_error(), bytes)


def test_argument_errors_are_reported_without_a_gpu(lib):
    from eeg_multimodal_b200 import _lib

    # NULL DP -> argument error before any CUDA call
    rc = lib.pgf_dp_coeffs(None, None, 1, 16, 1, None, None, None, None)
    assert rc != 0 and b"pgf_dp_coeffs" in lib.pgf_last_error()
    with pytest.raises(RuntimeError, match="pgf_adam_step"):
        _lib.call("pgf_adam_step", None, None, None, None, None, 8, 0, 1e-6, 0.9, 0.999, 1e-8, 1.0, None)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "eeg_multimodal_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), f"{fn} references oracle/"


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from eeg_multimodal_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_cpu_tensors_are_rejected():
    import torch

    from eeg_multimodal_b200 import ops

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.dp_coeffs(torch.zeros(16), 2.7)


def test_gemm_object_is_tcgen05_tma_code():
    """The large-batch GEMM is tcgen05 / TMEM / TMA code for sm_100a, not a recompiled mma.sync kernel: the SASS of the built
    object carries UTCHMMA (tcgen05.mma, also in its 2-CTA form), LDTM (tcgen05.ld), UTMALDG / UTMASTG / UTMAREDG (TMA loads,
    stores, reduce-adds; 2-D and -- for the fp32-parity plane walk -- 3-D) and no HMMA (profiles/sass_gemm_tc.txt)."""
    import shutil
    import subprocess

    obj = os.path.join(ROOT, "eeg_multimodal_b200", "build", "gemm_tc.o")
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(obj) or not os.path.exists(tool):
        pytest.skip("gemm_tc.o / cuobjdump not available (the library was not built in this tree)")
    sass = subprocess.run([tool, "-sass", obj], capture_output=True, text=True, check=True).stdout
    for needle in ("UTCHMMA.2CTA", "UTCHMMA", "LDTM", "UTMALDG.2D.2CTA", "UTMALDG.3D", "UTMASTG.2D", "UTMAREDG.2D.ADD"):
        assert needle in sass, needle
    assert "HMMA." not in sass.replace("UTCHMMA", "")
    assert "sm_100a" in subprocess.run([tool, "-lelf", obj], capture_output=True, text=True).stdout
