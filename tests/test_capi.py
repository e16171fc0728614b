"""The C-ABI library loads and exports every symbol include/accblas.h declares
(no compute calls here: this runs without a GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "accblas.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(accblas_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(ab):
    from accessor_blas_b200 import capi
    lib = capi.load()
    names = declared_symbols()
    assert len(names) >= 17
    for name in names:
        assert hasattr(lib, name), f"{name} declared in accblas.h but not exported"
    # and the binding table covers exactly the header
    assert sorted(capi.SYMBOLS) == names


def test_no_torch_or_oracle_dependency():
    """The product library is plain CUDA: it must not link torch, the oracle or
    the reference shim."""
    import subprocess
    lib = ROOT / "accessor-blas_b200" / "libaccblas_b200.so"
    out = subprocess.run(["ldd", str(lib)], capture_output=True, text=True).stdout
    for forbidden in ("torch", "oracle", "ref_kernels", "cublas"):
        assert forbidden not in out, out


def test_version_and_strings(ab):
    from accessor_blas_b200 import capi
    lib = capi.load()
    assert lib.accblas_version() == 100
    assert lib.accblas_status_string(0) == b"ok"
    assert lib.accblas_sizeof(capi.F16) == 2
    assert lib.accblas_sizeof(capi.F64) == 8


def test_create_fails_loudly_without_gpu(ab):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from accessor_blas_b200 import capi
    lib = capi.load()
    h = ctypes.c_void_p()
    rc = lib.accblas_create(ctypes.byref(h), 0)
    assert rc == 3  # ACCBLAS_ERR_CUDA: no CPU fallback
    assert b"no CPU path" in lib.accblas_last_error()
    with pytest.raises(ab.AccblasError):
        ab.Handle(0)


def test_product_sources_do_not_touch_the_oracle():
    """oracle/ is test infrastructure: nothing under the package or include/
    may include, import or link it."""
    for path in list((ROOT / "accessor-blas_b200").rglob("*")) + \
            list((ROOT / "include").rglob("*")):
        if path.is_file() and path.suffix in {".py", ".cu", ".cuh", ".h", ".hpp", ".cpp"}:
            text = path.read_text()
            assert "liboracle" not in text and "oracle_binding" not in text and \
                "oracle/" not in text.replace("oracle/ is", ""), path
