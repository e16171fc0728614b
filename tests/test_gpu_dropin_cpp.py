"""The C++ drop-in layer (include/accblas/*.cuh) and the benchmark drivers,
run as the compiled programs a user of the reference would run."""
import re
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

BIN = Path(__file__).resolve().parent.parent / "accessor-blas_b200" / "bin"
NUM = r"[+-]?\d\.\d{16}e[+-]\d{2}"


def run(name, *args, timeout=600):
    exe = BIN / name
    if not exe.exists():
        pytest.fail(f"{exe} missing: run __graft_entry__.build()")
    out = subprocess.run([str(exe), *args], capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


def test_dropin_headers_host_launchers_and_kernel_templates():
    assert run("dropin_test").strip().endswith("PASS")


def test_gemv_driver_csv_matches_reference_format():
    out = run("gemv_benchmark", "--size=300").splitlines()
    assert out[0] == ("Num rows;GEMV fp64;GEMV fp32;GEMV Acc<fp64, fp64>;GEMV Acc<fp64, fp32>;"
                      "GEMV Acc<fp32, fp32>;CUBLAS GEMV fp64;CUBLAS GEMV fp32")
    assert [l.split(";")[0] for l in out[1:]] == ["+100", "+200", "+300"]
    for line in out[1:]:
        cells = line.split(";")
        assert len(cells) == 8 and all(re.fullmatch(NUM, c) for c in cells[1:]), line
    err = run("gemv_benchmark", "--error", "--size=300").splitlines()
    assert err[0].startswith("Num rows;Error GEMV fp64;Error GEMV fp32;")
    row = [float(c) for c in err[3].split(";")]
    # variant 0 is the reference of the metric; Acc<fp64,fp64> is the same kernel
    assert row[1] == 0.0 and row[3] == 0.0
    assert 1e-8 < row[4] < 6e-8          # Acc<fp64,fp32>: the reference plots ~4e-8
    assert 2e-8 < row[2] < 3e-7 and 2e-8 < row[5] < 3e-7
    assert row[6] < 1e-14                # cuBLAS fp64


def test_dot_driver_csv_matches_reference_format():
    out = run("dot_benchmark", "--size=3000000").splitlines()
    names = ["DOT fp64", "DOT fp32", "DOT Acc<fp64, fp64>", "DOT Acc<fp64, fp32>",
             "DOT Acc<fp32, fp32>", "CUBLAS DOT fp64", "CUBLAS DOT fp32"]
    assert out[0] == "Vector Size;" + ";".join(names) + ";" + ";".join("Error " + n for n in names)
    assert [l.split(";")[0] for l in out[1:]] == ["1000000", "3000000"]
    assert all(len(l.split(";")) == 15 for l in out[1:])
    err = run("dot_benchmark", "--error", "--size=1000000").splitlines()
    assert err[0] == "Vector Size;" + ";".join("Error " + n for n in names)
    assert err[2].startswith("-----")
    assert err[3].startswith("Random iter;Vector Size;Result DOT fp64")
    assert len(err) == 4 + 10            # ten re-randomised rounds
    med = [float(c) for c in err[1].split(";")]
    assert med[1] == 0.0 and med[4] < 2e-7


def test_trsv_driver_csv_matches_reference_format():
    out = run("trsv_benchmark", "--size=300").splitlines()
    assert out[0] == ("Num rows;TRSV fp64;TRSV fp32;TRSV Acc<fp64, fp64>;TRSV Acc<fp64, fp32>;"
                      "TRSV Acc<fp32, fp32>;CUBLAS TRSV fp64;CUBLAS TRSV fp32")
    assert [l.split(";")[0] for l in out[1:]] == ["+100", "+200", "+300"]
    for flags in ((), ("--lower",)):
        err = run("trsv_benchmark", "--error", "--size=300", *flags).splitlines()
        row = [float(c) for c in err[3].split(";")]
        assert row[1] == 0.0 and row[3] == 0.0
        assert row[4] < 5e-6 and row[2] < 1e-3 and row[6] < 1e-11, row


# ---------------------------------------------------------------------------
# the reference's OWN, unmodified drivers on top of the drop-in headers
# (oracle/_ref/bin, built by `make -C oracle dropin` where /root/reference
# exists; see tests/test_dropin_build.py)
# ---------------------------------------------------------------------------
REF_BIN = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "bin"


def run_ref(name, *args, timeout=900):
    exe = REF_BIN / name
    if not exe.exists():
        pytest.skip(f"{exe} not present (built only where the reference sources are)")
    out = subprocess.run([str(exe), *args], capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


def _rows(text):
    lines = text.splitlines()
    return lines[0], [[float(c) for c in l.split(";")] for l in lines[1:] if l and l[0] in "+-0123456789"]


@pytest.mark.parametrize("name,size", [("gemv_benchmark", "300"), ("trsv_benchmark", "300")])
def test_reference_driver_on_dropin_headers_matches_repo_driver(name, size):
    """Same flags, same fixtures (the repo's drivers generate the reference's
    host stream on the device): identical CSV header, identical row labels,
    and the same error columns -- both programs end up in the same kernels."""
    ref_head, ref_rows = _rows(run_ref(name, "--error", f"--size={size}"))
    own_head, own_rows = _rows(run(name, "--error", f"--size={size}"))
    assert ref_head == own_head
    assert [r[0] for r in ref_rows] == [r[0] for r in own_rows]
    for a, b in zip(ref_rows, own_rows):
        for col, (va, vb) in enumerate(zip(a[1:], b[1:]), start=1):
            # accblas columns: same kernel on the same data; cuBLAS columns and the
            # TRSV fixture (LU by cuSOLVER vs torch-free device path) to rounding
            assert abs(va - vb) <= 0.5 * max(abs(va), abs(vb)) + 1e-15, (name, col, va, vb)
    timing_head, timing_rows = _rows(run_ref(name, f"--size={size}"))
    assert timing_head == _rows(run(name, f"--size={size}"))[0]
    assert all(v > 0 for r in timing_rows for v in r[1:])


def test_reference_dot_driver_on_dropin_headers():
    ref = run_ref("dot_benchmark", "--error", "--size=1000000").splitlines()
    own = run("dot_benchmark", "--error", "--size=1000000").splitlines()
    assert ref[0] == own[0] and ref[2].startswith("-----") and ref[3] == own[3]
    assert len(ref) == len(own)
    med_ref = [float(c) for c in ref[1].split(";")]
    med_own = [float(c) for c in own[1].split(";")]
    assert med_ref[0] == med_own[0]
    # Acc<fp64,fp32> median error of ten re-randomised rounds: same data, same kernel
    assert abs(med_ref[4] - med_own[4]) <= 0.5 * max(med_ref[4], med_own[4]) + 1e-15
