"""GPU tests against the REFERENCE'S OWN CUDA kernels (oracle/_ref, compiled
from /root/reference/cuda where it lies; skipped when that library was not
built).  They pin the CPU oracle's order-faithful restatement bit for bit and
hold the new kernels to the reference's error."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
NP = {torch.float64: np.float64, torch.float32: np.float32, torch.float16: np.float16}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def stored(oracle, count, st, seed=42, first=0):
    return oracle.convert(oracle.uniform(count, seed=seed, first_draw=first), NP[st])


@pytest.mark.parametrize("ar,st,plain", [
    (torch.float64, torch.float64, True), (torch.float32, torch.float32, True),
    (torch.float64, torch.float64, False), (torch.float64, torch.float32, False),
    (torch.float32, torch.float32, False), (torch.float64, torch.float16, False)])
@pytest.mark.parametrize("m,n,lda", [(100, 100, 24500), (333, 1000, 1000), (64, 5000, 5008)])
def test_oracle_restates_reference_gemv_bitwise(oracle, refk, ar, st, plain, m, n, lda):
    A = stored(oracle, m * lda, st)
    x = stored(oracle, n, st, first=m * lda)
    y = stored(oracle, m, st, first=m * lda + n)
    yd = dev(y)
    refk.gemv(ar, m, n, 1.0, dev(A), lda, dev(x), 1, 1.0, yd, 1, plain=plain)
    refk.sync()
    want = oracle.ref_gemv(NP[ar], A, m, n, lda, x, 1.0, 1.0, y)
    assert np.array_equal(yd.cpu().numpy().view(np.uint8), want.view(np.uint8))


@pytest.mark.parametrize("ar,st", [(torch.float64, torch.float64),
                                   (torch.float64, torch.float32),
                                   (torch.float32, torch.float32)])
def test_oracle_restates_reference_dot(oracle, refk, ar, st):
    """Block partials are combined by atomics in arbitrary order on the GPU, so
    the totals agree to rounding of that last step only; with fp64 arithmetic on
    fp32 storage the per-thread sums are exact products and the totals agree to
    a few ulps."""
    n = 1_000_000
    x = stored(oracle, n, st)
    y = stored(oracle, n, st, first=n)
    res = torch.full((1,), -999.0, dtype=ar, device=DEV)
    refk.dot(ar, n, dev(x), 1, dev(y), 1, res)
    refk.sync()
    want, partials = oracle.ref_dot(NP[ar], x, y, NP[ar], blocks=refk.sm_count() * 32)
    got = float(res.item())
    scale = np.abs(partials.astype(np.float64)).sum()
    eps = 2.3e-16 if ar == torch.float64 else 1.2e-7
    assert abs(got - float(want)) <= 4 * eps * scale


@pytest.mark.parametrize("ar,st,plain", [(torch.float64, torch.float64, True),
                                         (torch.float64, torch.float32, False),
                                         (torch.float32, torch.float32, False)])
@pytest.mark.parametrize("n", [32, 100, 1000])
@pytest.mark.parametrize("upper,unit", [(True, True), (False, True), (True, False),
                                        (False, False)])
def test_oracle_restates_reference_trsv(oracle, refk, ar, st, plain, n, upper, unit):
    from test_gpu_parity import lu_fixture
    LU = lu_fixture(n, seed=200 + n).reshape(n, n)
    if (upper and unit) or (not upper and not unit):
        LU = LU.T.copy()
    A = oracle.convert(LU.reshape(-1), NP[st])
    b = stored(oracle, n, st, seed=3)
    xd = dev(b)
    refk.trsv(ar, upper, unit, n, dev(A), n, xd, 1, plain=plain)
    refk.sync()
    got = xd.cpu().numpy()
    want = oracle.ref_trsv(NP[ar], A, n, n, b, upper, unit)
    if np.array_equal(got.view(np.uint8), want.view(np.uint8)):
        return
    # FMA contraction choices of nvcc inside the inversion may differ from the
    # restatement by an ulp here and there; the solutions must still agree to
    # rounding level relative to the accuracy oracle
    exact = oracle.exact_trsv(A, n, n, b, upper, unit)
    e_got = oracle.l1_rel_error(exact, got)
    e_want = oracle.l1_rel_error(exact, want)
    assert e_got <= 1.5 * e_want + 1e-15 and e_want <= 1.5 * e_got + 1e-15, (e_got, e_want)


@pytest.mark.parametrize("st", [torch.float64, torch.float32, torch.float16])
@pytest.mark.parametrize("ar", [torch.float64, torch.float32])
def test_new_gemv_no_worse_than_reference_kernel(oracle, refk, handle, ar, st):
    m = n = 2000
    A = stored(oracle, m * n, st)
    x = stored(oracle, n, st, first=m * n)
    y = stored(oracle, m, st, first=m * n + n)
    exact = oracle.exact_gemv(A, m, n, n, x, 1.0, 1.0, y)
    y_ref, y_new = dev(y), dev(y)
    refk.gemv(ar, m, n, 1.0, dev(A), n, dev(x), 1, 1.0, y_ref, 1)
    refk.sync()
    handle.gemv(ar, m, n, 1.0, dev(A), n, dev(x), 1, 1.0, y_new, 1)
    torch.cuda.synchronize()
    e_ref = oracle.l1_rel_error(exact, y_ref.cpu().numpy())
    e_new = oracle.l1_rel_error(exact, y_new.cpu().numpy())
    assert e_new <= 1.25 * e_ref + 1e-16, (e_new, e_ref)


@pytest.mark.parametrize("ar,st", [(torch.float64, torch.float32), (torch.float32, torch.float32),
                                   (torch.float64, torch.float64)])
def test_new_trsv_no_worse_than_reference_kernel(oracle, refk, ab, handle, ar, st):
    from test_gpu_parity import lu_fixture
    n = 2048
    LU = lu_fixture(n, seed=31)
    A = oracle.convert(LU, NP[st])
    b = stored(oracle, n, st, seed=4)
    exact = oracle.exact_trsv(A, n, n, b, False, True)
    x_ref, x_new = dev(b), dev(b)
    refk.trsv(ar, False, True, n, dev(A), n, x_ref, 1)
    refk.sync()
    handle.trsv(ar, ab.LOWER, ab.UNIT, n, dev(A), n, x_new, 1)
    torch.cuda.synchronize()
    e_ref = oracle.l1_rel_error(exact, x_ref.cpu().numpy())
    e_new = oracle.l1_rel_error(exact, x_new.cpu().numpy())
    assert e_new <= 2.0 * e_ref + 1e-15, (e_new, e_ref)
