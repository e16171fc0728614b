"""GPU tests at BASELINE.json's FULL sizes (configs 2, 3, 4), where the CPU
oracle is too slow to be run per pair: size-independent properties of the
domain plus one independent fp64 computation on the device (torch / cuBLAS on
the widened operands -- a different code path than the kernels under test).

  GEMV 16384^2   row-slab additivity (bit-exact), power-of-two scaling
                 (bit-exact), agreement with a fp64 GEMV of the widened data to
                 one rounding of the storage type; one pair against the oracle
  DOT  2^28      symmetry and scaling (bit-exact), determinism, split
                 additivity and agreement with fp64 torch.dot (tolerance)
  TRSV 16384     residual of the solve, error against the long-double oracle
                 no worse than 3x the reference restatement's
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
NP = {torch.float64: np.float64, torch.float32: np.float32, torch.float16: np.float16}
EPS = {torch.float64: 2.0 ** -53, torch.float32: 2.0 ** -24, torch.float16: 2.0 ** -11}


def fixture(handle, rows, cols, st, first=0):
    t = torch.empty(rows * cols, dtype=st, device=DEV)
    handle.fill_uniform(rows, cols, t, cols, seed=42, first_draw=first)
    return t


@pytest.mark.parametrize("ar", [torch.float64, torch.float32])
@pytest.mark.parametrize("st", [torch.float64, torch.float32, torch.float16])
def test_gemv_config2_properties(handle, ar, st):
    m = n = 16384
    A = fixture(handle, m, n, st)
    x = fixture(handle, n, 1, st, first=m * n)
    y0 = fixture(handle, m, 1, st, first=m * n + n)

    def run(rows, A_part, alpha, beta, y_init):
        y = y_init.clone()
        handle.gemv(ar, rows, n, alpha, A_part, n, x, 1, beta, y, 1)
        return y

    full = run(m, A, 1.0, 0.0, y0)
    # 1. a row's result does not depend on which slab of rows it is computed in
    cuts = [0, 4 * 1021, 8192, 8192 + 4 * 777, m]
    parts = [run(b - a, A[a * n:b * n], 1.0, 0.0, y0[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    assert torch.equal(torch.cat(parts), full)
    # 2. scaling by a power of two commutes with every rounding involved
    doubled = run(m, A, 2.0, 0.0, y0)
    assert torch.equal(doubled.double(), 2.0 * full.double())
    # 3. alpha A x + beta y against fp64 arithmetic on the widened operands
    got = run(m, A, 1.0, 1.0, y0).double()
    want = A.view(m, n).double() @ x.double() + y0.double()
    scale = (A.view(m, n).double().abs() @ x.double().abs()) + y0.double().abs()
    # rounding of the result to storage + accumulation in `ar` (n terms, pairwise-ish)
    bar = EPS[st] * want.abs() + (EPS[ar] * 64 + 2.0 ** -50) * scale
    if st == torch.float16 or (ar == torch.float32 and st == torch.float64):
        # half ulps of fp16 are coarse near binade edges; fp64 data is first
        # narrowed to fp32 by the accessor: allow twice the bar
        bar = 2 * bar
    assert bool(((got - want).abs() <= bar).all()), float(((got - want).abs() / bar).max())
    l1 = float((got - want).abs().sum() / want.abs().sum())
    assert l1 <= {torch.float64: 2e-15, torch.float32: 6e-8, torch.float16: 5e-4}[st] + \
        (4e-7 if ar == torch.float32 else 0.0), l1


GEMV_BAR = {(torch.float64, torch.float64): 2e-15, (torch.float64, torch.float32): 6e-8,
            (torch.float64, torch.float16): 5e-4, (torch.float32, torch.float64): 4e-7,
            (torch.float32, torch.float32): 4e-7, (torch.float32, torch.float16): 5e-4}


@pytest.mark.parametrize("ar", [torch.float64, torch.float32])
@pytest.mark.parametrize("st", [torch.float64, torch.float32, torch.float16])
def test_gemv_config2_against_oracle(oracle, handle, ar, st, request):
    """Config 2 (m = n = 16384, all three storages x both arithmetics) against
    the double-double oracle on sampled rows (the CPU side stays within
    seconds), and -- when the reference's own kernels are on the box -- no worse
    than 1.5x THEIR error on the same rows (cuda/gemv_benchmark.cu:219-232)."""
    m = n = 16384
    A = fixture(handle, m, n, st)
    x = fixture(handle, n, 1, st, first=m * n)
    y0 = fixture(handle, m, 1, st, first=m * n + n)
    y = y0.clone()
    handle.gemv(ar, m, n, 1.0, A, n, x, 1, 1.0, y, 1)
    rows = np.r_[0:8, 4093:4101, 8190:8198, 12283:12291, m - 8:m]
    A_rows = A.view(m, n)[torch.from_numpy(rows).to(DEV)].contiguous().cpu().numpy().reshape(-1)
    exact = oracle.exact_gemv(A_rows, len(rows), n, n, x.cpu().numpy(), 1.0, 1.0,
                              y0.cpu().numpy()[rows])
    err = oracle.l1_rel_error(exact, y.cpu().numpy()[rows])
    assert err <= GEMV_BAR[(ar, st)], err
    from oracle_binding import REF_LIB
    if REF_LIB.exists():
        refk = request.getfixturevalue("refk")
        yr = y0.clone()
        refk.gemv(ar, m, n, 1.0, A, n, x, 1, 1.0, yr, 1)
        refk.sync()
        ref_err = oracle.l1_rel_error(exact, yr.cpu().numpy()[rows])
        assert err <= 1.5 * ref_err + 1e-18, (err, ref_err)


@pytest.mark.parametrize("ar", [torch.float64, torch.float32])
@pytest.mark.parametrize("st", [torch.float64, torch.float32, torch.float16])
def test_dot_config3_properties(oracle, handle, ar, st):
    n = 2 ** 28
    x = fixture(handle, 1, n, st)
    y = fixture(handle, 1, n, st, first=n)
    res = torch.zeros(4, dtype=ar, device=DEV)

    def run(a, b, count, slot):
        handle.dot(ar, count, a, 1, b, 1, res[slot:slot + 1])
        return res[slot].clone()

    d = run(x, y, n, 0)
    assert torch.equal(run(y, x, n, 1), d)            # products commute, same order
    assert torch.equal(run(x, y, n, 1), d)            # run-to-run determinism
    x2 = x * 2                                         # exact in every storage type
    assert torch.equal(run(x2, y, n, 1), 2 * d)       # power-of-two scaling
    # the double-double oracle (OpenMP) on the very same stored data
    xh, yh = x.cpu().numpy(), y.cpu().numpy()
    want = oracle.exact_dot(xh, yh)
    scale = 0.0
    step = 2 ** 26
    for i in range(0, n, step):
        scale += float(torch.dot(x[i:i + step].double().abs(), y[i:i + step].double().abs()))
    del xh, yh
    tol = {torch.float64: 5e-14, torch.float32: 2e-5}[ar] * scale
    assert abs(float(d) - want) <= tol, (float(d), want)
    half = n // 2
    d1 = run(x[:half], y[:half], half, 2)
    d2 = run(x[half:], y[half:], n - half, 3)
    assert abs(float(d1) + float(d2) - float(d)) <= tol


@pytest.mark.parametrize("ar,st", [(torch.float64, torch.float32), (torch.float64, torch.float64),
                                   (torch.float32, torch.float32)])
def test_trsv_config4(oracle, ab, handle, ar, st):
    """n = 16384, lower / unit on the L factor of the partially pivoted LU of
    the uniform(-1,1) fixture (cuda/trsv_memory.cuh:131-168)."""
    n = 16384
    g = fixture(handle, n, n, torch.float64)
    LU, _ = torch.linalg.lu_factor(g.view(n, n))
    del g
    A = LU.contiguous().view(-1).to(st)
    del LU
    b = fixture(handle, n, 1, st, first=n * n)
    x = b.clone()
    handle.trsv(ar, ab.LOWER, ab.UNIT, n, A, n, x, 1)
    # residual in fp64 on the device: (I + strict_lower(A)) x - b
    L = torch.tril(A.view(n, n).double(), diagonal=-1)
    r = L @ x.double() + x.double() - b.double()
    growth = L.abs() @ x.double().abs() + x.double().abs()
    del L
    bar = (EPS[st] * 4 + EPS[ar] * 256) * growth
    assert bool((r.abs() <= bar).all()), float((r.abs() / bar).max())
    # forward error against the long-double oracle, relative to the reference kernel
    A_h, b_h = A.cpu().numpy(), b.cpu().numpy()
    exact = oracle.exact_trsv(A_h, n, n, b_h, False, True)
    err = oracle.l1_rel_error(exact, x.cpu().numpy())
    ref = oracle.ref_trsv(NP[ar], A_h, n, n, b_h, False, True)
    ref_err = oracle.l1_rel_error(exact, ref)
    assert err <= 3.0 * ref_err + 1e-15, (err, ref_err)


# ---------------------------------------------------------------------------
# 64-bit indexing (the reference's plain DOT is int32, cuda/dot_kernels.cuh:89-97)
# ---------------------------------------------------------------------------
def test_dot_beyond_int32(handle):
    """n = 2^31 + 2^20 + 3 halves per operand (8.6 GB): the streaming kernel's
    tile and tail indices are 64-bit.  Checked against fp64 sums of 2^26-element
    chunks; a planted pair of large products beyond index 2^31 must be seen."""
    n = 2 ** 31 + 2 ** 20 + 3
    st = torch.float16
    x = torch.empty(n, dtype=st, device=DEV)
    y = torch.empty(n, dtype=st, device=DEV)
    handle.fill_uniform(1, n, x, n, seed=42, first_draw=0)
    handle.fill_uniform(1, n, y, n, seed=42, first_draw=n)
    x[2 ** 31 + 5] = 100.0
    y[2 ** 31 + 5] = 200.0
    x[n - 1] = -64.0
    y[n - 1] = 32.0
    want = 0.0
    scale = 0.0
    step = 2 ** 26
    for i in range(0, n, step):
        xa, ya = x[i:i + step].double(), y[i:i + step].double()
        want += float(torch.dot(xa, ya))
        scale += float(torch.dot(xa.abs(), ya.abs()))
    for ar, tol in ((torch.float64, 5e-14), (torch.float32, 2e-5)):
        res = torch.zeros(2, dtype=ar, device=DEV)
        handle.dot(ar, n, x, 1, y, 1, res[0:1])
        handle.dot(ar, n, x, 1, y, 1, res[1:2])
        assert torch.equal(res[0], res[1])
        assert abs(float(res[0]) - want) <= tol * scale, (float(res[0]), want)
    # the same through the scalar (strided) kernel on a 2^31+ index range
    res = torch.zeros(1, dtype=torch.float64, device=DEV)
    handle.dot(torch.float64, (n + 1) // 2, x, 2, y, 2, res)     # indices 0, 2, ..., n - 1
    want2 = 0.0
    for i in range(0, n, step):
        want2 += float(torch.dot(x[i:i + step:2].double(), y[i:i + step:2].double()))
    assert abs(float(res) - want2) <= 5e-14 * scale, (float(res), want2)


def test_fill_uniform_beyond_uint32(oracle, handle):
    """rows x cols > 2^32 elements (8.6 GB of halves): windows of the device
    stream against the oracle's closed form at the start, around 2^32 and at the
    very end; the row-major matrix view draws r*cols + c."""
    rows, cols = 65539, 65541            # 4 295 491 599 elements
    total = rows * cols
    assert total > 2 ** 32
    out = torch.empty(total, dtype=torch.float16, device=DEV)
    handle.fill_uniform(rows, cols, out, cols, seed=42, first_draw=7)
    for start in (0, 2 ** 31 - 50, 2 ** 32 - 100, 2 ** 32 + 12345, total - 1000):
        cnt = min(1000, total - start)
        want = oracle.convert(oracle.uniform(cnt, seed=42, first_draw=7 + start), np.float16)
        got = out[start:start + cnt].cpu().numpy()
        assert np.array_equal(got.view(np.uint16), want.view(np.uint16)), start
