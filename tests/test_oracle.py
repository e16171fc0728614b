"""CPU tests of the oracle itself: pins it against the reference's generator
(libstdc++), CUDA's own fp16 host conversions and the reference's published
error curves.  No GPU needed."""
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden"


# --- input stream ------------------------------------------------------------
def test_uniform_first_values_match_survey(oracle):
    # first four uniform(-1,1) doubles for seed 42
    want = [0.049174203583507659, -0.47338891843144348, -0.60742834883854191,
            0.024636217393456228]
    for closed in (False, True):
        got = oracle.uniform(4, seed=42, closed_form=closed)
        assert got.tolist() == want


def test_closed_form_equals_libstdcxx(oracle):
    n = 200_000
    a = oracle.uniform(n, seed=42, closed_form=False)
    b = oracle.uniform(n, seed=42, closed_form=True)
    assert np.array_equal(a, b)
    assert a.min() >= -1.0 and a.max() < 1.0


@pytest.mark.parametrize("seed,first", [(42, 12345), (1, 10**9 + 7), (2**31 - 1, 5),
                                        (0, 0), (42, 24500 * 24500)])
def test_closed_form_jump_ahead(oracle, seed, first):
    a = oracle.uniform(1000, seed=seed, first_draw=first, closed_form=False)
    b = oracle.uniform(1000, seed=seed, first_draw=first, closed_form=True)
    assert np.array_equal(a, b)


# --- fp16 ----------------------------------------------------------------------
def test_fp16_conversions_match_cuda_host_routines(oracle):
    g = np.load(GOLDEN / "fp16_cuda_host.npz")
    x = g["x"]
    h64 = oracle.convert(x, np.float16).view(np.uint16)
    assert np.array_equal(h64, g["half_from_f64"])
    h32 = oracle.convert(x.astype(np.float32), np.float16).view(np.uint16)
    assert np.array_equal(h32, g["half_from_f32"])
    back = oracle.convert(g["half_from_f64"].view(np.float16), np.float32)
    assert np.array_equal(back.view(np.uint32), g["f32_from_half"].view(np.uint32))
    wide = oracle.convert(g["half_from_f64"].view(np.float16), np.float64)
    assert np.array_equal(wide, g["f32_from_half"].astype(np.float64))


def test_fp16_single_rounding_probe(oracle):
    # via float this would be 0x3c00
    v = np.array([1 + 2.0 ** -11 + 2.0 ** -30])
    assert oracle.convert(v, np.float16).view(np.uint16)[0] == 0x3C01
    assert oracle.convert(v.astype(np.float32), np.float16).view(np.uint16)[0] == 0x3C00


def test_convert_matches_numpy_casts(oracle):
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, 10000)
    assert np.array_equal(oracle.convert(x, np.float32), x.astype(np.float32))
    assert np.array_equal(oracle.convert(x.astype(np.float32), np.float64),
                          x.astype(np.float32).astype(np.float64))
    # numpy's double -> half is a single RNE rounding as well
    assert np.array_equal(oracle.convert(x, np.float16).view(np.uint16),
                          x.astype(np.float16).view(np.uint16))


# --- reference fixtures: draw order of the drivers -----------------------------
def gemv_fixture(oracle, n, max_size, seed=42):
    """Sub-problem of size n of the reference GEMV fixture generated for
    --size=max_size: matrix (stride max_size), x, res drawn in that order
    (cuda/gemv_memory.cuh:39-55, cuda/gemv_benchmark.cu:213-217)."""
    A = np.empty((n, n), dtype=np.float64)
    for r in range(n):
        A[r] = oracle.uniform(n, seed=seed, first_draw=r * max_size)
    x = oracle.uniform(n, seed=seed, first_draw=max_size * max_size)
    res = oracle.uniform(n, seed=seed, first_draw=max_size * max_size + max_size)
    return A, x, res


def test_gemv_error_curves_reproduce_published_values(oracle):
    """at n=1000 (default --size=24500) the reference
    plots 3.8e-8..4.2e-8 for Acc<fp64,fp32> and 9.5e-8 for plain fp32, both
    relative to its own fp64 kernel.  The order-faithful restatement of those
    kernels must land on the same numbers (digitisation accuracy ~1 %)."""
    n, N = 1000, 24500
    A, x, res = gemv_fixture(oracle, n, N)
    ref = oracle.ref_gemv(np.float64, A.reshape(-1), n, n, n, x, 1.0, 1.0, res)
    A32, x32, res32 = (v.astype(np.float32) for v in (A, x, res))
    acc = oracle.ref_gemv(np.float64, A32.reshape(-1), n, n, n, x32, 1.0, 1.0, res32)
    plain = oracle.ref_gemv(np.float32, A32.reshape(-1), n, n, n, x32, 1.0, 1.0, res32)
    e_acc = oracle.l1_rel_error(ref, acc)
    e_plain = oracle.l1_rel_error(ref, plain)
    assert 3.6e-8 < e_acc < 4.4e-8, e_acc
    assert 8.5e-8 < e_plain < 1.05e-7, e_plain
    # Acc<fp64,fp64> is the same arithmetic as the plain fp64 kernel
    acc64 = oracle.ref_gemv(np.float64, A.reshape(-1), n, n, n, x, 1.0, 1.0, res)
    assert np.array_equal(acc64, ref)


def test_ref_gemv_close_to_exact(oracle):
    rng = np.random.default_rng(3)
    m, n, lda = 37, 1500, 1600
    A = rng.uniform(-1, 1, m * lda).astype(np.float32)
    x = rng.uniform(-1, 1, n).astype(np.float32)
    y = rng.uniform(-1, 1, m).astype(np.float32)
    exact = oracle.exact_gemv(A, m, n, lda, x, 0.5, -2.0, y)
    want = (0.5 * (A.reshape(m, lda)[:, :n].astype(np.float64) @ x.astype(np.float64))
            - 2.0 * y)
    assert np.allclose(exact, want, rtol=1e-13, atol=1e-13)
    got = oracle.ref_gemv(np.float64, A, m, n, lda, x, 0.5, -2.0, y)
    assert got.dtype == np.float32
    assert np.array_equal(got, exact.astype(np.float32)) or \
        oracle.l1_rel_error(exact, got) < 6e-8
    # beta == 0 must not read y (NaN there stays out of the result)
    ynan = np.full(m, np.nan, dtype=np.float32)
    got0 = oracle.ref_gemv(np.float64, A, m, n, lda, x, 1.0, 0.0, ynan)
    assert np.isfinite(got0).all()


@pytest.mark.parametrize("st", [np.float64, np.float32, np.float16])
def test_ref_dot_close_to_exact(oracle, st):
    n = 100_003
    x = oracle.convert(oracle.uniform(n, seed=42), st)
    y = oracle.convert(oracle.uniform(n, seed=42, first_draw=n), st)
    exact = oracle.exact_dot(x, y)
    want = float(np.dot(x.astype(np.float64), y.astype(np.float64)))
    assert abs(exact - want) < 1e-10
    got, partials = oracle.ref_dot(np.float64, x, y, np.float64, blocks=148 * 32)
    assert abs(got - exact) < 1e-12 * max(1, abs(exact)) + 1e-13
    assert partials.shape == (148 * 32,)
    got32, _ = oracle.ref_dot(np.float32, x, y, np.float32, blocks=148 * 32)
    assert abs(float(got32) - exact) < 2e-5 * np.sqrt(n)


def lu_triangles(n, seed):
    """Unit-lower L and upper U of a partially pivoted LU of a uniform(-1,1)
    matrix -- the conditioning of the reference's TRSV fixture
    (cuda/trsv_memory.cuh:131-168)."""
    import scipy.linalg
    rng = np.random.default_rng(seed)
    M = rng.uniform(-1, 1, (n, n))
    _, L, U = scipy.linalg.lu(M)
    return L, U


@pytest.mark.parametrize("n", [1, 31, 32, 33, 100, 257])
@pytest.mark.parametrize("upper,unit", [(False, True), (True, False), (False, False),
                                        (True, True)])
def test_ref_trsv_close_to_exact(oracle, n, upper, unit):
    L, U = lu_triangles(n, seed=n)
    T = (L.T if upper else L).copy() if unit else (U if upper else U.T).copy()
    T = T + (np.triu(np.ones((n, n)), 1).T if upper else np.triu(np.ones((n, n)), 1)) * 7.0
    rng = np.random.default_rng(n + 1)
    b = rng.uniform(-1, 1, n)
    for st, tol in ((np.float64, 1e-11), (np.float32, 5e-5)):
        Ts = T.astype(st).reshape(-1)
        bs = b.astype(st)
        exact = oracle.exact_trsv(Ts, n, n, bs, upper, unit)
        # the other triangle (filled with 7s) must never be read
        tri = np.triu(Ts.reshape(n, n).astype(np.float64)) if upper else \
            np.tril(Ts.reshape(n, n).astype(np.float64))
        if unit:
            np.fill_diagonal(tri, 1.0)
        want = np.linalg.solve(tri, bs.astype(np.float64))
        scale = np.abs(want).sum()
        assert np.abs(exact - want).sum() <= 1e-9 * scale
        got = oracle.ref_trsv(np.float64, Ts, n, n, bs, upper, unit)
        assert oracle.l1_rel_error(exact, got) < tol * 50
        base = oracle.cpu_trsv(np.float64, Ts, n, n, bs.copy(), upper, unit)
        assert oracle.l1_rel_error(exact, base) < tol * 50


def test_l1_rel_error_metric(oracle):
    ref = np.array([1.0, -2.0, 3.0, -4.0, 5.0])
    res = np.array([1.5, -2.0, 2.0, -4.0, 5.0], dtype=np.float32)
    assert oracle.l1_rel_error(ref, res) == pytest.approx(1.5 / 15.0)
    assert oracle.l1_rel_error(np.array([2.0]), np.array([1.0])) == pytest.approx(0.5)


def test_cpu_baseline_matches_exact(oracle):
    rng = np.random.default_rng(5)
    m, n = 64, 4099
    for st in (np.float64, np.float32, np.float16):
        A = oracle.convert(rng.uniform(-1, 1, m * n), st)
        x = oracle.convert(rng.uniform(-1, 1, n), st)
        y = oracle.convert(rng.uniform(-1, 1, m), st)
        exact = oracle.exact_gemv(A, m, n, n, x, 1.0, 1.0, y)
        got = oracle.cpu_gemv(np.float64, A, m, n, n, x, 1.0, 1.0, y.copy())
        tol = {np.float64: 1e-14, np.float32: 1e-7, np.float16: 6e-4}[st]
        assert oracle.l1_rel_error(exact, got) < tol
        d = oracle.cpu_dot(np.float64, x, A[:n].copy(), np.float64)
        assert abs(d - oracle.exact_dot(x, A[:n].copy())) < 1e-11


# --- golden outputs of the reference's own CUDA kernels (generated on a B200) ---
def _reference_golden():
    path = GOLDEN / "reference_kernels_b200.npz"
    if not path.exists():
        pytest.skip("tests/golden/reference_kernels_b200.npz not generated yet")
    return np.load(path)


_DT = {"f64": np.float64, "f32": np.float32}


def test_oracle_matches_reference_cuda_kernels_gemv_bitwise():
    """The order-faithful GEMV restatement equals what the reference's kernels
    produced on the B200, bit for bit (tests/golden/make_reference_golden.py)."""
    from oracle_binding import Oracle
    orc = Oracle()
    g = _reference_golden()
    keys = [k for k in g.files if k.startswith("gemv_")]
    assert keys
    for key in keys:
        _, m, n, lda, ar, st, plain = key.split("_")
        m, n, lda = int(m), int(n), int(lda)
        A = orc.convert(orc.uniform(m * lda, seed=42), _DT[st])
        x = orc.convert(orc.uniform(n, seed=42, first_draw=m * lda), _DT[st])
        y = orc.convert(orc.uniform(m, seed=42, first_draw=m * lda + n), _DT[st])
        want = orc.ref_gemv(_DT[ar], A, m, n, lda, x, 1.0, 1.0, y)
        assert np.array_equal(want.view(np.uint8), g[key].view(np.uint8)), key


def test_oracle_matches_reference_cuda_kernels_dot_and_trsv():
    from oracle_binding import Oracle
    orc = Oracle()
    g = _reference_golden()
    blocks = int(g["sm_count"][0]) * 32
    for key in [k for k in g.files if k.startswith("dot_")]:
        _, n, ar, st, res, plain = key.split("_")
        n = int(n)
        x = orc.convert(orc.uniform(n, seed=42), _DT[st])
        y = orc.convert(orc.uniform(n, seed=42, first_draw=n), _DT[st])
        want, partials = orc.ref_dot(_DT[ar], x, y, _DT[res], blocks=blocks)
        # the GPU combines the block partials with atomics in arbitrary order
        eps = 2.3e-16 if ar == "f64" else 1.2e-7
        slack = 4 * eps * np.abs(partials.astype(np.float64)).sum() + \
            (6e-8 * abs(float(want)) if res == "f32" else 0.0)
        assert abs(float(g[key][0]) - float(want)) <= slack, key
    for key in [k for k in g.files if k.startswith("trsv_")]:
        _, n, upper, unit, ar, st, plain = key.split("_")
        n, upper, unit = int(n), int(upper), int(unit)
        base = orc.uniform(n * n, seed=7).reshape(n, n) * 0.02
        T = base.copy()
        np.fill_diagonal(T, 1.0 + 0.5 * orc.uniform(n, seed=8))
        A = orc.convert(T.reshape(-1), _DT[st])
        b = orc.convert(orc.uniform(n, seed=9), _DT[st])
        want = orc.ref_trsv(_DT[ar], A, n, n, b, upper, unit)
        got = g[key]
        if np.array_equal(want.view(np.uint8), got.view(np.uint8)):
            continue
        exact = orc.exact_trsv(A, n, n, b, upper, unit)
        e_got, e_want = orc.l1_rel_error(exact, got), orc.l1_rel_error(exact, want)
        assert e_got <= 1.5 * e_want + 1e-15 and e_want <= 1.5 * e_got + 1e-15, key
