"""GPU parity tests: the sm_100a kernels, called through the C ABI, against the
CPU oracle on identical seeded inputs.

Bars (north_star): storage<->arithmetic conversion and the input stream are
BIT-EXACT; GEMV / DOT / TRSV agree with the accuracy oracle within the
relative error stated per (arithmetic, storage) pair below, and are no worse
than the reference's own kernels (order-faithful restatement).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
ST = [torch.float64, torch.float32, torch.float16]
AR = [torch.float64, torch.float32]
NP = {torch.float64: np.float64, torch.float32: np.float32, torch.float16: np.float16}

# L1-relative error bar of GEMV (sum|ref-res|/sum|ref|, ref = exact result of
# the STORED operands) per (arithmetic, storage): the floor is the rounding of
# the result to storage (u_st/2 on average), plus accumulation error in Ar.
GEMV_TOL = {
    (torch.float64, torch.float64): 2e-15,
    (torch.float64, torch.float32): 6e-8,     # reference plots 3.8e-8..4.2e-8
    (torch.float64, torch.float16): 5e-4,     # 0.67 * 2^-11 predicted floor
    (torch.float32, torch.float64): 4e-7,     # inputs cast to fp32, fp32 sums
    (torch.float32, torch.float32): 4e-7,     # reference plots 0.95e-7..1.5e-7
    (torch.float32, torch.float16): 5e-4,
}
# |res-ref|/|ref| bar for DOT with the result kept in the arithmetic type,
# scaled by sqrt(n) conditioning of a random uniform(-1,1) dot product
DOT_TOL = {torch.float64: 5e-14, torch.float32: 2e-5}


def dev(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def host(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


def stored(oracle, count, st, seed=42, first=0):
    return oracle.convert(oracle.uniform(count, seed=seed, first_draw=first), NP[st])


# ---------------------------------------------------------------------------
# conversion + generation: bit-exact
# ---------------------------------------------------------------------------
def special_values():
    h = np.arange(0, 0x7c00, 37, dtype=np.uint16).view(np.float16).astype(np.float64)
    mid = (h[:-1] + h[1:]) / 2
    f = np.array([1 + 2.0 ** -11 + 2.0 ** -30, 65504.0, 65519.99, 65520.0, 1e6,
                  2.0 ** -24, 2.0 ** -25, 0.0, -0.0, np.inf, -np.inf, 1e-310,
                  1.0000000596046448, 1 + 2.0 ** -24, 1 + 2.0 ** -24 + 2.0 ** -50,
                  3.4028235677973366e38, 1e39, 1e-46, 1.4e-45])
    return np.concatenate([h, -h, mid, -mid, np.nextafter(mid, np.inf), f, -f])


@pytest.mark.parametrize("src", ST)
@pytest.mark.parametrize("dst", ST)
def test_convert_bit_exact(oracle, handle, src, dst):
    rng = np.random.default_rng(7)
    x64 = np.concatenate([special_values(), rng.uniform(-1, 1, 100_003),
                          rng.uniform(-1e-5, 1e-5, 5000)])
    x = oracle.convert(x64, NP[src])
    want = oracle.convert(x, NP[dst])
    out = torch.empty(x.size, dtype=dst, device=DEV)
    handle.convert(1, x.size, dev(x), x.size, out, x.size)
    got = host(out)
    view = {2: np.uint16, 4: np.uint32, 8: np.uint64}[got.itemsize]
    assert np.array_equal(got.view(view), want.view(view))


@pytest.mark.parametrize("dst", ST)
def test_convert_strided_rows_and_unaligned(oracle, handle, dst):
    rows, cols, ld_in, ld_out = 37, 101, 131, 117
    x = oracle.uniform(rows * ld_in + 1, seed=3)
    src = dev(x)[1:]  # 8-byte but not 32-byte aligned
    out = torch.zeros(rows * ld_out, dtype=dst, device=DEV)
    handle.convert(rows, cols, src, ld_in, out, ld_out)
    got = host(out).reshape(rows, ld_out)
    want = oracle.convert(x[1:].reshape(rows, ld_in)[:, :cols].copy(), NP[dst])
    assert np.array_equal(got[:, :cols].view(np.uint8), want.view(np.uint8))
    assert not got[:, cols:].any()  # padding untouched


@pytest.mark.parametrize("dst", ST)
@pytest.mark.parametrize("first", [0, 12345, 24500 * 24500, 2 ** 34 + 7])
def test_fill_uniform_bit_exact(oracle, handle, dst, first):
    rows, cols, ld = 13, 1001, 1024
    out = torch.zeros(rows * ld, dtype=dst, device=DEV)
    handle.fill_uniform(rows, cols, out, ld, seed=42, first_draw=first)
    got = host(out).reshape(rows, ld)
    want = oracle.convert(oracle.uniform(rows * cols, seed=42, first_draw=first),
                          NP[dst]).reshape(rows, cols)
    assert np.array_equal(got[:, :cols].view(np.uint8), want.view(np.uint8))
    assert not got[:, cols:].any()


def test_fill_uniform_matches_libstdcxx_engine(oracle, handle):
    n = 50_000
    out = torch.empty(n, dtype=torch.float64, device=DEV)
    for seed in (42, 1, 0, 2 ** 31 - 1, 123456789):
        handle.fill_uniform(1, n, out, n, seed=seed)
        assert np.array_equal(host(out), oracle.uniform(n, seed=seed, closed_form=False))


# ---------------------------------------------------------------------------
# GEMV
# ---------------------------------------------------------------------------
def run_gemv(handle, ar, A, m, n, lda, x, alpha, beta, y, incx=1, incy=1):
    yd = dev(y)
    handle.gemv(ar, m, n, alpha, dev(A), lda, dev(x), incx, beta, yd, incy)
    torch.cuda.synchronize()
    return host(yd)


@pytest.mark.parametrize("dst", ST)
def test_fill_fast_equals_generic(ab, handle, dst):
    """The streaming generator (constant-multiplier jumps, reciprocal-multiply
    division) against the per-row kernel (__ddiv_rn) on 2^27 draws, several
    starting points incl. beyond 2^32, bit for bit -- on the device."""
    count = 2 ** 27
    try:
        for first in (0, 999_983, 2 ** 32 + 12345, 2 ** 40 + 1):
            a = torch.empty(count, dtype=dst, device=DEV)
            b = torch.empty(count, dtype=dst, device=DEV)
            ab.tune("fill_generic", 0)
            handle.fill_uniform(1, count, a, count, seed=42, first_draw=first)
            ab.tune("fill_generic", 1)
            handle.fill_uniform(1, count, b, count, seed=42, first_draw=first)
            assert torch.equal(a.view(torch.uint8), b.view(torch.uint8)), first
            # a ragged count and a matrix view take the same path
            ab.tune("fill_generic", 0)
            c = torch.empty(1000 * 777, dtype=dst, device=DEV)
            handle.fill_uniform(1000, 777, c, 777, seed=42, first_draw=first)
            assert torch.equal(c.view(torch.uint8), b[:1000 * 777].view(torch.uint8))
            del a, b, c
    finally:
        ab.tune("fill_generic", 0)


@pytest.mark.parametrize("ar", AR)
@pytest.mark.parametrize("st", ST)
@pytest.mark.parametrize("m,n,lda", [(1, 1, 1), (7, 3, 8), (100, 100, 100),
                                     (515, 1030, 1032), (33, 4100, 4104),
                                     (1000, 1000, 24500 if False else 1000),
                                     (300, 2049, 2056), (2500, 517, 520),
                                     (20000, 64, 64)])
def test_gemv_parity(oracle, handle, ar, st, m, n, lda):
    A = stored(oracle, m * lda, st, first=0)
    x = stored(oracle, n, st, first=m * lda)
    y = stored(oracle, m, st, first=m * lda + n)
    got = run_gemv(handle, ar, A, m, n, lda, x, 1.0, 1.0, y)
    exact = oracle.exact_gemv(A, m, n, lda, x, 1.0, 1.0, y)
    err = oracle.l1_rel_error(exact, got)
    tol = GEMV_TOL[(ar, st)] * max(1.0, np.sqrt(n / 16384))
    assert err <= tol, (err, tol)
    # no worse than the reference kernel (same inputs, its summation order)
    ref = oracle.ref_gemv(NP[ar], A, m, n, lda, x, 1.0, 1.0, y)
    ref_err = oracle.l1_rel_error(exact, ref)
    assert err <= 1.5 * ref_err + 1e-16, (err, ref_err)


@pytest.mark.parametrize("ar", AR)
@pytest.mark.parametrize("st", ST)
@pytest.mark.parametrize("lda_pad,base_off", [(0, 0), (4, 0), (2, 0), (1, 0), (6, 2), (0, 4), (0, 1),
                                              (8, 8)])
def test_gemv_alignment_classes(oracle, handle, ar, st, lda_pad, base_off):
    """Rows that are only 8-, 4- (or element-) aligned take the cp.async path
    with smaller pieces (or the scalar kernel): same answers as the 16-byte
    aligned layout of the same numbers -- bit-identical for fp64 arithmetic
    (same chunk ownership and accumulators), within tolerance for fp32.  Sizes
    include several full chunks per warp, a ragged tail and the stride of the
    reference driver's default sweep (24500)."""
    for m, n in ((37, 12401), (5, 24500)):
        lda = n + lda_pad
        lda_ref = (n + 7) // 8 * 8
        vals = stored(oracle, m * n, st, first=7)
        x = stored(oracle, n + 8, st, first=10 ** 7)
        y = stored(oracle, m, st, first=2 * 10 ** 7)
        A_ref = np.zeros(m * lda_ref, dtype=NP[st])
        A_ref.reshape(m, lda_ref)[:, :n] = vals.reshape(m, n)
        A = np.zeros(base_off + m * lda, dtype=NP[st])
        A[base_off:].reshape(m, lda)[:, :n] = vals.reshape(m, n)
        want = run_gemv(handle, ar, A_ref, m, n, lda_ref, x[:n].copy(), 1.0, 1.0, y)
        yd = dev(y)
        xo = base_off % 8
        handle.gemv(ar, m, n, 1.0, dev(A)[base_off:], lda, dev(x)[xo:], 1, 1.0, yd, 1)
        got = host(yd)
        x_used = x[xo:xo + n].copy()
        if xo:
            want = run_gemv(handle, ar, A_ref, m, n, lda_ref, x_used, 1.0, 1.0, y)
        if ar == torch.float64:
            assert np.array_equal(got, want), (m, n, lda, base_off)
        exact = oracle.exact_gemv(A_ref, m, n, lda_ref, x_used, 1.0, 1.0, y)
        err = oracle.l1_rel_error(exact, got)
        assert err <= GEMV_TOL[(ar, st)] * max(1.0, np.sqrt(n / 16384)), (err, lda, base_off)


@pytest.mark.parametrize("st", ST)
def test_gemv_alpha_beta_strides_alignment(oracle, handle, st):
    m, n, lda, incx, incy = 257, 1031, 1040, 3, 2
    A = stored(oracle, m * lda + 3, st)
    x = stored(oracle, n * incx, st, first=10 ** 6)
    y = stored(oracle, m * incy, st, first=2 * 10 ** 6)
    for off in (0, 1, 3):                       # misaligned matrix base
        for alpha, beta in ((0.5, -2.0), (1.0, 0.0), (-1.25, 1.0)):
            Ao = A[off:off + m * lda].copy()
            yd = dev(y)
            Ad = dev(A)[off:]
            handle.gemv(torch.float64, m, n, alpha, Ad, lda, dev(x), incx, beta, yd, incy)
            got = host(yd)
            exact = oracle.exact_gemv(Ao, m, n, lda, x, alpha, beta, y, incx, incy)
            err = oracle.l1_rel_error(exact, got[::incy].copy())
            assert err <= GEMV_TOL[(torch.float64, st)], (off, alpha, beta, err)
            # elements between the strided outputs are untouched
            assert np.array_equal(got[1::incy], y[1::incy])


def test_gemv_beta_zero_ignores_output_nans(oracle, handle):
    m, n = 129, 513
    A = stored(oracle, m * n, torch.float32)
    x = stored(oracle, n, torch.float32, first=m * n)
    y = np.full(m, np.nan, dtype=np.float32)
    got = run_gemv(handle, torch.float64, A, m, n, n, x, 1.0, 0.0, y)
    assert np.isfinite(got).all()


def test_gemv_empty_and_errors(ab, handle):
    y = torch.ones(4, dtype=torch.float32, device=DEV)
    a = torch.ones(16, dtype=torch.float32, device=DEV)
    handle.gemv(torch.float64, 0, 4, 1.0, a, 4, y, 1, 1.0, y, 1)  # m == 0: no-op
    handle.gemv(torch.float64, 4, 0, 1.0, a, 4, y, 1, 2.0, y, 1)  # n == 0: y *= beta
    torch.cuda.synchronize()
    assert host(y).tolist() == [2.0] * 4
    with pytest.raises(ab.AccblasError):
        handle.gemv(torch.float64, 4, 4, 1.0, a, 3, y, 1, 1.0, y, 1)  # lda < n
    with pytest.raises(ab.AccblasError):
        handle.gemv(torch.float16, 4, 4, 1.0, a, 4, y, 1, 1.0, y, 1)  # fp16 arithmetic


# ---------------------------------------------------------------------------
# DOT
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("ar", AR)
@pytest.mark.parametrize("st", ST)
@pytest.mark.parametrize("n", [0, 1, 5, 1000, 4097, 2 ** 20, 3_000_001])
def test_dot_parity_and_determinism(oracle, handle, ar, st, n):
    x = stored(oracle, max(n, 1), st)[:n]
    y = stored(oracle, max(n, 1), st, first=max(n, 1))[:n]
    xd, yd = dev(x) if n else torch.empty(0, dtype=st, device=DEV), \
        dev(y) if n else torch.empty(0, dtype=st, device=DEV)
    res = torch.full((1,), -999.0, dtype=ar, device=DEV)  # the reference's sentinel
    handle.dot(ar, n, xd, 1, yd, 1, res)
    first = res.clone()
    exact = oracle.exact_dot(x, y) if n else 0.0
    got = float(first.item())
    scale = max(np.sqrt(n / 3.0) / 3.0, abs(exact), 1e-30)
    assert abs(got - exact) <= DOT_TOL[ar] * scale * max(1.0, np.log2(max(n, 2)) / 20), \
        (got, exact)
    for _ in range(3):                              # bit-reproducible run to run
        res.fill_(-999.0)
        handle.dot(ar, n, xd, 1, yd, 1, res)
        assert torch.equal(res, first)
    if n:
        ref, _ = oracle.ref_dot(NP[ar], x, y, NP[ar], blocks=148 * 32)
        assert abs(got - exact) <= 2.0 * abs(float(ref) - exact) + DOT_TOL[ar] * scale * 0.05


@pytest.mark.parametrize("ar", AR)
@pytest.mark.parametrize("st", ST)
@pytest.mark.parametrize("off", [1, 3, 5])
def test_dot_common_misalignment_is_peeled(oracle, handle, ar, st, off):
    """x[off:] . y[off:] (both operands misaligned by the same amount) runs the
    streaming kernel on the aligned body plus `head` scalar products; different
    misalignments fall back to the strided kernel.  Both must agree with the
    oracle."""
    n = 300_007
    x = stored(oracle, n + 16, st, seed=31)
    y = stored(oracle, n + 16, st, seed=32)
    for ox, oy in ((off, off), (off, off + 1), (0, off)):
        xs, ys = x[ox:ox + n].copy(), y[oy:oy + n].copy()
        res = torch.zeros(1, dtype=ar, device=DEV)
        handle.dot(ar, n, dev(x)[ox:], 1, dev(y)[oy:], 1, res)
        exact = oracle.exact_dot(xs, ys)
        scale = float(np.abs(xs.astype(np.float64) * ys.astype(np.float64)).sum())
        assert abs(float(res.item()) - exact) <= DOT_TOL[ar] * scale, (ox, oy)
    for tiny in (0, 1, 2, 7):       # shorter than the head
        res = torch.full((1,), -1.0, dtype=ar, device=DEV)
        handle.dot(ar, tiny, dev(x)[off:], 1, dev(y)[off:], 1, res)
        want = float(np.dot(x[off:off + tiny].astype(np.float64), y[off:off + tiny].astype(np.float64)))
        assert abs(float(res.item()) - want) <= 1e-6 * max(1.0, abs(want))


@pytest.mark.parametrize("ar", AR)
@pytest.mark.parametrize("st", ST)
def test_dot_differently_misaligned_operands_keep_the_vector_path(oracle, handle, ar, st):
    """x 16-byte aligned, y shifted by 1..7 elements: y's 16 bytes are fetched
    in 8- / 4- / 2-byte pieces into the registers of the aligned kernel, so the
    result must be BIT-identical to the same data with both operands aligned
    (same launch shape, same summation order)."""
    n = 1_000_003
    x = stored(oracle, n, st, seed=41)
    y = stored(oracle, n + 16, st, seed=43)
    xd = dev(x)
    per16 = 16 // np.dtype(NP[st]).itemsize
    for shift in sorted({1, 2, 3, per16 // 2, per16 - 1} - {0}):
        if shift >= per16:
            continue
        ys = y[shift:shift + n].copy()
        aligned = torch.zeros(1, dtype=ar, device=DEV)
        handle.dot(ar, n, xd, 1, dev(ys), 1, aligned)            # fresh, aligned copy
        shifted = torch.zeros(1, dtype=ar, device=DEV)
        ybuf = dev(y)
        assert ybuf.data_ptr() % 16 == 0
        handle.dot(ar, n, xd, 1, ybuf[shift:], 1, shifted)
        swapped = torch.zeros(1, dtype=ar, device=DEV)
        handle.dot(ar, n, ybuf[shift:], 1, xd, 1, swapped)        # the misaligned one first
        exact = oracle.exact_dot(x, ys)
        scale = float(np.abs(x.astype(np.float64) * ys.astype(np.float64)).sum())
        assert torch.equal(aligned, shifted), (shift, aligned.item(), shifted.item())
        assert abs(float(swapped.item()) - exact) <= DOT_TOL[ar] * scale
        assert abs(float(shifted.item()) - exact) <= DOT_TOL[ar] * scale


def test_dot_launch_shapes_and_integer_widening(oracle, ab, handle):
    """Every CTA size / unroll the tuner can select, and the integer-pipe
    widening of Acc<fp64,fp32> (bit-identical to the conversion-pipe path of
    the same shape on ordinary data; Inf/NaN inputs take the exact fallback)."""
    n = 3_000_017
    try:
        for ar, st in ((torch.float64, torch.float32), (torch.float32, torch.float16),
                       (torch.float64, torch.float64)):
            x = stored(oracle, n, st, seed=51)
            y = stored(oracle, n, st, seed=52)
            exact = oracle.exact_dot(x, y)
            scale = float(np.abs(x.astype(np.float64) * y.astype(np.float64)).sum())
            for block in (256, 512, 1024):
                for unroll in (2, 4):
                    ab.tune("dot_block", block)
                    ab.tune("dot_unroll", unroll)
                    outs = []
                    for mix in ((0, 1) if (ar, st) == (torch.float64, torch.float32) else (0,)):
                        ab.tune("dot_intmix", mix)
                        res = torch.zeros(1, dtype=ar, device=DEV)
                        handle.dot(ar, n, dev(x), 1, dev(y), 1, res)
                        outs.append(res.clone())
                        assert abs(float(res.item()) - exact) <= DOT_TOL[ar] * scale, (block, unroll, mix)
                    if len(outs) == 2:
                        assert torch.equal(outs[0], outs[1]), (block, unroll)
        # non-finite inputs through the integer-widening path
        ab.tune("dot_block", 0)
        ab.tune("dot_unroll", 0)
        x = stored(oracle, n, torch.float32, seed=53)
        y = stored(oracle, n, torch.float32, seed=54)
        for bad, where in ((np.inf, 12345), (-np.inf, n - 7), (np.nan, 2_000_000)):
            xb = x.copy()
            xb[where] = bad
            got = []
            for mix in (0, 1):
                ab.tune("dot_intmix", mix)
                res = torch.zeros(1, dtype=torch.float64, device=DEV)
                handle.dot(torch.float64, n, dev(xb), 1, dev(y), 1, res)
                got.append(float(res.item()))
            assert (np.isnan(got[0]) and np.isnan(got[1])) or got[0] == got[1], (bad, got)
            assert not np.isfinite(got[1])
    finally:
        ab.tune("dot_block", 0)
        ab.tune("dot_unroll", 0)
        ab.tune("dot_intmix", 0)


@pytest.mark.parametrize("st", ST)
def test_dot_result_types_strides_alignment(oracle, handle, st):
    n, incx, incy = 70_001, 2, 3
    x = stored(oracle, n * incx + 1, st)
    y = stored(oracle, n * incy + 1, st, first=10 ** 6)
    exact = oracle.exact_dot(x[1:], y[1:], n, incx, incy)
    for res_t in ST:
        res = torch.zeros(1, dtype=res_t, device=DEV)
        handle.dot(torch.float64, n, dev(x)[1:], incx, dev(y)[1:], incy, res)
        want = oracle.convert(np.array([exact]), NP[res_t])[0]
        got = host(res)[0]
        # the fp64 accumulation is far more accurate than one ulp of fp32/fp16,
        # so the rounded result must be the correctly rounded exact value
        # (or its neighbour when the exact value sits on a rounding boundary)
        assert got == want or abs(float(got) - exact) <= 1.5e-12 * abs(exact) + \
            abs(float(np.spacing(want))), (got, want)
    # unaligned but contiguous
    res = torch.zeros(1, dtype=torch.float64, device=DEV)
    handle.dot(torch.float64, n, dev(x)[1:], 1, dev(y)[1:], 1, res)
    assert abs(res.item() - oracle.exact_dot(x[1:n + 1].copy(), y[1:n + 1].copy())) < 1e-10


# ---------------------------------------------------------------------------
# TRSV
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("ar", AR)
@pytest.mark.parametrize("st", ST)
def test_dot_dynamic_pool_is_deterministic(oracle, ab, handle, ar, st):
    """The last share of the tiles is handed out by an atomic counter; every
    chunk has its own partial, folded in chunk order, so which CTA took which
    chunk must not show in the bits -- and the pool must not lose or double
    anything (ragged n, more chunks than CTAs)."""
    n = 2 ** 25 + 12345
    x = stored(oracle, n, st, seed=42, first=0)
    y = stored(oracle, n, st, seed=42, first=n)
    xd, yd = dev(x), dev(y)
    exact = oracle.exact_dot(x, y, n)
    res = torch.zeros(1, dtype=ar, device=DEV)
    try:
        ab.tune("dot_ctas_per_sm", 1)   # small grid: the pool is active from 32 tiles per CTA on
        seen = {}
        for pct, ct in ((12, 4), (12, 4), (12, 4), (50, 1), (50, 1), (0, 4)):
            ab.tune("dot_pool_pct", pct)
            ab.tune("dot_chunk_tiles", ct)
            for _ in range(3):
                handle.dot(ar, n, xd, 1, yd, 1, res)
                torch.cuda.synchronize()
                seen.setdefault((pct, ct), set()).add(res.cpu().numpy().tobytes())
        for key, bits in seen.items():
            assert len(bits) == 1, key
        tol = {torch.float64: 1e-13, torch.float32: 2e-5}[ar]
        scale = float(np.abs(x.astype(np.float64) * y.astype(np.float64)).sum())
        for key, bits in seen.items():
            got = float(np.frombuffer(next(iter(bits)), dtype=NP[ar])[0])
            assert abs(got - exact) <= tol * scale, (key, got, exact)
    finally:
        ab.tune("dot_ctas_per_sm", 0)
        ab.tune("dot_pool_pct", 12)
        ab.tune("dot_chunk_tiles", 4)


def lu_fixture(n, seed, lda=None):
    """Row-major matrix whose triangles are well conditioned the way the
    reference's fixture is (partially pivoted LU of uniform(-1,1) data,
    cuda/trsv_memory.cuh:131-168): strict lower = L (|l_ij| <= 1), upper incl.
    diagonal = U."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    M = (torch.rand(n, n, generator=g, dtype=torch.float64) * 2 - 1).to(DEV)
    LU, _ = torch.linalg.lu_factor(M)
    lda = lda or n
    out = torch.zeros(n, lda, dtype=torch.float64, device=DEV)
    out[:, :n] = LU
    return host(out).reshape(-1)


TRSV_TOL = {
    (torch.float64, torch.float64): 1e-11,
    (torch.float64, torch.float32): 2e-6,
    (torch.float64, torch.float16): 2e-2,
    (torch.float32, torch.float64): 5e-4,
    (torch.float32, torch.float32): 5e-4,
    (torch.float32, torch.float16): 2e-2,
}


@pytest.mark.parametrize("ar", AR)
@pytest.mark.parametrize("st", ST)
@pytest.mark.parametrize("n", [1, 31, 33, 128, 129, 300, 1000])
@pytest.mark.parametrize("upper,unit,transpose", [(False, True, False),
                                                  (True, True, True),
                                                  (True, False, False),
                                                  (False, False, True)])
def test_trsv_parity(oracle, ab, handle, ar, st, n, upper, unit, transpose):
    """unit triangles hold L (or L^T), non-unit ones U (or U^T)."""
    lda = n + (8 if n > 1 else 0)
    LU = lu_fixture(n, seed=100 + n, lda=lda).reshape(n, lda)
    if transpose:
        T = np.zeros_like(LU)
        T[:, :n] = LU[:, :n].T
        LU = T
    A = oracle.convert(LU.reshape(-1), NP[st])
    b = stored(oracle, n, st, seed=9)
    xd = dev(b)
    handle.trsv(ar, ab.UPPER if upper else ab.LOWER, ab.UNIT if unit else ab.NON_UNIT,
                n, dev(A), lda, xd, 1)
    torch.cuda.synchronize()
    got = host(xd)
    exact = oracle.exact_trsv(A, n, lda, b, upper, unit)
    ref = oracle.ref_trsv(NP[ar], A, n, lda, b, upper, unit)
    err = oracle.l1_rel_error(exact, got)
    ref_err = oracle.l1_rel_error(exact, ref)
    if not np.isfinite(ref_err):
        pytest.skip("fixture not representable in this storage type")
    # conditioning of U grows with n: bar relative to the reference kernel,
    # absolute bar for the well-conditioned unit case
    assert err <= 3.0 * ref_err + 1e-15, (err, ref_err)
    if unit:
        assert err <= TRSV_TOL[(ar, st)] * max(1.0, n / 300), (err, ref_err)


@pytest.mark.parametrize("st", ST)
@pytest.mark.parametrize("pad", [0, 1, 2, 4, 8])
def test_trsv_row_alignment_classes(oracle, ab, handle, st, pad):
    """Rows aligned to 16 bytes (128-bit loads), to 8 bytes (64-bit loads) or
    less (scalar loads) must give bit-identical solutions of the same system."""
    n = 700
    LU = lu_fixture(n, seed=3, lda=n).reshape(n, n)
    b = stored(oracle, n, st, seed=14)
    outs = []
    for lda in (n + 8 - n % 8 + 8, n + pad):
        M = np.zeros((n, lda))
        M[:, :n] = LU
        A = oracle.convert(M.reshape(-1), NP[st])
        xd = dev(b)
        handle.trsv(torch.float64, ab.LOWER, ab.UNIT, n, dev(A), lda, xd, 1)
        torch.cuda.synchronize()
        outs.append(host(xd))
    assert np.array_equal(outs[0], outs[1], equal_nan=True), (st, pad)


@pytest.mark.parametrize("variant,whole,group", [(0, 1, 0), (0, 1, 4096), (1, 0, 1024), (1, 1, 0),
                                                 (1, 0, 0), (1, 1, 4096)])
def test_trsv_wait_modes_give_identical_results(oracle, ab, handle, variant, whole, group):
    """Both kernels (0 = clusters with the hand-off through distributed shared
    memory, 1 = one CTA per block row through L2): how a caught-up CTA waits
    for x and whether tiles are requested into L2 ahead of time must not change
    a bit of the result (many block rows: the chain, the x staging and every
    wait mode are exercised); and each kernel stays within the error bar."""
    n, lda = 3000, 3008
    LU = lu_fixture(n, seed=31, lda=lda)
    try:
        ab.tune("trsv_variant", variant)
        for ar, st, upper, unit in ((torch.float64, torch.float32, False, True),
                                    (torch.float32, torch.float16, False, True),
                                    (torch.float64, torch.float64, True, False)):
            A = oracle.convert(LU, NP[st])
            b = stored(oracle, n, st, seed=13)
            outs = []
            for w, g in ((1, 1024), (whole, group)):
                ab.tune("trsv_whole_block_spin", w)
                ab.tune("trsv_l2_ahead", g)
                xd = dev(b)
                handle.trsv(ar, ab.UPPER if upper else ab.LOWER, ab.UNIT if unit else ab.NON_UNIT,
                            n, dev(A), lda, xd, 1)
                torch.cuda.synchronize()
                outs.append(host(xd))
            assert np.array_equal(outs[0], outs[1], equal_nan=True), (ar, st, variant, whole, group)
            if unit:
                exact = oracle.exact_trsv(A, n, lda, b, upper, unit)
                assert oracle.l1_rel_error(exact, outs[0]) <= TRSV_TOL[(ar, st)] * n / 300
    finally:
        ab.tune("trsv_variant", -1)
        ab.tune("trsv_whole_block_spin", 1)
        ab.tune("trsv_l2_ahead", -1)


@pytest.mark.parametrize("ar,st", [(torch.float64, torch.float32), (torch.float32, torch.float32),
                                   (torch.float64, torch.float16), (torch.float64, torch.float64)])
@pytest.mark.parametrize("n", [127, 128, 1025, 2500])
def test_trsv_both_kernels_meet_the_reference_bar(oracle, ab, handle, ar, st, n):
    """The single-CTA-per-block-row kernel stays in the library as the
    fallback for devices where no cluster fits: same bar for both."""
    LU = lu_fixture(n, seed=5)
    A = oracle.convert(LU, NP[st])
    b = stored(oracle, n, st, seed=6)
    try:
        for upper, unit in ((False, True), (True, False)):
            exact = oracle.exact_trsv(A, n, n, b, upper, unit)
            ref = oracle.ref_trsv(NP[ar], A, n, n, b, upper, unit)
            ref_err = oracle.l1_rel_error(exact, ref)
            for variant in (0, 1):
                ab.tune("trsv_variant", variant)
                xd = dev(b)
                handle.trsv(ar, ab.UPPER if upper else ab.LOWER, ab.UNIT if unit else ab.NON_UNIT,
                            n, dev(A), n, xd, 1)
                torch.cuda.synchronize()
                err = oracle.l1_rel_error(exact, host(xd))
                assert err <= 3.0 * ref_err + 1e-15, (variant, upper, unit, err, ref_err)
    finally:
        ab.tune("trsv_variant", -1)


def _substitute_lower_unit(A, b, st_np):
    """Row-by-row forward substitution in fp64 with every solved entry rounded
    through the storage type (what the accessor write/re-read does)."""
    n = b.shape[0]
    x = np.zeros(n, dtype=np.float64)
    with np.errstate(all="ignore"):
        for r in range(n):
            x[r] = np.float64(st_np(b[r] - np.dot(A[r, :r].astype(np.float64), x[:r])))
    return x


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("ar", AR)
@pytest.mark.parametrize("st", ST)
def test_trsv_nonfinite_matrix_entries_propagate(ab, handle, ar, st, variant):
    """Inf / NaN in off-diagonal tiles poison the solution from their row on and
    nothing in front of their 32-row sub-block."""
    n = 300
    rng = np.random.default_rng(17)
    A = np.tril(rng.uniform(-1.0, 1.0, (n, n)) / 64.0, -1).astype(NP[st])
    A[200, 10] = np.inf
    A[290, 150] = np.nan
    A[260, 3] = -np.inf
    b = rng.uniform(-1.0, 1.0, n).astype(NP[st])
    want = _substitute_lower_unit(A, b, NP[st])
    try:
        ab.tune("trsv_variant", variant)
        xd = dev(b)
        handle.trsv(ar, ab.LOWER, ab.UNIT, n, dev(A.reshape(-1)), n, xd, 1)
        torch.cuda.synchronize()
        got = host(xd).astype(np.float64)
    finally:
        ab.tune("trsv_variant", -1)
    # The diagonal tile is solved in product form (x_g = Inv_g rhs_g - ...): a
    # non-finite rhs entry meets the structural zeros of Inv_g, so the rows of
    # ITS 32-row sub-block in front of it may come out NaN where substitution
    # keeps them finite (documented in DESIGN.md).  Everything in front of that
    # sub-block is exact, everything from the poisoned row on is non-finite.
    fin = np.isfinite(want)
    assert fin[:200].all() and not fin[200:].any()
    assert np.isfinite(got[:192]).all()
    assert not np.isfinite(got[200:]).any()
    assert np.abs(got[:192] - want[:192]).sum() <= TRSV_TOL[(ar, st)] * np.abs(want[:192]).sum()


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("ar", AR)
@pytest.mark.parametrize("st", ST)
def test_trsv_subnormal_entries_times_huge_solution(ab, handle, ar, st, variant):
    """Off-diagonal tiles made of SUBNORMAL storage values, solution entries
    near the top of the storage range: every row of the second and third block
    row is a sum of (subnormal * huge) products, so a conversion that flushed
    subnormals shows up entry by entry."""
    if ar == torch.float32 and st == torch.float64:
        pytest.skip("fp64 subnormals are zero in fp32 arithmetic")
    n = 300
    rng = np.random.default_rng(23)
    info = np.finfo(NP[st])
    tiny = float(info.smallest_subnormal)
    mant = 2 ** (info.nmant - 1)
    huge = {torch.float16: 3.0e4, torch.float32: 1.0e38, torch.float64: 1.0e300}[st]
    if ar == torch.float32:
        huge = min(huge, 1.0e38)
    A = np.zeros((n, n), dtype=np.float64)
    A[128:, :128] = rng.integers(-mant, mant, (n - 128, 128)) * tiny
    A = A.astype(NP[st])
    assert (np.abs(A[128:, :128]) < float(info.tiny)).all()
    b = np.zeros(n, dtype=np.float64)
    b[:128] = rng.uniform(0.5, 1.0, 128) * huge * rng.choice([-1.0, 1.0], 128)
    b = b.astype(NP[st])
    want = _substitute_lower_unit(A, b, NP[st])
    try:
        ab.tune("trsv_variant", variant)
        xd = dev(b)
        handle.trsv(ar, ab.LOWER, ab.UNIT, n, dev(A.reshape(-1)), n, xd, 1)
        torch.cuda.synchronize()
        got = host(xd).astype(np.float64)
    finally:
        ab.tune("trsv_variant", -1)
    assert np.isfinite(got).all()
    assert np.array_equal(got[:128], want[:128])
    tail = want[128:]
    assert (tail != 0).sum() > 100
    # bound per row: storage rounding of the result + accumulation of 128 terms
    rel = max(2.0 ** -(info.nmant - 1), 2.0 ** -44) if ar == torch.float64 else 1e-4
    terms = np.abs(A[128:, :128].astype(np.float64)) @ np.abs(want[:128])
    floor = float(info.smallest_subnormal)
    assert (np.abs(got[128:] - tail) <= rel * terms + floor).all()


def test_strided_gemv_then_trsv_share_the_workspace(oracle, ab, handle):
    """GEMV packs a strided x into the region TRSV uses as its progress vector;
    the next solve must find it armed again."""
    n, lda = 900, 904
    LU = lu_fixture(n, seed=41, lda=lda)
    A = oracle.convert(LU, np.float32)
    b = stored(oracle, n, torch.float32, seed=15)
    exact = oracle.exact_trsv(A, n, lda, b, False, True)
    Ad = dev(A)
    for _ in range(2):
        m, k, incx = 64, 5000, 3
        G = stored(oracle, m * k, torch.float32, first=123)
        xg = stored(oracle, k * incx, torch.float32, first=10 ** 6)
        yg = np.zeros(m, dtype=np.float32)
        got = run_gemv(handle, torch.float64, G, m, k, k, xg, 1.0, 0.0, yg, incx=incx)
        ex = oracle.exact_gemv(G, m, k, k, xg, 1.0, 0.0, yg, incx, 1)
        assert oracle.l1_rel_error(ex, got) <= GEMV_TOL[(torch.float64, torch.float32)]
        xd = dev(b)
        handle.trsv(torch.float64, ab.LOWER, ab.UNIT, n, Ad, lda, xd, 1)
        torch.cuda.synchronize()
        assert oracle.l1_rel_error(exact, host(xd)) < 2e-6 * max(1, n / 300)


def test_trsv_repeated_calls_and_strided_x(oracle, ab, handle):
    n, lda, incx = 700, 704, 2
    LU = lu_fixture(n, seed=5, lda=lda)
    A = oracle.convert(LU, np.float32)
    b = stored(oracle, n * incx, torch.float32, seed=11)
    exact = oracle.exact_trsv(A, n, lda, b, False, True, incb=incx)
    Ad = dev(A)
    outs = []
    for _ in range(3):  # the workspace must re-arm itself between calls
        xd = dev(b)
        handle.trsv(torch.float64, ab.LOWER, ab.UNIT, n, Ad, lda, xd, incx)
        torch.cuda.synchronize()
        outs.append(host(xd))
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[1], outs[2])
    assert np.array_equal(outs[0][1::incx], b[1::incx])
    assert oracle.l1_rel_error(exact, outs[0][::incx].copy()) < 2e-6
    # a smaller and a larger system afterwards
    for n2 in (130, 1500):
        LU2 = lu_fixture(n2, seed=n2)
        A2 = oracle.convert(LU2, np.float32)
        b2 = stored(oracle, n2, torch.float32, seed=12)
        xd = dev(b2)
        handle.trsv(torch.float64, ab.LOWER, ab.UNIT, n2, dev(A2), n2, xd, 1)
        torch.cuda.synchronize()
        ex2 = oracle.exact_trsv(A2, n2, n2, b2, False, True)
        assert oracle.l1_rel_error(ex2, host(xd)) < 2e-6 * max(1, n2 / 300)


# ---------------------------------------------------------------------------
# multi-GPU DOT with the all-reduce inside the kernel (peer mailboxes)
# ---------------------------------------------------------------------------
def _peer_group(ab, devices):
    handles = [ab.Handle(d) for d in devices]
    boxes = [h.peer_mailbox() for h in handles]
    for r, h in enumerate(handles):
        h.peer_connect_ptrs(len(devices), r, boxes, devices)
    return handles


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("ar,st", [(torch.float64, torch.float32), (torch.float32, torch.float16),
                                   (torch.float64, torch.float64)])
def test_dot_allreduce_in_kernel_matches_rank_order_sum(oracle, ab, world, ar, st):
    """`world` ranks as handles of ONE process (all on the visible devices,
    round robin; several ranks may share a GPU -- their kernels then run one
    after the other and the earlier one's last CTA waits for the later one).
    Every rank must return the rank-order sum of the per-rank partials, with
    identical bits, call after call (the mailbox alternates between two epochs)."""
    ndev = torch.cuda.device_count()
    devices = [r % ndev for r in range(world)]
    handles = _peer_group(ab, devices)
    n = 1_000_003
    x = stored(oracle, n, st, seed=21)
    y = stored(oracle, n, st, seed=22)
    from accessor_blas_b200.sharded import range_partition
    parts = [range_partition(n, world, r) for r in range(world)]
    streams = [torch.cuda.Stream(device=d) for d in devices]
    xs, ys, outs, partials = [], [], [], []
    for r, (first, count) in enumerate(parts):
        with torch.cuda.device(devices[r]):
            xs.append(torch.from_numpy(x[first:first + count].copy()).cuda(devices[r]))
            ys.append(torch.from_numpy(y[first:first + count].copy()).cuda(devices[r]))
            outs.append(torch.zeros(1, dtype=ar, device=f"cuda:{devices[r]}"))
            p = torch.zeros(1, dtype=ar, device=f"cuda:{devices[r]}")
            handles[r].dot(ar, count, xs[r], 1, ys[r], 1, p)
            partials.append(p.cpu())
    want = torch.zeros(1, dtype=ar)
    for p in partials:  # rank order, in the arithmetic type
        want = want + p
    for call in range(3):
        for r, (first, count) in enumerate(parts):
            with torch.cuda.device(devices[r]):
                outs[r].zero_()
                handles[r].dot_allreduce(ar, count, xs[r], 1, ys[r], 1, outs[r], streams[r])
        for d in set(devices):
            torch.cuda.synchronize(d)
        for r in range(world):
            got = outs[r].cpu()
            assert torch.equal(got, want), (call, r, got.item(), want.item())
    exact = oracle.exact_dot(x, y)
    scale = float(np.abs(x.astype(np.float64) * y.astype(np.float64)).sum())
    tol = 5e-14 if ar == torch.float64 else 2e-5
    assert abs(want.item() - exact) <= tol * scale


# ---------------------------------------------------------------------------
# host-buffer entry points (the e2e path of bench.py)
# ---------------------------------------------------------------------------
def test_host_entry_points(oracle, ab, handle):
    m, n = 300, 1001
    A = stored(oracle, m * n, torch.float32)
    x = stored(oracle, n, torch.float32, first=m * n)
    y = stored(oracle, m, torch.float32, first=m * n + n)
    exact = oracle.exact_gemv(A, m, n, n, x, 1.0, 1.0, y)
    out = y.copy()
    handle.gemv_host(torch.float64, m, n, 1.0, A, n, x, 1, 1.0, out, 1)
    assert oracle.l1_rel_error(exact, out) < 6e-8
    res = np.zeros(1, dtype=np.float64)
    handle.dot_host(torch.float64, n, x, 1, A[:n].copy(), 1, res)
    assert abs(res[0] - oracle.exact_dot(x, A[:n].copy())) < 1e-12
    nt = 400
    LU = lu_fixture(nt, seed=77)
    At = oracle.convert(LU, np.float32)
    b = stored(oracle, nt, torch.float32, seed=5)
    sol = b.copy()
    handle.trsv_host(torch.float64, ab.LOWER, ab.UNIT, nt, At, nt, sol, 1)
    assert oracle.l1_rel_error(oracle.exact_trsv(At, nt, nt, b, False, True), sol) < 3e-6


def test_l1_error_metric_on_device(oracle, handle):
    n = 100_001
    ref = oracle.uniform(n, seed=1)
    res = (ref * (1 + 1e-7)).astype(np.float32)
    got = handle.l1_error(n, dev(ref), 1, dev(res), 1)
    want = oracle.l1_rel_error(ref, res)
    assert got == pytest.approx(want, rel=1e-10)


# ---------------------------------------------------------------------------
# every launch shape the tuner can select must give the same answers
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("stages", [0, 2, 3, 4])
@pytest.mark.parametrize("variant", [2, 3, 4, 5])
def test_gemv_all_launch_shapes(oracle, ab, handle, variant, stages):
    shapes = [(515, 1030, 1032), (37, 4100, 4104), (1000, 8192, 8192)]
    try:
        for unroll in (1, 2, 4):
            ab.tune("gemv_variant", variant)
            ab.tune("gemv_stages", stages)
            ab.tune("gemv_unroll", unroll)
            for m, n, lda in shapes:
                for ar, st in ((torch.float64, torch.float32), (torch.float64, torch.float16),
                               (torch.float32, torch.float16), (torch.float64, torch.float64)):
                    A = stored(oracle, m * lda, st)
                    x = stored(oracle, n, st, first=m * lda)
                    y = stored(oracle, m, st, first=m * lda + n)
                    got = run_gemv(handle, ar, A, m, n, lda, x, 1.0, 1.0, y)
                    exact = oracle.exact_gemv(A, m, n, lda, x, 1.0, 1.0, y)
                    err = oracle.l1_rel_error(exact, got)
                    assert err <= GEMV_TOL[(ar, st)], (variant, stages, unroll, m, n, err)
    finally:
        ab.tune("gemv_variant", 0)
        ab.tune("gemv_stages", 0)
        ab.tune("gemv_unroll", 2)


@pytest.mark.parametrize("pipe", [0, 1, 3, 4])
def test_gemv_pipelined_streams_are_bit_identical(oracle, ab, handle, pipe):
    """The register-pipelined (1) and cp.async-ring (3, 4 stages) streaming
    loops own the same chunks and use the same accumulators as the plain loop
    (0): results must be bit-identical, for every pair, incl. ragged edges."""
    shapes = [(515, 1030, 1032), (37, 4100, 4104), (1000, 8192, 8192), (5, 16384, 16384)]
    try:
        for m, n, lda in shapes:
            for ar, st in ((torch.float64, torch.float32), (torch.float64, torch.float16),
                           (torch.float32, torch.float16), (torch.float64, torch.float64),
                           (torch.float32, torch.float32), (torch.float32, torch.float64)):
                A = stored(oracle, m * lda, st)
                x = stored(oracle, n, st, first=m * lda)
                y = stored(oracle, m, st, first=m * lda + n)
                ab.tune("gemv_variant", 4)
                ab.tune("gemv_pipe", 0)
                want = run_gemv(handle, ar, A, m, n, lda, x, 1.0, 1.0, y)
                ab.tune("gemv_pipe", pipe)
                for iw in ((0, 1, 2, 3, 4) if pipe == 1 and st == torch.float16 and ar == torch.float64 else (3,)):
                    ab.tune("gemv_intwords", iw)
                    got = run_gemv(handle, ar, A, m, n, lda, x, 1.0, 1.0, y)
                    assert np.array_equal(got, want), (pipe, iw, m, n, ar, st)
    finally:
        ab.tune("gemv_variant", 0)
        ab.tune("gemv_pipe", -1)
        ab.tune("gemv_intwords", 2)


def test_gemv_fp16_fp64_fast_path_is_bit_identical_and_handles_non_finite(oracle, handle):
    """The integer widening used for Acc<fp64,fp16> must equal the plain
    conversion bit for bit (checked against the reference's summation order is
    not possible, so against the same kernel with the chunk loop disabled via a
    short n) and fall back when Inf/NaN halves appear."""
    m, n = 64, 8192
    rng = np.random.default_rng(11)
    vals = rng.uniform(-1, 1, m * n)
    vals[::97] *= 1e-6          # fp16 subnormals
    vals[5::1013] = 0.0
    A = oracle.convert(vals, np.float16)
    x = oracle.convert(rng.uniform(-4, 4, n), np.float16)
    y = np.zeros(m, dtype=np.float16)
    got = run_gemv(handle, torch.float64, A, m, n, n, x, 1.0, 0.0, y)
    exact = oracle.exact_gemv(A, m, n, n, x, 1.0, 0.0, y)
    assert np.array_equal(got, exact.astype(np.float16)) or \
        oracle.l1_rel_error(exact, got) < 3.5e-4
    # non-finite entries: rows containing them must come out non-finite exactly
    # like the straightforward computation, the others untouched
    A2 = A.copy().reshape(m, n)
    A2[3, 100] = np.inf
    A2[7, 4097] = np.nan
    A2[9, 12] = -np.inf
    got2 = run_gemv(handle, torch.float64, A2.reshape(-1), m, n, n, x, 1.0, 0.0, y)
    want2 = (A2.astype(np.float64) @ x.astype(np.float64)).astype(np.float16)
    assert np.isposinf(got2[3]) == np.isposinf(want2[3]) and not np.isfinite(got2[3])
    assert np.isnan(got2[7]) and not np.isfinite(got2[9])
    keep = np.ones(m, dtype=bool)
    keep[[3, 7, 9]] = False
    assert np.array_equal(got2[keep], got[keep])


# ---------------------------------------------------------------------------
# programmatic dependent launch: producer -> consumer chains on one stream
# ---------------------------------------------------------------------------
def test_pdl_chains_match_serialised_launches(oracle, ab, handle):
    """With PDL the next kernel's CTAs become resident while the previous
    kernel still writes y; nothing it wrote may be read before
    griddepcontrol.wait.  GEMV -> GEMV ping-pong (the output of one is the x of
    the next, beta != 0 so y is read too) and GEMV -> DOT, 40 links each,
    bit for bit against the same chain with PDL switched off."""
    m = n = 4096 + 256
    st, ar = torch.float32, torch.float64
    A = dev(stored(oracle, m * n, st, seed=61) * np.float32(0.02))
    v0 = dev(stored(oracle, n, st, seed=62))
    w0 = dev(stored(oracle, n, st, seed=63))

    def chain(pdl):
        ab.tune("gemv_pdl", pdl)
        ab.tune("dot_pdl", pdl)
        v, w = v0.clone(), w0.clone()
        dots = torch.zeros(40, dtype=torch.float64, device=DEV)
        for i in range(40):
            handle.gemv(ar, m, n, 1.0, A, n, v, 1, 0.5, w, 1)      # w = A v + w/2
            handle.dot(ar, n, w, 1, v, 1, dots[i:i + 1])            # reads what GEMV just wrote
            handle.gemv(ar, m, n, 1.0, A, n, w, 1, 0.25, v, 1)     # v = A w + v/4
        torch.cuda.synchronize()
        return v.clone(), w.clone(), dots.clone()

    try:
        ref = chain(0)
        for _ in range(3):
            got = chain(1)
            for a, b in zip(ref, got):
                assert torch.equal(a, b)
        assert bool(torch.isfinite(ref[0]).all()) and float(ref[2].abs().max()) > 0
    finally:
        ab.tune("gemv_pdl", 1)
        ab.tune("dot_pdl", 1)
