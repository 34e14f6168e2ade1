"""Is the accuracy of the 6-digit INT8 contraction (csrc/gemm_i8.cu, experimental) enough for the north-star trajectory bar?
A vectorised numpy statement of the same arithmetic (48-bit fixed point per row / column, balanced base-256 digits as the
bytes of (X + 0x8080808080) ^ 0x8080808080, the 26 digit-plane products with p + q <= 6, combination smallest weights first)
replaces the oracle's MTTKRP inside its ALS loop; the fit trajectory over 100 sweeps must stay within 1e-9 of the plain oracle's."""
import numpy as np

from oracle import cpals

NDIG, NDIG_B, NACC, FRAC, FRAC_B = 6, 7, 7, 48, 56


def digits(X, nd):
    """X: int64 array -> list of nd float64 digit planes, most significant first (i8_fields<ND> in gemm_i8.cu)"""
    C = sum(0x80 << (8 * j) for j in range(nd - 1))
    Z = (X + C) ^ C
    planes = [((Z >> (8 * j)) & 0xff).astype(np.uint8).view(np.int8).reshape(X.shape) for j in range(nd)]   # lowest plane first; the top byte is signed
    out = [p.astype(np.float64) for p in planes[::-1]]
    assert np.array_equal(sum(out[p].astype(np.int64) << (8 * (nd - 1 - p)) for p in range(nd)), X)         # the digits ARE X
    return out


def exponents(A, axis):
    amax = np.max(np.abs(A), axis=axis, keepdims=True)
    f, e = np.frexp(np.where(amax > 0, amax, 1.0))        # amax = f 2^e, f in [0.5, 1)
    return e + 1 + (f >= 1.0 - 2.0 ** -7)                 # |a| 2^-E < 1/2 - 2^-8 (i8_exponent: six leading mantissa ones -> next exponent)


def gemm_i8(A, B):
    """A (M x K) @ B (K x N) through the digit-split scheme"""
    ea, eb = exponents(A, 1), exponents(B, 0)
    XA = np.rint(np.ldexp(A, FRAC - ea)).astype(np.int64)
    XB = np.rint(np.ldexp(B, FRAC_B - eb)).astype(np.int64)
    dA, dB = digits(XA, NDIG), digits(XB, NDIG_B)
    v = np.zeros((A.shape[0], B.shape[1]))
    for t in range(NACC - 1, -1, -1):
        acc = np.zeros_like(v)
        for p in range(NDIG):
            if 0 <= t - p < NDIG_B:
                acc += dA[p] @ dB[t - p]                  # exact: |entries| < 2^53
        assert np.max(np.abs(acc)) < 2 ** 31              # what the int32 TMEM accumulators must hold
        v = acc * 2.0 ** (-8 * t) + v
    return np.ldexp(v, ea + eb - FRAC - FRAC_B + 8 * (NDIG - 1 + NDIG_B - 1))


def mttkrp_i8(T, factors, n):
    N = T.ndim
    others = [m for m in range(N) if m != n]
    Kr = factors[others[0]]
    for m in others[1:]:
        Kr = (factors[m][:, None, :] * Kr[None, :, :]).reshape(-1, Kr.shape[1])   # earlier modes fastest
    Tn = np.moveaxis(T, n, 0).reshape(T.shape[n], -1, order="F")
    return np.asfortranarray(gemm_i8(Tn, Kr))


def test_gemm_i8_model_accuracy():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((96, 512)) * np.exp2(rng.integers(-8, 9, size=(96, 1)))
    B = rng.standard_normal((512, 40))
    ref = (A.astype(np.longdouble) @ B.astype(np.longdouble)).astype(np.float64)
    assert np.linalg.norm(gemm_i8(A, B) - ref) / np.linalg.norm(ref) < 1e-12


def test_fit_trajectory_with_the_i8_contraction_stays_within_1e9():
    rng = np.random.default_rng(5)
    dims, R, nsweeps = (24, 20, 28), 10, 100
    T = np.asfortranarray(rng.standard_normal(dims))
    cp = cpals.random_CPD(T, R, np.random.default_rng(6))
    nT = float(np.linalg.norm(T))
    ref = cpals.FitCheck(0.0, nsweeps, nT)
    cpals.als_optimize(T, cp, alg=cpals.KRPNormal(), check=ref)
    # the same loop with the digit-split MTTKRP
    f = [x.copy() for x in cp.factors]
    grams = [cpals.gram(x) for x in f]
    fits = []
    lam = None
    for _ in range(nsweeps):
        M = None
        for n in range(3):
            M = mttkrp_i8(T, f, n)
            assert np.linalg.norm(M - cpals.mttkrp_krp_normal(T, f, n)) / np.linalg.norm(M) < 1e-12
            X = cpals.solve_ls_problem(cpals.compute_krp_gram(grams, n), M)
            f[n], lam = cpals.row_norm(X)
            grams[n] = cpals.gram(f[n])
        inner = float(np.sum(M * (f[-1] * lam[None, :])))
        norm2 = cpals.norm_factors(grams, lam)
        fits.append(1.0 - np.sqrt(abs(nT ** 2 + norm2 - 2 * abs(inner))) / nT)
    assert np.max(np.abs(np.array(fits) - np.array(ref.history))) < 1e-9
