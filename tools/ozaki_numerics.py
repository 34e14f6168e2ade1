"""Numerical feasibility of the INT8 tensor-core route past the FP64 DMMA bound: emulate an
Ozaki-style split of both GEMM operands into signed 7-bit digits (what tcgen05 kind::i8 would multiply exactly into int32
TMEM accumulators) on the CPU and measure the MTTKRP error against an 80-bit reference, for the operand statistics of the
CP-ALS path: T ~ N(0,1) (config B), K = Khatri-Rao product of column-normalised factors.
Digits are exact integers; digit-matrix products are done in float64 BLAS, which is exact here (|sum| < 2^53).
Usage: python tools/ozaki_numerics.py [M K R]"""
import sys

import numpy as np

W = 7  # bits per digit (balanced digits in [-64, 63] fit int8)


def split(A, axis, ndig):
    """fixed point w.r.t. the per-row (axis=1) / per-column (axis=0) power-of-two scale, then balanced base-128 digits,
    most significant first.  Returns (digits[ndig, ...] as float64 integers, exponents)."""
    amax = np.max(np.abs(A), axis=axis, keepdims=True)
    e = np.ceil(np.log2(np.where(amax > 0, amax, 1.0))) + 1          # |A| * 2^-e < 1/2
    X = np.rint(np.ldexp(A, (-e + W * ndig).astype(np.int64))).astype(object)   # exact big integers
    digs = []
    for _ in range(ndig):                                               # least significant first
        d = np.vectorize(lambda x: ((int(x) + 64) % 128) - 64, otypes=[object])(X)
        X = np.vectorize(lambda x, dd: (int(x) - int(dd)) // 128, otypes=[object])(X, d)
        digs.append(d.astype(np.float64))
    assert np.all(np.vectorize(lambda x: int(x) == 0)(X)), "top digit overflow"
    return digs[::-1], e                                               # digs[p] has weight 2^(-W (p+1))


def main():
    M, K, R = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (256, 1024, 64)
    rng = np.random.default_rng(0)
    T = rng.standard_normal((M, K))
    I1 = int(np.sqrt(K))
    f1 = rng.standard_normal((I1, R)); f1 /= np.linalg.norm(f1, axis=0)
    f2 = rng.standard_normal((K // I1, R)); f2 /= np.linalg.norm(f2, axis=0)
    Kr = (f1[:, None, :] * f2[None, :, :]).reshape(-1, R)[:K]          # Khatri-Rao rows
    ref = (T.astype(np.longdouble) @ Kr.astype(np.longdouble))
    nref = float(np.linalg.norm(ref.astype(np.float64)))
    e64 = float(np.linalg.norm((T @ Kr).astype(np.longdouble) - ref)) / nref
    print(f"M={M} K={K} R={R}   plain float64 GEMM vs 80-bit reference: {e64:.2e}")
    print(f"{'digits':>6} {'products':>9} {'rel. Frobenius error':>22} {'max |int32 acc|':>16}")
    for ndig in (6, 7, 8, 9):
        Td, eT = split(T, 1, ndig)
        Kd, eK = split(Kr, 0, ndig)
        for tmax in (ndig + 1, ndig + 2):                               # keep digit pairs with (p+1)+(q+1) <= tmax
            acc = np.zeros((M, R), dtype=np.longdouble)
            nprod, biggest = 0, 0.0
            for t in range(2, tmax + 1):                                # one int32 accumulator per weight 2^(-W t)
                S = np.zeros((M, R))
                for p in range(ndig):
                    q = t - 2 - p
                    if 0 <= q < ndig:
                        S += Td[p] @ Kd[q]
                        nprod += 1
                biggest = max(biggest, float(np.max(np.abs(S))))
                acc += np.ldexp(S.astype(np.longdouble), -W * t)
            C = acc * np.exp2((eT + eK).astype(np.longdouble))
            err = float(np.linalg.norm((C - ref).astype(np.float64))) / nref
            print(f"{ndig:6d} {nprod:9d} {err:22.2e} {biggest:16.3g}")


if __name__ == "__main__":
    main()
