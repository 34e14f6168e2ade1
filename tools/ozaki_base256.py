"""Digit-split comparison behind the choice made in csrc/gemm_i8.cu: 7 balanced base-128 digits with the 28 products p + q <= 6
(the first draft) against 6 balanced base-256 digits with the 21 products p + q <= 5 or the 26 products p + q <= 6 (adopted:
6 bytes per element instead of 7, fewer products, better accuracy).  Exact integer digit products, 80-bit reference, T ~ N(0,1),
Khatri-Rao of normalised factors; prints (products, relative Frobenius error, largest |accumulator|).
  python tools/ozaki_base256.py >> profiles/r1_ozaki_int8_numerics.txt"""
import numpy as np
def split(A, axis, ndig, W):
    B = 1 << W; H = B >> 1
    amax = np.max(np.abs(A), axis=axis, keepdims=True)
    e = np.ceil(np.log2(np.where(amax > 0, amax, 1.0))) + 1          # |A| 2^-e < 1/2
    X = np.rint(np.ldexp(A, (-e + W * ndig - (1 if W == 8 else 0)).astype(np.int64))).astype(object)   # base 256: one bit of headroom for the top digit
    digs = []
    for i in range(ndig):
        if i == ndig - 1:
            d = X.copy(); X = X * 0
            assert max(abs(int(v)) for v in d.ravel()) <= H
        else:
            d = np.vectorize(lambda x: ((int(x) + H) % B) - H, otypes=[object])(X)
            X = np.vectorize(lambda x, dd: (int(x) - int(dd)) // B, otypes=[object])(X, d)
        digs.append(d.astype(np.float64))
    return digs[::-1], e - (1 if W == 8 else 0) * 0, (W * ndig - (1 if W == 8 else 0))
def run(M, K, R, W, nd, tmax):
    rng = np.random.default_rng(0)
    T = rng.standard_normal((M, K))
    I1 = int(np.sqrt(K))
    f1 = rng.standard_normal((I1, R)); f1 /= np.linalg.norm(f1, axis=0)
    f2 = rng.standard_normal((K // I1, R)); f2 /= np.linalg.norm(f2, axis=0)
    Kr = (f1[:, None, :] * f2[None, :, :]).reshape(-1, R)[:K]
    ref = (T.astype(np.longdouble) @ Kr.astype(np.longdouble)); nref = float(np.linalg.norm(ref.astype(np.float64)))
    Td, eT, fT = split(T, 1, nd, W); Kd, eK, fK = split(Kr, 0, nd, W)
    acc = np.zeros((M, R), dtype=np.longdouble); nprod = 0; big = 0.0
    for t in range(0, tmax + 1):
        S = np.zeros((M, R))
        for p in range(nd):
            q = t - p
            if 0 <= q < nd:
                S += Td[p] @ Kd[q]; nprod += 1
        big = max(big, float(np.abs(S).max()))
        acc += np.ldexp(S.astype(np.longdouble), W * (2 * (nd - 1) - t))
    C = acc * np.exp2((eT + eK - fT - fK).astype(np.longdouble))
    return nprod, float(np.linalg.norm((C - ref).astype(np.float64))) / nref, big
for (M, K) in [(256, 1024), (128, 16384)]:
    print(M, K, "base128 7 digits p+q<=6 :", run(M, K, 64, 7, 7, 6))
    print(M, K, "base256 6 digits p+q<=5 :", run(M, K, 64, 8, 6, 5))
    print(M, K, "base256 6 digits p+q<=6 :", run(M, K, 64, 8, 6, 6))
