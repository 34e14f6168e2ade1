"""Synthetic inputs of bench.py / the full-size parity tests, regenerated on the CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  SURVEY.md 8(d) asks for a counter-based generator so that any
device shard and any CPU slice can regenerate the same values: the device side is `generate_kernel` /
`normal_at` in itensorcpd.jl_b200/csrc/kernels.cu (Philox4x32-10, Salmon et al. 2011, + Box-Muller on the
element's logical column-major index).  This file restates that generator with numpy so that the oracle can
run ALS on the very tensor the GPU generated, without a device-to-host copy.  Values agree with the device to
the last ulp or two of log / sincospi (tests/test_gpu_dense.py::test_generator_matches_cpu_restatement); the
parity bars (1e-12 MTTKRP, 1e-9 fit) are far above that.

The reference has no counterpart (its tests use unseeded `randn`, test/cp_als.jl:9); initial factors follow
src/cpd.jl:48-60 (`randn(I_n, R)` per mode from one generator, column-normalised).
"""
from __future__ import annotations

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def _philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Ten Philox rounds on uint64 arrays holding 32-bit words (kernels.cu: philox4x32_10)."""
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> _S32, p0 & _MASK
        hi1, lo1 = p1 >> _S32, p1 & _MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def normal_pairs(seed: int, pair0: int, npairs: int):
    """z0, z1 of pairs pair0 .. pair0+npairs-1 of stream `seed` (kernels.cu: normal_pair)."""
    idx = np.arange(pair0, pair0 + npairs, dtype=np.uint64)
    c0, c1, c2, c3 = _philox4x32_10(idx & _MASK, idx >> _S32, np.zeros_like(idx), np.zeros_like(idx),
                                    seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    a = (c1 << _S32) | c0
    b = (c3 << _S32) | c2
    u1 = ((a >> np.uint64(11)).astype(np.float64) + 1.0) * (1.0 / 9007199254740992.0)  # (0, 1]
    u2 = (b >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)         # [0, 1)
    rad = np.sqrt(-2.0 * np.log(u1))
    ang = 2.0 * u2
    # sincospi: reduce the argument exactly before multiplying by pi (what the device routine does)
    ang = ang - 2.0 * np.floor(ang * 0.5)
    return rad * _cospi(ang), rad * _sinpi(ang)


def _sinpi(x):
    """sin(pi x) for x in [0, 2) with the exact-argument reductions sinpi uses."""
    x = np.where(x >= 1.0, x - 2.0, x)           # (-1, 1)
    x = np.where(x > 0.5, 1.0 - x, x)
    x = np.where(x < -0.5, -1.0 - x, x)         # [-0.5, 0.5]
    return np.sin(np.pi * x)


def _cospi(x):
    """cos(pi x) = sin(pi (x + 1/2)) with the shift done exactly on the reduced argument"""
    x = np.where(x >= 1.0, x - 2.0, x)           # (-1, 1)
    x = np.abs(x)                                # even
    return np.where(x > 0.5, -np.sin(np.pi * (x - 0.5)), np.cos(np.pi * x))


def normal_range(seed: int, e0: int, count: int, chunk: int = 1 << 24) -> np.ndarray:
    """normal_at(seed, e) for e = e0 .. e0+count-1 (element e is z_{e&1} of pair e>>1)."""
    out = np.empty(count, dtype=np.float64)
    done = 0
    while done < count:
        n = min(chunk, count - done)
        a = e0 + done
        p0, p1 = a >> 1, (a + n - 1) >> 1
        z0, z1 = normal_pairs(seed, p0, p1 - p0 + 1)
        inter = np.empty(2 * (p1 - p0 + 1), dtype=np.float64)
        inter[0::2] = z0
        inter[1::2] = z1
        off = a - 2 * p0
        out[done:done + n] = inter[off:off + n]
        done += n
    return out


def generate_tensor(dims, seed: int = 0, elem_offset: int = 0) -> np.ndarray:
    """The tensor itcpd_generate_tensor(dims, seed, elem_offset) puts on the device, column-major (kernels.cu:
    generate_kernel: logical element e gets normal_at(seed, e + elem_offset))."""
    n = int(np.prod(dims))
    return normal_range(seed, elem_offset, n).reshape(tuple(int(d) for d in dims), order="F")


def init_factors(dims, R: int, seed: int = 1):
    """bench.py's initial guess: randn(I_n, R) from numpy default_rng(seed), one generator shared over the modes,
    column-normalised as src/cpd.jl:48-60."""
    rng = np.random.default_rng(seed)
    out = []
    for I in dims:
        X = np.asfortranarray(rng.standard_normal((int(I), int(R))))
        out.append(np.asfortranarray(X / np.sqrt(np.sum(X * X, axis=0))[None, :]))
    return out
