"""The three pivoted-Cholesky kernels of csrc/solve.cu, compiled VERBATIM for the host behind tests/simt_emu.h (one OS thread
per CUDA thread, pthread barriers for __syncthreads / bar.sync / __syncwarp, scratch-based warp collectives) and run
  * against LAPACK dpstrf (same pivots, rank and factor),
  * against each other (team kernel bitwise == block kernel: the property the GPU test checks on hardware),
  * under ThreadSanitizer (a missing barrier between a shared-memory write and another thread's read is a reported race).
This is the only execution the right-looking kernel (chol_alg=2) has had so far: its first hardware run is a round-2 item."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
from scipy.linalg import lapack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOLVE = os.path.join(ROOT, "itensorcpd.jl_b200", "csrc", "solve.cu")
BUILD = os.path.join(ROOT, "oracle", "_build")

DRIVER = r"""
double sm_dyn[129 * 128 + 256];
inline void team_sync(int nw, int team) { if (nw == 1) __syncwarp(); else emu_named_barrier(); }
#define ITCPD_SOLVE_CHOLESKY 0
#define ITCPD_SOLVE_QRCP 1
namespace itcpd_emu {  // not `itcpd`: libitcpd_b200.so (RTLD_GLOBAL) exports host stubs with the kernels' names
%(kernels)s
}
extern "C" int run_cholesky(int which, int n, double tol, const double *G, double *W, int *piv, int *status) {
    using namespace itcpd_emu;
    const int team = (n + 31) & ~31;
    if (which == 0) {
        const int threads = n <= 64 ? 64 : (n <= 128 ? 128 : 256);
        emu_launch(threads, 0, [&] { pivoted_cholesky_kernel(G, n, tol, W, piv, status, 1); });
    } else if (which == 1) {
        emu_launch(256, team, [&] { pivoted_cholesky_team_kernel(G, n, tol, W, piv, status); });
    } else if (which == 3) {
        emu_launch(256, 0, [&] { pivoted_cholesky_rl2_kernel(G, n, tol, W, piv, status); });     // two threads per column, n <= 128
    } else if (n <= 32) {
        emu_launch(256, 32, [&] { pivoted_cholesky_rl_kernel<32>(G, n, tol, W, piv, status); });
    } else {
        emu_launch(256, 64, [&] { pivoted_cholesky_rl_kernel<64>(G, n, tol, W, piv, status); });
    }
    return 0;
}
"""


def _build(tsan: bool, break_sync: bool = False):
    text = open(SOLVE).read()
    start = text.index("constexpr int CH_THREADS = 256;")
    end = text.index("// One warp per right-hand side")
    body = text[start:end]
    body = body.replace("extern __shared__ double sm_dyn[];", "")
    body, n = re.subn(r"__device__ __forceinline__ void team_sync\(int nw, int team\) \{.*?\n\}\n", "", body, flags=re.S)
    assert n == 1
    os.makedirs(BUILD, exist_ok=True)
    tag = ("tsan" if tsan else "plain") + ("_broken" if break_sync else "")
    cpp, so = os.path.join(BUILD, f"chol_emu_{tag}.cpp"), os.path.join(BUILD, f"chol_emu_{tag}.so")
    driver = DRIVER
    if break_sync:  # negative control: the single-warp team_sync between the swap / publish phase and its readers vanishes
        driver = driver.replace("if (nw == 1) __syncwarp(); else", "if (nw == 1) { } else")
        assert driver != DRIVER
    open(cpp, "w").write('#include "simt_emu.h"\n' + driver % {"kernels": body})
    flags = ["-O1", "-g", "-fsanitize=thread"] if tsan else ["-O2"]
    subprocess.run(["g++", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I", os.path.join(ROOT, "tests"), *flags, "-o", so, cpp,
                    "-lpthread", "-Wl,-Bsymbolic"], check=True, capture_output=True)
    return so


@pytest.fixture(scope="module")
def emu():
    return C.CDLL(_build(False))


def gamma(n, seed, deficient=False, ties=False):
    rng = np.random.default_rng(seed)
    I = max(2 * n, 8)
    A = rng.standard_normal((I, n)); A /= np.linalg.norm(A, axis=0)
    B = rng.standard_normal((I, n)); B /= np.linalg.norm(B, axis=0)
    if deficient and n > 3:
        A[:, 1] = A[:, 0]; B[:, 1] = B[:, 0]
        A[:, n - 1] = A[:, 2]; B[:, n - 1] = B[:, 2]
    G = np.asfortranarray((A.T @ A) * (B.T @ B))
    if ties:
        np.fill_diagonal(G, 1.0)
    return G


def run(lib, which, G, tol=1e-6):
    n = G.shape[0]
    ldw = n | 1
    W = np.zeros(ldw * n)
    piv = np.zeros(n, dtype=np.int32)
    st = np.zeros(3, dtype=np.int32)
    lib.run_cholesky(which, n, C.c_double(tol), G.ctypes.data_as(C.c_void_p), W.ctypes.data_as(C.c_void_p), piv.ctypes.data_as(C.c_void_p),
                     st.ctypes.data_as(C.c_void_p))
    return W.reshape((ldw, n), order="F")[:n, :], piv, st


CASES = [(1, False, False), (2, False, True), (7, False, False), (31, False, True), (32, False, False), (33, True, False),
         (50, False, True), (64, False, False), (64, True, True)]


@pytest.mark.parametrize("n,deficient,ties", CASES)
def test_emulated_kernels_match_lapack_and_each_other(emu, n, deficient, ties):
    G = gamma(n, 100 + n, deficient, ties)
    c, pl, rl, info = lapack.dpstrf(G, tol=1e-6, lower=0)
    res = [run(emu, which, G) for which in (0, 1, 2)]
    for which, (W, piv, st) in enumerate(res):
        assert st[1] == rl and st[0] == (0 if rl == n else 1), (which, st, rl)
        assert np.array_equal(piv[:rl], pl[:rl] - 1), (which, piv, pl - 1)
        U = np.triu(W)[:rl, :]
        assert np.allclose(U, np.triu(c)[:rl, :], rtol=0, atol=1e-10), which
    # the team kernel is the block kernel, operation for operation
    assert np.array_equal(np.triu(res[0][0])[:rl], np.triu(res[1][0])[:rl]) and np.array_equal(res[0][1], res[1][1])
    # the right-looking kernel: same pivots everywhere (bookkeeping of never-eliminated columns included)
    assert np.array_equal(res[2][1], res[1][1])


def test_emulated_team_kernel_up_to_128(emu):
    for n in (65, 100, 128):
        G = gamma(n, 7 + n)
        c, pl, rl, info = lapack.dpstrf(G, tol=1e-6, lower=0)
        W0, p0, s0 = run(emu, 0, G)
        W1, p1, s1 = run(emu, 1, G)
        assert s1[1] == rl and np.array_equal(p1, pl - 1)
        assert np.array_equal(np.triu(W0), np.triu(W1)) and np.array_equal(p0, p1)


@pytest.mark.parametrize("n,deficient,ties", CASES)
def test_two_thread_right_looking_kernel_is_bitwise_the_one_thread_kernel_up_to_64(emu, n, deficient, ties):
    """pivoted_cholesky_rl2_kernel splits every column over two threads but applies the same operations to every element"""
    G = gamma(n, 300 + n, deficient, ties)
    W2, p2, s2 = run(emu, 2, G)
    W3, p3, s3 = run(emu, 3, G)
    assert np.array_equal(s2, s3) and np.array_equal(p2, p3) and np.array_equal(W2, W3)


@pytest.mark.parametrize("n,deficient,ties", [(65, False, False), (66, True, False), (100, False, True), (127, False, False), (128, False, False), (128, True, True)])
def test_two_thread_right_looking_kernel_matches_lapack_up_to_128(emu, n, deficient, ties):
    G = gamma(n, 400 + n, deficient, ties)
    c, pl, rl, info = lapack.dpstrf(G, tol=1e-6, lower=0)
    W, piv, st = run(emu, 3, G)
    assert st[1] == rl and st[0] == (0 if rl == n else 1), (st, rl)
    assert np.array_equal(piv[:rl], pl[:rl] - 1)
    assert np.allclose(np.triu(W)[:rl, :], np.triu(c)[:rl, :], rtol=0, atol=1e-10)
    W1, p1, s1 = run(emu, 1, G)                                  # the shipped team kernel: same pivots, same status
    assert np.array_equal(p1, piv) and np.array_equal(s1, st)
    U = np.triu(W)[:rl, :]
    Gp = G[np.ix_(piv, piv)]
    if rl == n:
        assert np.linalg.norm(U.T @ U - Gp) / np.linalg.norm(Gp) < 1e-14


def test_emulated_nan_and_negative_tolerance(emu):
    G = gamma(12, 5)
    Gn = G.copy(); Gn[3, 3] = np.nan
    for which in (0, 1, 2, 3):
        assert run(emu, which, Gn)[2][1] == 0            # NaN on the diagonal: fail at column 0
        W, piv, st = run(emu, which, G, tol=-1.0)          # LAPACK default tolerance n * eps * max diag (leverage-score path)
        assert st[1] == 12


def _tsan_run(so, cases, kernels):
    code = f"""
import ctypes as C, numpy as np, sys
sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
from test_cholesky_emulation_cpu import gamma, run
lib = C.CDLL({so!r})
for n, d, t in {cases!r}:
    G = gamma(n, 3 + n, d, t)
    for which in {kernels!r}:
        run(lib, which, G)
print("TSAN_RUN_DONE")
"""
    import glob
    rt = sorted(glob.glob("/usr/lib/gcc/x86_64-linux-gnu/*/libtsan.so")) + sorted(glob.glob("/usr/lib/x86_64-linux-gnu/libtsan.so*"))
    if not rt:
        pytest.skip("libtsan not found")
    env = dict(os.environ, LD_PRELOAD=rt[0], TSAN_OPTIONS="report_bugs=1 exitcode=0 halt_on_error=0", PYTHONPATH=ROOT,
               OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")   # numpy's BLAS threads are not instrumented: keep them out of the picture
    out = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=900)
    assert "TSAN_RUN_DONE" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
    return out.stderr


def test_kernels_are_race_free_under_thread_sanitizer():
    err = _tsan_run(_build(True), [(5, False, False), (33, True, True), (64, False, False)], (0, 1, 2, 3))
    assert "WARNING: ThreadSanitizer" not in err, err[-6000:]
    err = _tsan_run(_build(True), [(70, False, True), (128, True, False)], (3,))
    assert "WARNING: ThreadSanitizer" not in err, err[-6000:]


def test_thread_sanitizer_sees_a_missing_warp_sync():
    """Negative control for the test above: without the single-warp __syncwarp of team_sync, the swap / u-publish phase races
    with its readers in both the team and the right-looking kernel, and the emulator + TSAN must say so."""
    so = _build(True, break_sync=True)
    for which in (1, 2):
        # TSAN keeps only a few recent accesses per memory cell, so one run can miss the race: several sizes, one verdict
        err = _tsan_run(so, [(20, False, False), (31, False, False), (27, False, True), (24, False, False)], (which,))
        if "WARNING: ThreadSanitizer: data race" not in err:
            err = _tsan_run(so, [(30, False, False), (22, False, False), (29, False, True), (32, False, False)], (which,))
        assert "WARNING: ThreadSanitizer: data race" in err, which
