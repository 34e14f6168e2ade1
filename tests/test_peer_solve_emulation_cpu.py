"""The DEFAULT multi-GPU exchange, executed on the host: api.cu's peer_signal_kernel and solve.cu's row-solve kernels -- the
warp-per-row chol_solve_warp_kernel (fused all-reduce + row solve behind the bounded flag wait peer_wait_all) -- are compiled
verbatim behind
tests/simt_emu.h and run as forked "ranks" over a MAP_SHARED exchange buffer.  Every rank must end with
X = (sum of the ranks' partial M) Gamma^{-1} and with the reduced M stored locally, for several consecutive exchanges
(epoch parity slots)."""
import os
import re
import subprocess

import numpy as np
import pytest
from scipy.linalg import lapack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "itensorcpd.jl_b200", "csrc")
BUILD = os.path.join(ROOT, "oracle", "_build")

HARNESS = r"""
#include "simt_emu.h"
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>
#include <cstdlib>
#define ITCPD_MAX_PEERS 16
#define ITCPD_PEER_TIMEOUT_NS 20000000000ull
#define ITCPD_SOLVE_CHOLESKY 0
#define ITCPD_SOLVE_QRCP 1
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __trap() { abort(); }
static unsigned long long emu_globaltimer() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (unsigned long long)ts.tv_sec * 1000000000ull + ts.tv_nsec; }
double sm_dyn[(65 * 64) + 8 * 1024 + 64] __attribute__((aligned(16)));
namespace itcpd_emu {
%(peersrc)s
%(signal)s
%(solve)s
}
using namespace itcpd_emu;

static std::vector<double> read_doubles(const char *path, size_t n) {
    std::vector<double> v(n);
    FILE *f = fopen(path, "rb");
    if (!f || fread(v.data(), 8, n, f) != n) { fprintf(stderr, "cannot read %%s\n", path); exit(9); }
    fclose(f);
    return v;
}

int main(int argc, char **argv) {
    const int G = atoi(argv[1]), n = atoi(argv[2]), rows = atoi(argv[3]), nexch = atoi(argv[4]);
    const char *dir = argv[5];
    const int ldw = n | 1;
    char path[512];
    snprintf(path, sizeof(path), "%%s/W.bin", dir);
    std::vector<double> W = read_doubles(path, (size_t)ldw * n);
    snprintf(path, sizeof(path), "%%s/piv.bin", dir);
    std::vector<double> pivd = read_doubles(path, n);
    std::vector<int> piv(n);
    for (int i = 0; i < n; ++i) piv[i] = (int)pivd[i];
    int status[3] = {ITCPD_SOLVE_CHOLESKY, n, 0};
    const size_t slot_doubles = (size_t)rows * n;
    const size_t xchg_bytes = ((256 + 2 * slot_doubles * 8) + 4095) & ~(size_t)4095;
    char *shared = (char *)mmap(nullptr, xchg_bytes * G, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (shared == MAP_FAILED) return 2;
    memset(shared, 0, xchg_bytes * G);
    std::vector<pid_t> kids;
    for (int r = 0; r < G; ++r) {
        pid_t pid = fork();
        if (pid == 0) {
            srand(5 + r);
            std::vector<double> X((size_t)rows * n), Mred((size_t)rows * n);
            long long epoch = 0;
            for (int x = 0; x < nexch; ++x) {
                ++epoch;                                                      // api.cu: const int64_t epoch = ++c->peer_epoch;
                const size_t slot_off = 256 + (size_t)(epoch & 1) * slot_doubles * 8;
                snprintf(path, sizeof(path), "%%s/M_%%d_%%d.bin", dir, x, r);
                std::vector<double> Mp = read_doubles(path, slot_doubles);
                if (rand() %% 2) usleep(rand() %% 500);
                memcpy(shared + r * xchg_bytes + slot_off, Mp.data(), slot_doubles * 8);   // second level: my partial -> my slot
                PeerFlags pf;
                memset(&pf, 0, sizeof(pf));
                pf.n = G; pf.rank = r;
                for (int q = 0; q < G; ++q) pf.dst[q] = (long long *)(shared + q * xchg_bytes);
                emu_launch(32, 0, [&] { peer_signal_kernel(pf, epoch); });
                if (rand() %% 2) usleep(rand() %% 500);
                PeerSrc src;
                memset(&src, 0, sizeof(src));
                src.n = G;
                for (int q = 0; q < G; ++q) src.p[q] = (const double *)(shared + q * xchg_bytes + slot_off);
                src.flags = (const volatile long long *)(shared + r * xchg_bytes);
                src.epoch = epoch;
                src.reduced_out = Mred.data();
                const int ctas = (rows + TSW_WARPS - 1) / TSW_WARPS;
                for (int b = 0; b < ctas; ++b)
                    emu_launch(TSW_WARPS * 32, 0, [&] { blockIdx.x = b; chol_solve_warp_kernel<%(E)d>(W.data(), piv.data(), status, src, rows, n, X.data(), 1, 0, 1, QrcpWs{nullptr, nullptr, nullptr}); });

                snprintf(path, sizeof(path), "%%s/X_%%d_%%d.bin", dir, x, r);
                FILE *f = fopen(path, "wb"); fwrite(X.data(), 8, X.size(), f); fclose(f);
                snprintf(path, sizeof(path), "%%s/R_%%d_%%d.bin", dir, x, r);
                f = fopen(path, "wb"); fwrite(Mred.data(), 8, Mred.size(), f); fclose(f);
            }
            _exit(0);
        }
        kids.push_back(pid);
    }
    int bad = 0;
    for (pid_t k : kids) {
        int st = 0;
        waitpid(k, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad++;
    }
    printf(bad ? "PEER_SOLVE_FAILED %%d\n" : "PEER_SOLVE_OK\n", bad);
    return bad ? 1 : 0;
}
"""


def _extract():
    common = open(os.path.join(CSRC, "common.cuh")).read()
    m = re.search(r"struct PeerSrc \{.*?\n\};\n", common, flags=re.S)
    peersrc = m.group(0)
    api = open(os.path.join(CSRC, "api.cu")).read()
    a0 = api.index("struct PeerFlags")
    a1 = api.index("static int peer_signal(")
    signal = api[a0:a1]
    solve = open(os.path.join(CSRC, "solve.cu")).read()
    h0 = solve.index("// fused all-reduce, first half")
    h1 = solve.index("constexpr int CH_THREADS = 256;")
    helper, n = re.subn(r'asm volatile\("mov\.u64 %0, %%globaltimer;" : "=l"\(now\)\);', "now = emu_globaltimer();", solve[h0:h1])
    assert n == 1
    rows_h = open(os.path.join(CSRC, "qrcp_rows.cuh")).read()
    q0 = rows_h.index("struct QrcpWs")
    q1 = rows_h.index("int qrcp_workspace(")
    s0 = solve.index("constexpr int TSW_WARPS = 8;")
    s1 = solve.index("// dynamic shared memory budget")
    body = solve[s0:s1].replace("extern __shared__ double sm_dyn[];", "")
    return peersrc, signal, helper + rows_h[q0:q1] + body


@pytest.mark.parametrize("G,n,rows", [(2, 12, 20), (3, 40, 17)])
def test_default_peer_exchange_and_fused_solve_emulated(tmp_path, G, n, rows):
    peersrc, signal, body = _extract()
    os.makedirs(BUILD, exist_ok=True)
    E = (n + 31) // 32
    cpp, exe = os.path.join(BUILD, f"peer_solve_emu_{E}.cpp"), os.path.join(BUILD, f"peer_solve_emu_{E}")
    open(cpp, "w").write(HARNESS % {"peersrc": peersrc, "signal": signal, "solve": body, "E": E})
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-I", os.path.join(ROOT, "tests"), "-o", exe, cpp, "-lpthread"], check=True,
                   capture_output=True)
    rng = np.random.default_rng(n)
    A = rng.standard_normal((3 * n, n)); A /= np.linalg.norm(A, axis=0)
    B = rng.standard_normal((3 * n, n)); B /= np.linalg.norm(B, axis=0)
    Gam = np.asfortranarray((A.T @ A) * (B.T @ B))
    c, piv, rank, info = lapack.dpstrf(Gam, tol=1e-6, lower=0)
    assert rank == n
    ldw = n | 1
    W = np.zeros((ldw, n), order="F")
    W[:n, :] = np.triu(c)
    W.reshape(-1, order="F").tofile(tmp_path / "W.bin")
    (piv - 1).astype(np.float64).tofile(tmp_path / "piv.bin")
    nexch = 4
    parts = {}
    for x in range(nexch):
        for r in range(G):
            parts[x, r] = np.asfortranarray(rng.standard_normal((rows, n)))
            parts[x, r].reshape(-1, order="F").tofile(tmp_path / f"M_{x}_{r}.bin")
    out = subprocess.run([exe, str(G), str(n), str(rows), str(nexch), str(tmp_path)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "PEER_SOLVE_OK" in out.stdout, out.stdout[-1000:] + out.stderr[-3000:]
    for x in range(nexch):
        M = np.zeros((rows, n))
        for r in range(G):
            M = M + parts[x, r]           # rank order, like the kernel
        want = np.linalg.solve(Gam, M.T).T
        first = None
        for r in range(G):
            X = np.fromfile(tmp_path / f"X_{x}_{r}.bin").reshape((rows, n), order="F")
            Mred = np.fromfile(tmp_path / f"R_{x}_{r}.bin").reshape((rows, n), order="F")
            assert np.array_equal(Mred, M), (x, r)
            assert np.linalg.norm(X - want) / np.linalg.norm(want) < 1e-10 * max(10.0, np.linalg.cond(Gam)), (x, r)
            first = X if first is None else first
            assert np.array_equal(X, first), "ranks must hold bitwise identical solutions"
