"""The kernels of csrc/peer_graph.cu (device-side exchange epochs, flag wait, peer small all-reduce), compiled VERBATIM for the
host behind tests/simt_emu.h and run as G forked processes -- one per "GPU" -- that share the exchange buffers through a
MAP_SHARED mapping, the way the real ranks share them through CUDA IPC.  Every slot value is tagged with (sweep, exchange,
rank), so a stale or torn read, a wrong offset or a lost flag fails; a dead-lock trips the kernels' own bounded wait.
These kernels have not run on hardware yet (option peer_graph, off by default)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "itensorcpd.jl_b200", "csrc", "peer_graph.cu")
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
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __trap() { abort(); }
static unsigned long long emu_globaltimer() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (unsigned long long)ts.tv_sec * 1000000000ull + ts.tv_nsec; }
namespace itcpd_emu {
%(kernels)s
}
using namespace itcpd_emu;

static double tag(int sweep, int x, int rank, int i) { return sweep * 1000.0 + x * 100.0 + rank + i * 0.001; }

int main(int argc, char **argv) {
    const int G = atoi(argv[1]), nexch = atoi(argv[2]), sweeps = atoi(argv[3]);
    const int64_t slot_doubles = 96, small_doubles = 40;
    const int nslots = nexch > 2 ? nexch : 2;
    const size_t small_off = 256 + (size_t)nslots * slot_doubles * 8;
    const size_t xchg_bytes = ((small_off + 2 * small_doubles * 8) + 4095) & ~(size_t)4095;
    char *shared = (char *)mmap(nullptr, xchg_bytes * G, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (shared == MAP_FAILED) return 2;
    memset(shared, 0, xchg_bytes * G);
    std::vector<pid_t> kids;
    for (int r = 0; r < G; ++r) {
        pid_t pid = fork();
        if (pid == 0) {
            srand(77 + r);
            PeerPtrs f;
            memset(&f, 0, sizeof(f));
            f.n = G; f.rank = r;
            for (int q = 0; q < G; ++q) f.base[q] = shared + q * xchg_bytes;
            long long epochs[2] = {0, 0};   // private "device memory" of this rank
            double buf[64];
            for (int s = 0; s < sweeps; ++s) {
                for (int x = 0; x < nexch; ++x) {
                    if (rand() %% 3 == 0) usleep(rand() %% 300);
                    double *slot = (double *)(f.base[r] + 256 + (size_t)x * slot_doubles * 8);   // second-level kernel: my partial
                    for (int i = 0; i < slot_doubles; ++i) slot[i] = tag(s, x, r, i);
                    emu_launch(32, 0, [&] { peer_signal_dev_kernel(f, &epochs[0]); });
                    if (rand() %% 3 == 0) usleep(rand() %% 300);
                    emu_launch(32, 0, [&] { peer_wait_dev_kernel((const volatile long long *)f.base[r], G, &epochs[0]); });
                    for (int q = 0; q < G; ++q) {                                                   // row-solve kernel: sum the peers' slots
                        const volatile double *ps = (const volatile double *)(f.base[q] + 256 + (size_t)x * slot_doubles * 8);
                        for (int i = 0; i < slot_doubles; ++i)
                            if (ps[i] != tag(s, x, q, i)) { fprintf(stderr, "rank %%d sweep %%d exch %%d: peer %%d slot[%%d] = %%f\n", r, s, x, q, i, ps[i]); _exit(3); }
                    }
                }
                for (int which = 0; which < 2; ++which) {                                           // last mode: column norms, Gram
                    const int n = which ? 37 : 5;
                    for (int i = 0; i < n; ++i) buf[i] = r + i + which * 0.5 + s;
                    if (rand() %% 3 == 0) usleep(rand() %% 300);
                    emu_launch(256, 0, [&] { peer_allreduce_small_kernel(f, &epochs[1], small_off, small_doubles, buf, n); });
                    for (int i = 0; i < n; ++i) {
                        double want = 0.0;
                        for (int q = 0; q < G; ++q) want += q + i + which * 0.5 + s;
                        if (buf[i] != want) { fprintf(stderr, "rank %%d sweep %%d allreduce %%d: buf[%%d] = %%f want %%f\n", r, s, which, i, buf[i], want); _exit(4); }
                    }
                }
            }
            if (epochs[0] != (long long)sweeps * nexch || epochs[1] != 2ll * sweeps) _exit(5);
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
    printf(bad ? "PEER_EMU_FAILED %%d\n" : "PEER_EMU_OK\n", bad);
    return bad ? 1 : 0;
}
"""


@pytest.fixture(scope="module")
def harness():
    hdr = open(os.path.join(os.path.dirname(SRC), "peer.cuh")).read()      # PeerPtrs + bounded_wait (csrc/peer.cuh)
    hdr = hdr[hdr.index("struct PeerPtrs"):hdr.rindex("}  // namespace itcpd")]
    text = open(SRC).read()
    start = text.index("__global__ void peer_signal_dev_kernel")
    end = text.index("bool peer_graph_active")
    body = hdr + text[start:end]
    body, n = re.subn(r"static inline PeerPtrs peer_ptrs\(const itcpd_ctx \*c\) \{.*?\n\}\n", "", body, flags=re.S)
    assert n == 1
    body, n = re.subn(r'asm volatile\("mov\.u64 %0, %%globaltimer;" : "=l"\(now\)\);', "now = emu_globaltimer();", body)
    assert n == 1
    os.makedirs(BUILD, exist_ok=True)
    cpp, exe = os.path.join(BUILD, "peer_emu.cpp"), os.path.join(BUILD, "peer_emu")
    open(cpp, "w").write(HARNESS % {"kernels": body})
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "tests"), "-o", exe, cpp, "-lpthread"], check=True, capture_output=True)
    return exe


@pytest.mark.parametrize("G,nexch", [(2, 2), (3, 2), (4, 3)])
def test_peer_graph_kernels_emulated_as_forked_ranks(harness, G, nexch):
    out = subprocess.run([harness, str(G), str(nexch), "25"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "PEER_EMU_OK" in out.stdout, out.stdout[-1000:] + out.stderr[-3000:]
