// Standalone hardware probe for the building blocks of csrc/gemm_i8.cu (round-2 opener; seconds on one B200):
//   one CTA, operands written to shared memory with ordinary stores in the kernel's canonical layouts, ONE or a few
//   tcgen05.mma kind::i8 instructions through the kernel's own descriptor / PTX wrappers, TMEM read back with the kernel's
//   tcgen05.ld wrapper, compared with exact integer arithmetic on the host.
//     case 0: A K-major  (KIND 0 planes), N = 64
//     case 1: A MN-major (KIND 1 planes), N = 64
//     case 2: A K-major, B = 4 stacked planes, N = 256, D at a column offset, then a second accumulate pass
//     case 3: A MN-major, B = 3 stacked planes, N = 192, D at column 64
//   If a case fails the probe diagnoses itself in the same run: it searches the (LBO, SBO) pair of each operand that
//   reproduces exact row sums against an all-ones partner, re-checks the full product with the pair found, and prints
//   where single impulses land in TMEM (lane / column map) for the kernel's own descriptors.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -DITCPD_I8_PROBE -o tools/i8_probe tools/i8_probe.cu
#include "../itensorcpd.jl_b200/csrc/gemm_i8.cu"

#include <cstdlib>
#include <vector>

using namespace itcpd;

__global__ void __launch_bounds__(128, 1) i8_probe_kernel(const uint8_t *__restrict__ Ag, const uint8_t *__restrict__ Bg, int a_bytes, int b_bytes,
                                                          int a_mn_major, int n, int col_off, int passes, int a_lbo, int a_sbo, int b_lbo,
                                                          int b_sbo, int *__restrict__ D /* 128 x 512 */) {
    extern __shared__ uint8_t probe_raw[];
    const uint32_t base = (i8_smem_u32(probe_raw) + 1023u) & ~1023u;
    uint8_t *gen = probe_raw + (base - i8_smem_u32(probe_raw));
    const uint32_t sA = base, sB = base + 4096, bar = sB + 16384, slot = bar + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < a_bytes; i += 128) gen[i] = Ag[i];
    for (int i = threadIdx.x; i < b_bytes; i += 128) gen[4096 + i] = Bg[i];
    if (threadIdx.x == 0) {
        i8_mbar_init(bar, 1);
        i8_fence_mbar_init();
    }
    i8_fence_proxy_async();
    if (warp == 0) i8_tmem_alloc(slot, 512);
    i8_tc_fence_before();
    __syncthreads();
    i8_tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(gen + (slot - base));
    if (threadIdx.x == 0) {
        const uint64_t adesc = i8_smem_desc(sA, (uint32_t)a_lbo, (uint32_t)a_sbo);
        const uint64_t bdesc = i8_smem_desc(sB, (uint32_t)b_lbo, (uint32_t)b_sbo);
        for (int p = 0; p < passes; ++p) i8_mma(tmem + (uint32_t)col_off, adesc, bdesc, i8_idesc(n, a_mn_major), p > 0 ? 1u : 0u);
        i8_commit(bar);
    }
    i8_mbar_wait(bar, 0);
    i8_tc_fence_after();
    for (int c0 = 0; c0 < 512; c0 += 32) {
        int v[32];
        i8_tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, v);
        if (c0 >= col_off && c0 < col_off + n)
            for (int c = 0; c < 32; ++c) D[(32 * warp + lane) * 512 + c0 + c] = v[c];
    }
    i8_tc_fence_before();
    __syncthreads();
    if (warp == 0) i8_tmem_dealloc(tmem, 512);
}

// byte offsets of element (row, k) in the layouts the library's kernels write (csrc/gemm_i8.cu)
static int off_kmajor(int row, int k) { return (row % 8) * 16 + (row / 8) * 256 + (k / 16) * 128 + k % 16; }
static int off_mnmajor(int row, int k) { return (k % 8) * 16 + (k / 8) * 128 + (row / 16) * 512 + row % 16; }

struct Probe {
    uint8_t *dA = nullptr, *dB = nullptr;
    int *dD = nullptr;
    Probe() {
        cudaMalloc(&dA, 4096); cudaMalloc(&dB, 16384); cudaMalloc(&dD, 128 * 512 * 4);
        cudaFuncSetAttribute(i8_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);  // room for the widest descriptor of the search
    }
    ~Probe() { cudaFree(dA); cudaFree(dB); cudaFree(dD); }
    // returns false on a CUDA error (sticky: the caller stops)
    bool run(const std::vector<uint8_t> &Ab, const std::vector<uint8_t> &Bb, int a_mn, int n, int col_off, int passes, int a_lbo, int a_sbo, int b_lbo,
             int b_sbo, std::vector<int> &D) {
        cudaMemcpy(dA, Ab.data(), 4096, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, Bb.data(), 16384, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0xff, 128 * 512 * 4);
        i8_probe_kernel<<<1, 128, 200 * 1024>>>(dA, dB, 4096, 16384, a_mn, n, col_off, passes, a_lbo, a_sbo, b_lbo, b_sbo, dD);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return false; }
        D.resize(128 * 512);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        return true;
    }
};

static void random_operands(int seed, int n, std::vector<int8_t> &A, std::vector<int8_t> &B) {
    A.resize(128 * 32); B.resize((size_t)n * 32);
    srand(seed);
    for (auto &x : A) x = (int8_t)(rand() % 129 - 64);
    for (auto &x : B) x = (int8_t)(rand() % 129 - 64);
}
static void lay_out(const std::vector<int8_t> &A, const std::vector<int8_t> &B, int a_mn, int n, std::vector<uint8_t> &Ab, std::vector<uint8_t> &Bb) {
    Ab.assign(4096, 0); Bb.assign(16384, 0);
    for (int m = 0; m < 128; ++m)
        for (int k = 0; k < 32; ++k) Ab[a_mn ? off_mnmajor(m, k) : off_kmajor(m, k)] = (uint8_t)A[m * 32 + k];
    for (int r = 0; r < n; ++r)
        for (int k = 0; k < 32; ++k) Bb[off_kmajor(r, k)] = (uint8_t)B[(size_t)r * 32 + k];
}

static int run_case(Probe &P, int id, int a_mn_major, int nplanes, int col_off, int passes) {
    const int n = 64 * nplanes;
    std::vector<int8_t> A, B;
    random_operands(17 + id, n, A, B);
    std::vector<uint8_t> Ab, Bb;
    lay_out(A, B, a_mn_major, n, Ab, Bb);
    std::vector<int> D;
    if (!P.run(Ab, Bb, a_mn_major, n, col_off, passes, 128, a_mn_major ? 512 : 256, 128, 256, D)) return -1;
    long bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int r = 0; r < n; ++r) {
            int ref = 0;
            for (int k = 0; k < 32; ++k) ref += (int)A[m * 32 + k] * (int)B[(size_t)r * 32 + k];
            ref *= passes;
            if (D[m * 512 + col_off + r] != ref) {
                if (bad < 5) printf("  case %d mismatch at (m=%d, r=%d): got %d want %d\n", id, m, r, D[m * 512 + col_off + r], ref);
                ++bad;
            }
        }
    printf("case %d (A %s-major, N=%d, column offset %d, %d pass(es)): %s (%ld mismatches)\n", id, a_mn_major ? "MN" : "K", n, col_off, passes, bad ? "FAIL" : "ok", bad);
    return bad != 0;
}

// Self-diagnosis after a failed case, so that ONE hardware run says what to change:
//  stage 1: B = all ones (its layout cannot matter)  ->  D[m, r] = sum_k A[m, k]: search A's (LBO, SBO)
//  stage 2: A = all ones                             ->  D[m, r] = sum_k B[r, k]: search B's (LBO, SBO)
//  stage 3: full product with the pair found (checks that the k order of the two operands pairs up)
//  and, if nothing matches, the accumulator map of single impulses: where does A[m0, k0] * B[r0, k0] land in TMEM?
static int diagnose(Probe &P, int a_mn) {
    static const int cand[] = {16, 32, 64, 128, 256, 512, 1024, 2048, 4096};
    const int nc = sizeof(cand) / sizeof(cand[0]), n = 64;
    std::vector<int8_t> A, B;
    random_operands(91 + a_mn, n, A, B);
    std::vector<uint8_t> Ab, Bb, ones(16384, 1), onesA(4096, 1);
    lay_out(A, B, a_mn, n, Ab, Bb);
    std::vector<int> D;
    int fa_l = -1, fa_s = -1, fb_l = -1, fb_s = -1;
    printf("diagnose: A %s-major\n", a_mn ? "MN" : "K");
    for (int i = 0; i < nc && fa_l < 0; ++i)
        for (int j = 0; j < nc && fa_l < 0; ++j) {
            if (!P.run(Ab, ones, a_mn, n, 0, 1, cand[i], cand[j], 128, 256, D)) return -1;
            bool ok = true;
            for (int m = 0; m < 128 && ok; ++m) {
                int ref = 0;
                for (int k = 0; k < 32; ++k) ref += A[m * 32 + k];
                for (int r = 0; r < n && ok; ++r) ok = D[m * 512 + r] == ref;
            }
            if (ok) { fa_l = cand[i]; fa_s = cand[j]; }
        }
    printf("  A descriptor reproducing the row sums of A: LBO=%d SBO=%d (the kernel uses LBO=128 SBO=%d)\n", fa_l, fa_s, a_mn ? 512 : 256);
    for (int i = 0; i < nc && fb_l < 0; ++i)
        for (int j = 0; j < nc && fb_l < 0; ++j) {
            if (!P.run(onesA, Bb, a_mn, n, 0, 1, 128, a_mn ? 512 : 256, cand[i], cand[j], D)) return -1;
            bool ok = true;
            for (int r = 0; r < n && ok; ++r) {
                int ref = 0;
                for (int k = 0; k < 32; ++k) ref += B[(size_t)r * 32 + k];
                for (int m = 0; m < 128 && ok; ++m) ok = D[m * 512 + r] == ref;
            }
            if (ok) { fb_l = cand[i]; fb_s = cand[j]; }
        }
    printf("  B descriptor reproducing the row sums of B: LBO=%d SBO=%d (the kernel uses LBO=128 SBO=256)\n", fb_l, fb_s);
    if (fa_l > 0 && fb_l > 0) {
        if (!P.run(Ab, Bb, a_mn, n, 0, 1, fa_l, fa_s, fb_l, fb_s, D)) return -1;
        long bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int r = 0; r < n; ++r) {
                int ref = 0;
                for (int k = 0; k < 32; ++k) ref += (int)A[m * 32 + k] * (int)B[(size_t)r * 32 + k];
                bad += D[m * 512 + r] != ref;
            }
        printf("  full product with these descriptors: %ld mismatches\n", bad);
    }
    // impulse map with the kernel's own descriptors: one nonzero in A and a full column of ones in B at the same k
    for (int t = 0; t < 6; ++t) {
        const int m0 = (t * 37 + 5) % 128, k0 = (t * 11 + 3) % 32;
        std::vector<uint8_t> Ai(4096, 0), Bi(16384, 0);
        Ai[a_mn ? off_mnmajor(m0, k0) : off_kmajor(m0, k0)] = 1;
        for (int r = 0; r < n; ++r) Bi[off_kmajor(r, k0)] = (uint8_t)(r + 1);
        if (!P.run(Ai, Bi, a_mn, n, 0, 1, 128, a_mn ? 512 : 256, 128, 256, D)) return -1;
        int nz = 0, lane0 = -1, col0 = -1, val0 = 0;
        for (int m = 0; m < 128; ++m)
            for (int r = 0; r < n; ++r)
                if (D[m * 512 + r] != 0) { if (!nz) { lane0 = m; col0 = r; val0 = D[m * 512 + r]; } ++nz; }
        printf("  impulse A[m=%d,k=%d] x B[r,k]=r+1: %d nonzeros, first at lane %d column %d value %d (want %d nonzeros in lane %d, value = column + 1)\n",
               m0, k0, nz, lane0, col0, val0, n, m0);
    }
    return 0;
}

int main() {
    Probe P;
    int bad = 0, r;
    const int cases[4][4] = {{0, 1, 0, 1}, {1, 1, 0, 1}, {0, 4, 128, 2}, {1, 3, 64, 1}};
    for (int i = 0; i < 4; ++i) {
        r = run_case(P, i, cases[i][0], cases[i][1], cases[i][2], cases[i][3]);
        if (r < 0) { printf("I8_PROBE_FAILED (CUDA error, context lost)\n"); return 2; }
        bad += r;
    }
    if (bad) {
        if (diagnose(P, 0) < 0 || diagnose(P, 1) < 0) { printf("I8_PROBE_FAILED (CUDA error during diagnosis)\n"); return 2; }
    }
    printf(bad ? "I8_PROBE_FAILED\n" : "I8_PROBE_OK\n");
    return bad;
}
