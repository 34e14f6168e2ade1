// Standalone hardware probe for the building blocks of csrc/gemm_i8.cu (round-2 opener; seconds on one B200):
//   one CTA, operands written to shared memory with ordinary stores in the kernel's canonical layouts, ONE or a few
//   tcgen05.mma kind::i8 instructions through the kernel's own descriptor / PTX wrappers, TMEM read back with the kernel's
//   tcgen05.ld wrapper, compared with exact integer arithmetic on the host.
//     case 0: A K-major  (KIND 0 planes), N = 64
//     case 1: A MN-major (KIND 1 planes), N = 64
//     case 2: A K-major, B = 4 stacked planes, N = 256, D at a column offset, then a second accumulate pass
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -DITCPD_I8_PROBE -o tools/i8_probe tools/i8_probe.cu
#include "../itensorcpd.jl_b200/csrc/gemm_i8.cu"

#include <cstdlib>
#include <vector>

using namespace itcpd;

__global__ void __launch_bounds__(128, 1) i8_probe_kernel(const uint8_t *__restrict__ Ag, const uint8_t *__restrict__ Bg, int a_bytes, int b_bytes,
                                                          int a_mn_major, int n, int col_off, int passes, int *__restrict__ D /* 128 x 512 */) {
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
        const uint64_t adesc = a_mn_major ? i8_smem_desc(sA, 128, 512) : i8_smem_desc(sA, 128, 256);
        const uint64_t bdesc = i8_smem_desc(sB, 128, 256);
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

static int run_case(int id, int a_mn_major, int nplanes, int col_off, int passes) {
    const int n = 64 * nplanes;
    std::vector<int8_t> A(128 * 32), B((size_t)n * 32);
    srand(17 + id);
    for (auto &x : A) x = (int8_t)(rand() % 129 - 64);
    for (auto &x : B) x = (int8_t)(rand() % 129 - 64);
    std::vector<uint8_t> Ab(4096, 0), Bb(16384, 0);
    for (int m = 0; m < 128; ++m)
        for (int k = 0; k < 32; ++k) {
            const int off = a_mn_major ? (k % 8) * 16 + (k / 8) * 128 + (m / 16) * 512 + m % 16 : (m % 8) * 16 + (m / 8) * 256 + (k / 16) * 128 + k % 16;
            Ab[off] = (uint8_t)A[m * 32 + k];
        }
    for (int r = 0; r < n; ++r)
        for (int k = 0; k < 32; ++k) Bb[(r % 8) * 16 + (r / 8) * 256 + (k / 16) * 128 + k % 16] = (uint8_t)B[(size_t)r * 32 + k];
    uint8_t *dA, *dB;
    int *dD;
    cudaMalloc(&dA, 4096); cudaMalloc(&dB, 16384); cudaMalloc(&dD, 128 * 512 * 4);
    cudaMemcpy(dA, Ab.data(), 4096, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bb.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 128 * 512 * 4);
    cudaFuncSetAttribute(i8_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    i8_probe_kernel<<<1, 128, 32768>>>(dA, dB, 4096, n * 32, a_mn_major, n, col_off, passes, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("case %d: CUDA error %s\n", id, cudaGetErrorString(e)); return 1; }
    std::vector<int> D(128 * 512);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
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
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return bad != 0;
}

int main() {
    int bad = 0;
    bad += run_case(0, 0, 1, 0, 1);
    bad += run_case(1, 1, 1, 0, 1);
    bad += run_case(2, 0, 4, 128, 2);
    bad += run_case(3, 1, 3, 64, 1);
    printf(bad ? "I8_PROBE_FAILED\n" : "I8_PROBE_OK\n");
    return bad;
}
