// The dominant kernel: partial contraction of the dimension tree as a TMA-staged FP64 DMMA GEMM.
//
// Replaces the R-pass per-rank contraction loop of the reference's MTTKRP
// (src/algebra/had_contract.jl:72-124, reached from algorithms/.../standard/tensor.jl:32-44) and the
// permute+GEMM of KRPNormal (tensor.jl:12-20).  The dense tensor is never permuted or copied: it is
// viewed as a column-major matrix (modes [0,split) x modes [split,N)) and streamed from HBM once per
// pass through 2-D TMA boxes (cp.async.bulk.tensor, 128B swizzle) into a 4-stage mbarrier ring; the
// Khatri-Rao operand is pre-packed in MMA fragment order and staged with 1-D bulk copies.  Eight
// consumer warps issue mma.sync m8n8k4 f64 (SASS DMMA.8x8x4; FP64 has no tcgen05 kind) on 32 x 8*NB
// register tiles; the TMA refill duty rotates over the warps (one elected lane).  Persistent CTAs.
//
//  kind 0 ("A"): out[m, r] = sum_k T[m + M*k]  * K[k, r]   (free index m contiguous in memory)
//  kind 1 ("B"): out[n, r] = sum_m T[m + Mc*n] * K[m, r]   (contraction index m contiguous in memory)
#include "common.cuh"

namespace itcpd {

constexpr int BK = 16;      // contraction tile: 16 doubles = one 128-byte swizzle row
constexpr int STAGES = 4;

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// row permutation of kind-1 A fragments: makes the 8 lanes of a quarter warp hit 8 distinct
// 16-byte bank groups under the 128B swizzle (see DESIGN.md "fragment maps")
__device__ __forceinline__ int sigma_b(int g) { return (g >> 1) + ((g & 1) << 2); }

// ------------------------------------------------------------------------------------------------
// the GEMM kernel
// ------------------------------------------------------------------------------------------------
template <int NB, int WARPS>
struct GemmSmem {
    static constexpr int BM = 32 * WARPS;
    static constexpr int T_BYTES = BM * BK * 8;
    static constexpr int K_BYTES = BK * 8 * NB * 8;
    static constexpr int STAGE_BYTES = T_BYTES + K_BYTES;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 1024 /*alignment slack*/;
};

template <int NB, int WARPS, int KIND>
__global__ void __launch_bounds__(WARPS * 32, (WARPS <= 4 ? 2 : 1))
partial_gemm_kernel(const __grid_constant__ CUtensorMap tmap, const double *__restrict__ Kp, double *__restrict__ out,
                    int64_t rows_out, int R, int num_row_tiles, int num_rblocks, int kt_count, int swz_mask, int map3d,
                    int full_waves, double *__restrict__ slots) {
    using S = GemmSmem<NB, WARPS>;
    constexpr int BM = S::BM;
    constexpr int BN = 8 * NB;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sT0 = smem_base;
    const uint32_t sK0 = smem_base + STAGES * S::T_BYTES;
    const uint32_t bar0 = sK0 + STAGES * S::K_BYTES;  // full[STAGES], empty[STAGES]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar0 + 8 * s, 1);
            mbar_init(bar0 + 8 * (STAGES + s), WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // Hybrid stream-K work split.  The first `full_waves` rounds are data parallel (tile = cta + w * grid: the CTAs
    // of a round stream neighbouring rows of the same k columns, which keeps DRAM pages / TLB entries shared);
    // the tiles of the last, partial round are cut along k into gridDim.x equal ranges so that every SM runs the
    // same number of DMMA iterations whatever the tile count.  A range that begins or ends inside a tile leaves a
    // partial tile in `slots`; streamk_fixup_kernel adds the parts in ascending-k order (deterministic).
    const int G = (int)gridDim.x, cta = (int)blockIdx.x;
    const int dp_iters = full_waves * kt_count;
    const int rem_units = max(0, num_row_tiles * num_rblocks - full_waves * G) * kt_count;
    const int r0 = (int)((int64_t)rem_units * cta / G), r1 = (int)((int64_t)rem_units * (cta + 1) / G);
    const int total_iters = dp_iters + (r1 - r0);
    auto locate = [&](int j, int &tile, int &kt) {
        if (j < dp_iters) {
            const int w = j / kt_count;
            kt = j - w * kt_count;
            tile = cta + w * G;
        } else {
            const int u = r0 + (j - dp_iters);
            const int q = u / kt_count;
            kt = u - q * kt_count;
            tile = full_waves * G + q;
        }
    };

    // Stage refill for pipeline iteration j (one elected lane; the duty rotates over the warps so that
    // no warp's DMMA stream carries the whole TMA-issue overhead).
    auto issue = [&](int j) {
        const int sj = j % STAGES;
        if (j >= STAGES) mbar_wait(bar0 + 8 * (STAGES + sj), (uint32_t)((j / STAGES - 1) & 1));
        int tile, kt;
        locate(j, tile, kt);
        const int row_tile = tile / num_rblocks;
        const int rb = tile - row_tile * num_rblocks;
        const int row0 = row_tile * BM;
        const uint32_t full = bar0 + 8 * sj;
        mbar_expect_tx(full, S::STAGE_BYTES);
        const uint32_t dT = sT0 + sj * S::T_BYTES;
        if (KIND == 0) {
            if (map3d) {
                // one box (16 m, 16 k, BM/16 row groups): lands as [row group][k][16 m], same image as the 2-D loop
                tma_load_3d(dT, &tmap, 0, kt * BK, row0 >> 4, full);
            } else {
#pragma unroll 1
                for (int gg = 0; gg < BM / 16; ++gg) tma_load_2d(dT + gg * (BK * 128), &tmap, row0 + 16 * gg, kt * BK, full);
            }
        } else {
            tma_load_2d(dT, &tmap, kt * BK, row0, full);
        }
        const double *src = Kp + ((size_t)kt * num_rblocks + rb) * (size_t)(BK * 8 * NB);
        bulk_load_1d(sK0 + sj * S::K_BYTES, src, S::K_BYTES, full);
    };

    if (warp == 0 && lane == 0) {
        for (int j = 0; j < STAGES - 1 && j < total_iters; ++j) issue(j);
    }
    __syncwarp();

    const int g = lane >> 2;   // MMA groupID
    const int t = lane & 3;    // MMA threadID_in_group
    int it = 0;

    while (it < total_iters) {
        int tile, kt_begin;
        locate(it, tile, kt_begin);
        const int kt_end = min(kt_count, kt_begin + (total_iters - it));
        const int row_tile = tile / num_rblocks;
        const int rb = tile - row_tile * num_rblocks;
        const int64_t row0 = (int64_t)row_tile * BM;

        double acc[4][NB][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int kt = kt_begin; kt < kt_end; ++kt, ++it) {
            {
                const int j = it + STAGES - 1;
                if (j < total_iters && warp == (it % WARPS)) {
                    if (lane == 0) issue(j);
                    __syncwarp();
                }
            }
            const int stage = it % STAGES;
            mbar_wait(bar0 + 8 * stage, (uint32_t)((it / STAGES) & 1));
            const uint32_t sT = sT0 + stage * S::T_BYTES;
            const uint32_t sK = sK0 + stage * S::K_BYTES;
#pragma unroll
            for (int pair = 0; pair < 2; ++pair) {
                double2 b[NB];
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) b[nb] = lds128(sK + (((pair * NB + nb) * 32 + lane) << 4));
                double a[4][2];  // [row block][k-step of the pair]
                if (KIND == 0) {
                    // smem box = [k row (128 B)][16 m]; one LDS.128 = rows m=2g,2g+1 of a 16-row group at one k
#pragma unroll
                    for (int gq = 0; gq < 2; ++gq) {
                        const uint32_t base = sT + (2 * warp + gq) * (BK * 128);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int rho = 8 * pair + 2 * t + e;
                            const double2 v = lds128(base + rho * 128 + ((g ^ (rho & swz_mask)) << 4));
                            a[2 * gq + 0][e] = v.x;
                            a[2 * gq + 1][e] = v.y;
                        }
                    }
                } else {
                    // smem box = [n row (128 B)][16 m]; one LDS.128 = k = 2c, 2c+1 of one output row
#pragma unroll
                    for (int rbk = 0; rbk < 4; ++rbk) {
                        const int nl = 32 * warp + 8 * rbk + sigma_b(g);
                        const int c = 4 * pair + t;
                        const double2 v = lds128(sT + nl * 128 + ((c ^ (nl & swz_mask)) << 4));
                        a[rbk][0] = v.x;
                        a[rbk][1] = v.y;
                    }
                }
#pragma unroll
                for (int e = 0; e < 2; ++e)
#pragma unroll
                    for (int rbk = 0; rbk < 4; ++rbk)
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb)
                            dmma884(acc[rbk][nb][0], acc[rbk][nb][1], a[rbk][e], e == 0 ? b[nb].x : b[nb].y);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * (STAGES + stage));
        }

        // ---- epilogue: registers -> global.  Full tile: column-major rows_out x R output.  Partial tile (the
        // range started or stopped inside it): dense BM x BN slot, 2*cta (+1 when it holds the head of the tile).
        const bool full_tile = (kt_begin == 0 && kt_end == kt_count);
        double *dst_base;
        int64_t ldo, row_lim;
        int col0, col_lim;
        if (full_tile) {
            dst_base = out + row0;
            ldo = rows_out;
            row_lim = rows_out - row0;
            col0 = BN * rb;
            col_lim = R;
        } else {
            dst_base = slots + (size_t)(2 * cta + (kt_begin == 0 ? 1 : 0)) * (size_t)(BM * BN);
            ldo = BM;
            row_lim = BM;
            col0 = 0;
            col_lim = BN;
        }
        if (KIND == 0) {
#pragma unroll
            for (int gq = 0; gq < 2; ++gq) {
                const int m = 32 * warp + 16 * gq + 2 * g;
                if (m < row_lim) {
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const int r = col0 + 8 * nb + 2 * t + j;
                            if (r < col_lim) {
                                double *dst = dst_base + m + ldo * (int64_t)r;
                                if (m + 1 < row_lim) {
                                    *reinterpret_cast<double2 *>(dst) = make_double2(acc[2 * gq][nb][j], acc[2 * gq + 1][nb][j]);
                                } else {
                                    dst[0] = acc[2 * gq][nb][j];
                                }
                            }
                        }
                    }
                }
            }
        } else {
#pragma unroll
            for (int rbk = 0; rbk < 4; ++rbk) {
                const int n = 32 * warp + 8 * rbk + sigma_b(g);
                if (n < row_lim) {
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const int r = col0 + 8 * nb + 2 * t + j;
                            if (r < col_lim) dst_base[n + ldo * (int64_t)r] = acc[rbk][nb][j];
                        }
                    }
                }
            }
        }
    }
}

// Adds the partial tiles left by the stream-K split, in ascending-k order, and writes the finished tile.
__global__ void __launch_bounds__(256) streamk_fixup_kernel(const int *__restrict__ tile_ids, const int *__restrict__ slot_ptr,
                                                            const int *__restrict__ slot_list, const double *__restrict__ slots,
                                                            double *__restrict__ out, int64_t rows_out, int R, int num_rblocks,
                                                            int BM, int BN) {
    const int s = blockIdx.x;
    const int tile = tile_ids[s];
    const int row_tile = tile / num_rblocks;
    const int rb = tile - row_tile * num_rblocks;
    const int64_t row0 = (int64_t)row_tile * BM;
    const int p0 = slot_ptr[s], p1 = slot_ptr[s + 1];
    for (int e = blockIdx.y * 256 + threadIdx.x; e < BM * BN; e += 256 * gridDim.y) {
        const int lc = e / BM, lr = e - lc * BM;
        const int64_t row = row0 + lr;
        const int r = rb * BN + lc;
        if (row < rows_out && r < R) {
            double v = 0.0;
#pragma unroll 4
            for (int p = p0; p < p1; ++p) v += slots[(size_t)slot_list[p] * (size_t)(BM * BN) + e];   // ascending k: fixed order, loads in flight together
            out[row + rows_out * (int64_t)r] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Khatri-Rao operand packing: K[k, r] = prod_f A_f[i_f(k), r] written in DMMA B-fragment order
//   Kp[kt][rb][pair][nb][lane][e],  k = 16 kt + 8 pair + 2 (lane&3) + e,  r = 8 (rb NB + nb) + (lane>>2)
// zero for k >= K extent, padded rows of mode 0 and r >= R.
// ------------------------------------------------------------------------------------------------
struct PackArgs {
    const double *fac[ITCPD_MAX_ORDER];
    int64_t ext[ITCPD_MAX_ORDER];   // extent used to decode the linear k (ld0 for a padded mode 0)
    int64_t dim[ITCPD_MAX_ORDER];   // logical rows of the factor
    int nf;
    int64_t kext;                   // product of ext
    int R, NB, num_rblocks, kt_count;
};

// One thread per (k-tile, pair, lane): it owns the two contraction indices k0 = 16 kt + 8 pair + 2 t and k0 + 1,
// decodes their mode coordinates ONCE, then walks the rank columns r = 8 (rb NB + nb) + g and writes one 16-byte
// element per (rb, nb) -- a warp writes 512 contiguous bytes.  (The first version decoded every element separately:
// ~10 64-bit divisions per double made the pack kernel cost more than the contraction it feeds when the operand is
// a large Khatri-Rao product.)
__global__ void __launch_bounds__(256) pack_krp_kernel(PackArgs a, double *__restrict__ Kp, int64_t nthreads) {
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (tid >= nthreads) return;
    const int lane = (int)(tid & 31);
    const int pair = (int)((tid >> 5) & 1);
    const int64_t kt = tid >> 6;
    const int t = lane & 3, g = lane >> 2;
    const int64_t k0 = 16 * kt + 8 * pair + 2 * t;
    const double *p0[ITCPD_MAX_ORDER], *p1[ITCPD_MAX_ORDER];  // row pointers of every factor for k0 and k0 + 1 (null = zero)
    bool ok0 = k0 < a.kext, ok1 = k0 + 1 < a.kext;
    {
        int64_t r0 = k0, r1 = k0 + 1;
        for (int f = 0; f < a.nf; ++f) {
            const int64_t i0 = r0 % a.ext[f], i1 = r1 % a.ext[f];
            r0 /= a.ext[f];
            r1 /= a.ext[f];
            ok0 = ok0 && i0 < a.dim[f];
            ok1 = ok1 && i1 < a.dim[f];
            p0[f] = a.fac[f] + (i0 < a.dim[f] ? i0 : 0);
            p1[f] = a.fac[f] + (i1 < a.dim[f] ? i1 : 0);
        }
    }
    double2 *dst = reinterpret_cast<double2 *>(Kp) + ((kt * a.num_rblocks * 2 + pair) * (int64_t)a.NB) * 32 + lane;
    for (int rb = 0; rb < a.num_rblocks; ++rb) {
        for (int nb = 0; nb < a.NB; ++nb) {
            const int r = 8 * (rb * a.NB + nb) + g;
            double v0 = 0.0, v1 = 0.0;
            if (r < a.R) {
                v0 = ok0 ? 1.0 : 0.0;
                v1 = ok1 ? 1.0 : 0.0;
                for (int f = 0; f < a.nf; ++f) {
                    const int64_t off = a.dim[f] * (int64_t)r;
                    if (ok0) v0 *= p0[f][off];
                    if (ok1) v1 *= p1[f][off];
                }
            }
            dst[((int64_t)rb * 2 * a.NB + nb) * 32] = make_double2(v0, v1);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// kind-0 tile as ONE 3-D box: the M x K matrix viewed as (16, K, M/16) with strides (8, 8M, 128) bytes
static int make_tmap_3d(CUtensorMap *map, const double *base, uint64_t M, uint64_t K, uint32_t groups, bool swz) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return ITCPD_ERR_CUDA;
    cuuint64_t gdim[3] = {16, K, M / 16};
    cuuint64_t gstr[2] = {M * 8, 128};
    cuuint32_t box[3] = {16, (cuuint32_t)BK, groups};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? ITCPD_OK : ITCPD_ERR_CUDA;
}

static int make_tmap(CUtensorMap *map, const double *base, uint64_t d0, uint64_t d1, uint32_t box0, uint32_t box1, bool swz) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not found"); return ITCPD_ERR_CUDA; }
    cuuint64_t gdim[2] = {d0, d1};
    cuuint64_t gstr[1] = {d0 * 8};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d) dims=(%llu,%llu) box=(%u,%u)", (int)r, (unsigned long long)d0,
                  (unsigned long long)d1, box0, box1);
        return ITCPD_ERR_CUDA;
    }
    return ITCPD_OK;
}

template <int NB, int WARPS, int KIND>
static int launch_cfg(itcpd_ctx *c, const CUtensorMap &map, const double *Kp, double *out, int64_t rows_out, int R,
                      int num_row_tiles, int num_rblocks, int kt_count, int map3d) {
    using S = GemmSmem<NB, WARPS>;
    constexpr int BM = S::BM, BN = 8 * NB;
    auto kern = partial_gemm_kernel<NB, WARPS, KIND>;
    static bool attr_set[64] = {false};  // function attributes are per device
    if (!attr_set[c->device & 63]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
        attr_set[c->device & 63] = true;
    }
    const int per_sm = (WARPS <= 4) ? 2 : 1;
    const int64_t tiles = (int64_t)num_row_tiles * num_rblocks;
    int grid = (int)std::min<int64_t>(tiles * kt_count, (int64_t)c->sm_count * per_sm);
    if (grid < 1) grid = 1;
    // stream_k: 0 never, 1 adaptive (only when the last data-parallel wave would idle >= 4 % of the SM-time), 2 always
    const double waves_exact = (double)tiles / grid, waves_dp = (double)ceil_div(tiles, grid);
    const bool use_sk = c->stream_k == 2 || (c->stream_k == 1 && tiles % grid != 0 && 1.0 - waves_exact / waves_dp >= 0.04);
    const int full_waves = use_sk ? (int)(tiles / grid) : (int)ceil_div(tiles, grid);
    if (!use_sk) grid = (int)std::min<int64_t>(tiles, grid);
    const int64_t rem_tiles = std::max<int64_t>(0, tiles - (int64_t)full_waves * grid);
    const int64_t rem_units = rem_tiles * kt_count;

    // ---- stream-K bookkeeping (host): which tiles of the last round are split, and which slots hold their parts ----
    StreamKTable &tb = c->sk_table[KIND];
    const int64_t key[6] = {tiles, kt_count, grid, BM, BN, full_waves};
    if (memcmp(key, tb.key, sizeof(key)) != 0) {
        std::vector<int> tile_ids, slot_ptr(1, 0), slot_list;
        int cur_tile = -1;
        for (int cta = 0; cta < grid && rem_units > 0; ++cta) {
            const int64_t a0 = rem_units * cta / grid, a1 = rem_units * (cta + 1) / grid;
            int64_t u = a0;
            while (u < a1) {
                const int64_t q = u / kt_count, kb = u - q * kt_count;
                const int64_t ke = std::min<int64_t>(kt_count, kb + (a1 - u));
                const int tile = (int)((int64_t)full_waves * grid + q);
                if (!(kb == 0 && ke == kt_count)) {
                    if (tile != cur_tile) {
                        if (cur_tile >= 0) slot_ptr.push_back((int)slot_list.size());
                        tile_ids.push_back(tile);
                        cur_tile = tile;
                    }
                    slot_list.push_back(2 * cta + (kb == 0 ? 1 : 0));
                }
                u += ke - kb;
            }
        }
        if (cur_tile >= 0) slot_ptr.push_back((int)slot_list.size());
        tb.nsplit = (int)tile_ids.size();
        const size_t n1 = tile_ids.size(), n2 = slot_ptr.size(), n3 = slot_list.size();
        TRY(tb.dev.reserve((n1 + n2 + n3 + 4) * sizeof(int)));
        CUDA_TRY(cudaStreamSynchronize(c->stream));  // table rebuilds are rare (shape change); keep the copy simple
        if (tb.nsplit > 0) {
            CUDA_TRY(cudaMemcpy(tb.dev.as<int>(), tile_ids.data(), n1 * sizeof(int), cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(tb.dev.as<int>() + n1, slot_ptr.data(), n2 * sizeof(int), cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(tb.dev.as<int>() + n1 + n2, slot_list.data(), n3 * sizeof(int), cudaMemcpyHostToDevice));
        }
        tb.off_ptr = (int)n1;
        tb.off_list = (int)(n1 + n2);
        memcpy(tb.key, key, sizeof(key));
    }
    if (tb.nsplit > 0) TRY(c->sk_slots.reserve((size_t)2 * grid * BM * BN * 8));

    kern<<<grid, WARPS * 32, S::TOTAL, c->stream>>>(map, Kp, out, rows_out, R, num_row_tiles, num_rblocks, kt_count,
                                                c->swizzle ? 7 : 0, map3d, full_waves, c->sk_slots.as<double>());
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    if (tb.nsplit > 0) {
        // few split tiles with many parts each (the short-and-wide pass A of a slab: 8 tiles x 37 parts) need one element per thread to
        // cover the SMs; many split tiles (the last wave of pass B) are parallel enough at 4 elements per thread
        const int per_tile = tb.nsplit * 4 <= c->sm_count ? BM * BN / 256 : BM * BN / 1024;
        streamk_fixup_kernel<<<dim3((unsigned)tb.nsplit, (unsigned)std::max(1, std::min(per_tile, 16 * c->sm_count / tb.nsplit))), 256, 0, c->stream>>>(tb.dev.as<int>(), tb.dev.as<int>() + tb.off_ptr, tb.dev.as<int>() + tb.off_list,
                                                              c->sk_slots.as<double>(), out, rows_out, R, num_rblocks, BM, BN);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return ITCPD_OK;
}

template <int WARPS, int KIND>
static int launch_nb(itcpd_ctx *c, int NB, const CUtensorMap &map, const double *Kp, double *out, int64_t rows_out, int R,
                     int nrt, int nrb, int ktc, int map3d) {
    switch (NB) {
        case 1: return launch_cfg<1, WARPS, KIND>(c, map, Kp, out, rows_out, R, nrt, nrb, ktc, map3d);
        case 2: return launch_cfg<2, WARPS, KIND>(c, map, Kp, out, rows_out, R, nrt, nrb, ktc, map3d);
        case 3: return launch_cfg<3, WARPS, KIND>(c, map, Kp, out, rows_out, R, nrt, nrb, ktc, map3d);
        case 4: return launch_cfg<4, WARPS, KIND>(c, map, Kp, out, rows_out, R, nrt, nrb, ktc, map3d);
        case 5: return launch_cfg<5, WARPS, KIND>(c, map, Kp, out, rows_out, R, nrt, nrb, ktc, map3d);
        case 6: return launch_cfg<6, WARPS, KIND>(c, map, Kp, out, rows_out, R, nrt, nrb, ktc, map3d);
        case 7: return launch_cfg<7, WARPS, KIND>(c, map, Kp, out, rows_out, R, nrt, nrb, ktc, map3d);
        default: return launch_cfg<8, WARPS, KIND>(c, map, Kp, out, rows_out, R, nrt, nrb, ktc, map3d);
    }
}

int launch_partial_gemm(itcpd_ctx *c, int kind, int split, double *out) {
    const int N = c->order, R = c->rank;
    ARG_CHECK(split >= 1 && split < N, "bad dimension-tree split");
    // matrix view of the stored tensor: rows = modes [0,split) (mode 0 with leading dim ld0), cols = modes [split,N)
    int64_t Mrows = c->ld0, Ncols = 1;
    for (int n = 1; n < split; ++n) Mrows *= c->dims[n];
    for (int n = split; n < N; ++n) Ncols *= c->dims[n];

    // choose the r blocking: NB n-blocks of 8 columns per CTA pass, as balanced as possible
    const int nblk = (int)ceil_div(R, 8);
    const int num_rblocks = (int)ceil_div(nblk, 8);
    const int NB = (int)ceil_div(nblk, num_rblocks);

    const int64_t kext = (kind == 0) ? Ncols : Mrows;
    const int64_t rows_out = (kind == 0) ? Mrows : Ncols;
    const int kt_count = (int)ceil_div(kext, BK);

    // ---- pack the Khatri-Rao operand ----
    PackArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.R = R; pa.NB = NB; pa.num_rblocks = num_rblocks; pa.kt_count = kt_count; pa.kext = kext;
    if (kind == 0) {
        for (int n = split; n < N; ++n) {
            pa.fac[pa.nf] = c->A[n].as<double>(); pa.ext[pa.nf] = c->dims[n]; pa.dim[pa.nf] = c->dims[n]; pa.nf++;
        }
    } else {
        for (int n = 0; n < split; ++n) {
            pa.fac[pa.nf] = c->A[n].as<double>(); pa.ext[pa.nf] = (n == 0) ? c->ld0 : c->dims[n]; pa.dim[pa.nf] = c->dims[n]; pa.nf++;
        }
    }
    const int64_t total = (int64_t)kt_count * num_rblocks * 2 * NB * 64;
    TRY(c->packK.reserve((size_t)total * 8));
    {
        const int64_t nthreads = (int64_t)kt_count * 64;  // (k-tile, pair, lane)
        pack_krp_kernel<<<(unsigned)ceil_div(nthreads, 256), 256, 0, c->stream>>>(pa, c->packK.as<double>(), nthreads);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
    }

    // ---- tile shape ----
    int warps = c->tile_warps;
    if (warps != 4 && warps != 8) warps = 8;
    // small problems: prefer the 128-row tile so that more SMs get work
    if (ceil_div(rows_out, 256) * num_rblocks < c->sm_count) warps = 4;
    const int BM = 32 * warps;
    const int num_row_tiles = (int)ceil_div(rows_out, BM);

    CUtensorMap map;
    int map3d = 0;
    if (kind == 0) {
        if (c->tma3d && Mrows % 16 == 0 && make_tmap_3d(&map, c->T.as<double>(), (uint64_t)Mrows, (uint64_t)Ncols, BM / 16, c->swizzle != 0) == ITCPD_OK)
            map3d = 1;
        else
            TRY(make_tmap(&map, c->T.as<double>(), (uint64_t)Mrows, (uint64_t)Ncols, 16, BK, c->swizzle != 0));
    } else TRY(make_tmap(&map, c->T.as<double>(), (uint64_t)Mrows, (uint64_t)Ncols, BK, (uint32_t)BM, c->swizzle != 0));

    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->time_gemm) {
        if (c->gemm_events_used == c->gemm_events.size()) {
            cudaEvent_t a, b;
            CUDA_TRY(cudaEventCreate(&a));
            CUDA_TRY(cudaEventCreate(&b));
            c->gemm_events.push_back({a, b});
        }
        e0 = c->gemm_events[c->gemm_events_used].first;
        e1 = c->gemm_events[c->gemm_events_used].second;
        c->gemm_events_used++;
        CUDA_TRY(cudaEventRecord(e0, c->stream));
    }
    int st;
    const double *Kp = c->packK.as<double>();
    if (kind == 0) st = (warps == 8) ? launch_nb<8, 0>(c, NB, map, Kp, out, rows_out, R, num_row_tiles, num_rblocks, kt_count, map3d)
                                     : launch_nb<4, 0>(c, NB, map, Kp, out, rows_out, R, num_row_tiles, num_rblocks, kt_count, map3d);
    else st = (warps == 8) ? launch_nb<8, 1>(c, NB, map, Kp, out, rows_out, R, num_row_tiles, num_rblocks, kt_count, 0)
                           : launch_nb<4, 1>(c, NB, map, Kp, out, rows_out, R, num_row_tiles, num_rblocks, kt_count, 0);
    if (e1) CUDA_TRY(cudaEventRecord(e1, c->stream));
    return st;
}

// ------------------------------------------------------------------------------------------------
// FP64 peak probes (roofline denominators measured on the box the bench runs on)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) probe_dmma_kernel(double *sink, int iters) {
    double acc[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = 0.5, b0 = 1e-3, b1 = 2e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma884(acc[i][0], acc[i][1], (i & 1) ? a1 : a0, (i & 2) ? b1 : b0);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) sink[0] = s;
}

__global__ void __launch_bounds__(256) probe_dfma_kernel(double *sink, int iters) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-6 + i;
    const double a = 1.0000001, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 123.456) sink[0] = s;
}

static int run_probe(itcpd_ctx *c, bool dmma, double *tflops) {
    TRY(c->work.reserve(64));
    const int iters = 4096, blocks = c->sm_count * 4, threads = 256;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, c->stream));
        if (dmma) probe_dmma_kernel<<<blocks, threads, 0, c->stream>>>(c->work.as<double>(), iters);
        else probe_dfma_kernel<<<blocks, threads, 0, c->stream>>>(c->work.as<double>(), iters);
        c->launches++;
        CUDA_TRY(cudaEventRecord(e1, c->stream));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double warps = (double)blocks * threads / 32.0;
    const double flops = dmma ? warps * iters * 16.0 * (2.0 * 8 * 8 * 4) : (double)blocks * threads * iters * 16.0 * 2.0;
    *tflops = flops / (best * 1e-3) / 1e12;
    return ITCPD_OK;
}

int probe_dmma(itcpd_ctx *c, double *tflops) { return run_probe(c, true, tflops); }
int probe_dfma(itcpd_ctx *c, double *tflops) { return run_probe(c, false, tflops); }

}  // namespace itcpd
