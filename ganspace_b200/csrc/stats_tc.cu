// Per-group sufficient statistics of the small-d incremental PCA on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
//   mean_g[d]   = (1/nb) sum_r x[g nb + r, :]                          fp64 accumulation of the fp32 samples
//   gram_g[d,d] = sum_r (x[r,:] - mean_g)^T (x[r,:] - mean_g)          fp32-grade products, promoted accumulation
// for G consecutive groups of nb rows in ONE call.  These replace what sklearn's IncrementalPCA.partial_fit reads off a
// batch (estimators.py:68-76 -> _incremental_pca.py:332-357), exactly like stats.cu (the fp32 FMA form, kept for feature
// counts that are not a multiple of 128); the chain (subspace.cu / ipca.cu) consumes them in the reference's group order.
//
// Three kernels per call:
//   colsum_groups_kernel     column sums per group (fp64), coalesced reads                       -> mean, mean32
//   center_split_t_kernel    xt[g][c][r] = fp16 hi/lo split of (x[g nb + r][c] - mean32[c]) * 2^e_g (e_g: per-group power of two
//                            that puts 2 max|x| into [8192, 32768), so hi keeps 11 bits and nothing overflows), i.e. the centred samples
//                            TRANSPOSED so that the sample index is the contiguous (K) dimension of both MMA operands
//   gram_groups_tc_kernel    persistent, one CTA per SM, work items (group, upper tile pair) of 128 x 128 outputs:
//                              warp 0  TMA producer   3-D boxes {64 samples, 128 features, 1 group}, SWIZZLE_128B; samples
//                                                     beyond nb are zero-filled by the TMA unit (no padding is ever read)
//                              warp 1  MMA issuer     tcgen05.mma.cta_group::1.kind::f16 M128 N128 K16; 3 MMAs per product
//                                                     (hi hi, lo hi, hi lo), fp32 accumulation in TMEM
//                              warp 2  TMEM allocator (2 accumulators x 128 columns)
//                              warps 4-7 epilogue     TMEM accumulation truncates when aligning addends, so only 4 K-blocks
//                                                     (48 MMA steps) are summed in TMEM; the epilogue adds each partial into
//                                                     fp32 registers with round-to-nearest (as gram_tc.cu does) while the MMA
//                                                     warp fills the other accumulator, and finally scales by 2^-2e and
//                                                     stores the tile (and its mirror image) as fp64.
// Algorithmic work: 2 d^2 FLOP per sample (x3 MMAs); HBM traffic: x read twice (4 B), xt written and read (4 + 4 B per element,
// the re-reads of the 10 tile pairs come out of L2).
#include "tc_common.cuh"
#include <stdlib.h>
#include <string.h>

namespace gsb {

// ---- host: tensor maps (shared by the tensor-core kernels) ----------------------------------------------------------
TcEncodeTiledFn tc_encode_fn() {
    static TcEncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TcEncodeTiledFn>(ptr);
    }
    return fn;
}

int tc_make_tmap_f16(CUtensorMap *map, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                     const uint32_t *box) {
    TcEncodeTiledFn enc = tc_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return GSB_ERR_CUDA; }
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t bx[5], estr[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; estr[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstride[i] = strides_bytes[i];
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstride, bx, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return GSB_ERR_CUDA; }
    return GSB_OK;
}

namespace stc {

constexpr int BM = 128, BN = 128, BK = 64, UK = 16;
constexpr int STAGES = 3, ACC = 2, THREADS = 256;
constexpr int FLUSH_KB = 4;                 // K-blocks (of 64 samples) accumulated in TMEM before the promotion into registers
constexpr uint32_t TILE_BYTES = BM * BK * 2;               // 16 KB
constexpr uint32_t STAGE_BYTES = 4 * TILE_BYTES;           // 64 KB  (A_hi, A_lo, B_hi, B_lo)
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;

// ---- column sums per group ----------------------------------------------------------------------------------------
__global__ void colsum_groups_kernel(const float *__restrict__ x, int64_t nb, int d, int64_t ld, int rows_per_cta,
                                     double *__restrict__ sum /* [G][d], zeroed */, float *__restrict__ absmax /* [G], zeroed */) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.z;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
    const int64_t r1 = r0 + rows_per_cta < nb ? r0 + rows_per_cta : nb;
    const float *xg = x + (size_t)g * nb * ld;
    double a0 = 0.0, a1 = 0.0;              // fp64 sum of the fp32 samples, as sklearn's _safe_accumulator_op does
    float mx = 0.f;
    if (col < d) {
        int64_t r = r0;
        for (; r + 1 < r1; r += 2) {
            const float v0 = xg[r * ld + col], v1 = xg[(r + 1) * ld + col];
            a0 += (double)v0; a1 += (double)v1;
            mx = fmaxf(mx, fmaxf(fabsf(v0), fabsf(v1)));
        }
        if (r < r1) { const float v0 = xg[r * ld + col]; a0 += (double)v0; mx = fmaxf(mx, fabsf(v0)); }
        atomicAdd(&sum[(size_t)g * d + col], a0 + a1);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int *>(absmax + g), __float_as_int(mx));   // mx >= 0: int order == float order
}

// mean per (group, feature); scale exponent per group: |x - mean| <= 2 max|x|, and 2 max|x| 2^e lands in [8192, 32768)
__global__ void mean_groups_kernel(const double *__restrict__ sum, int n_groups, int d, double nd, double *__restrict__ mean,
                                   float *__restrict__ mean32, const float *__restrict__ absmax, int *__restrict__ exps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_groups * d) {
        const double m = sum[i] / nd;
        mean[i] = m;
        mean32[i] = (float)m;
    }
    if (i < n_groups) {
        const float a = absmax[i];
        int e = 0;
        if (a > 0.f && a < 1.0e38f) { int ex; frexpf(2.f * a, &ex); e = 14 - ex; }
        exps[i] = e;
    }
}

// ---- centre, scale, split, transpose ---------------------------------------------------------------------------------
// tile: 64 samples x 64 features; 256 threads.  xt_hi / xt_lo: [G][d][nbp] fp16 (nbp = nb rounded up to 64).
__global__ void __launch_bounds__(256)
center_split_t_kernel(const float *__restrict__ x, int64_t nb, int d, int64_t ld, int64_t nbp, const float *__restrict__ mean32,
                      const int *__restrict__ exps, __half *__restrict__ xt_hi, __half *__restrict__ xt_lo) {
    __shared__ float tile[64][65];
    const int g = blockIdx.z;
    const int64_t r0 = (int64_t)blockIdx.x * 64;
    const int c0 = blockIdx.y * 64;
    const int tid = threadIdx.x;
    const float *xg = x + (size_t)g * nb * ld;
    const float *mg = mean32 + (size_t)g * d;
    {
        const int c4 = (tid & 15) * 4, rr = tid >> 4;
        const float4 m = *reinterpret_cast<const float4 *>(mg + c0 + c4);
        const float sc = ldexpf(1.f, exps[g]);                  // power of two: exact
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int r = rr + 16 * p;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r0 + r < nb) {
                v = *reinterpret_cast<const float4 *>(xg + (r0 + r) * ld + c0 + c4);
                v.x = (v.x - m.x) * sc; v.y = (v.y - m.y) * sc; v.z = (v.z - m.z) * sc; v.w = (v.w - m.w) * sc;
            }
            tile[r][c4] = v.x; tile[r][c4 + 1] = v.y; tile[r][c4 + 2] = v.z; tile[r][c4 + 3] = v.w;
        }
    }
    __syncthreads();
    {
        const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int c = warp * 8 + q;
            uint32_t hi, lo;
            tc::split2(tile[2 * lane][c], tile[2 * lane + 1][c], hi, lo);
            const size_t o = ((size_t)g * d + c0 + c) * nbp + r0 + 2 * lane;
            *reinterpret_cast<uint32_t *>(xt_hi + o) = hi;
            *reinterpret_cast<uint32_t *>(xt_lo + o) = lo;
        }
    }
}

// ---- Gram of every group ---------------------------------------------------------------------------------------------
struct Params {
    double *gram;           // [G][d][d]
    const int *exps;        // [G] scale exponents of the split operands
    int d, nt, npairs;      // nt = d / 128 row blocks, npairs = nt (nt + 1) / 2
    int n_groups, nkb;      // nkb = ceil(nb / 64)
};

__device__ __forceinline__ void decode_pair(int pair, int nt, int &ti, int &tj) {
    ti = 0;
    while (pair >= nt - ti) { pair -= nt - ti; ++ti; }
    tj = ti + pair;
}

__global__ void __launch_bounds__(THREADS, 1)
gram_groups_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const Params p) {
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *full_bar = bars;                     // [STAGES]
    uint64_t *empty_bar = bars + STAGES;           // [STAGES]
    uint64_t *tfull_bar = bars + 2 * STAGES;       // [ACC]
    uint64_t *tempty_bar = tfull_bar + ACC;        // [ACC]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tempty_bar + ACC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_items = p.n_groups * p.npairs;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_hi); tma_prefetch_desc(&tm_lo); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < ACC; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 128); }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 256);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int g = item / p.npairs;
                int ti, tj;
                decode_pair(item % p.npairs, p.nt, ti, tj);
                const bool diag = (ti == tj);
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t *st = smem + stage * STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], diag ? 2 * TILE_BYTES : 4 * TILE_BYTES);
                    tma_load_3d(&tm_hi, &full_bar[stage], st, kb * BK, ti * BM, g);
                    tma_load_3d(&tm_lo, &full_bar[stage], st + TILE_BYTES, kb * BK, ti * BM, g);
                    if (!diag) {
                        tma_load_3d(&tm_hi, &full_bar[stage], st + 2 * TILE_BYTES, kb * BK, tj * BN, g);
                        tma_load_3d(&tm_lo, &full_bar[stage], st + 3 * TILE_BYTES, kb * BK, tj * BN, g);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_f16(BM, BN);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                int ti, tj;
                decode_pair(item % p.npairs, p.nt, ti, tj);
                const uint32_t boff = (ti == tj) ? 0u : 2 * TILE_BYTES;      // diagonal tile: B is A
                for (int g0 = 0; g0 < p.nkb; g0 += FLUSH_KB) {
                    const int g1 = (g0 + FLUSH_KB < p.nkb) ? g0 + FLUSH_KB : p.nkb;
                    mbar_wait(&tempty_bar[acc], acc_phase ^ 1);               // this accumulator's last partial has been promoted
                    fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                    for (int kb = g0; kb < g1; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        fence_after();
                        const uint32_t st = smem_u32(smem + stage * STAGE_BYTES);
                        const uint64_t d_ah = sw128_kmajor_desc(st), d_al = sw128_kmajor_desc(st + TILE_BYTES);
                        const uint64_t d_bh = sw128_kmajor_desc(st + boff), d_bl = sw128_kmajor_desc(st + boff + TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k) {
                            const uint64_t koff = (uint64_t)((k * UK * 2) >> 4);
                            mma_f16(tmem_d, d_ah + koff, d_bh + koff, idesc, (kb > g0 || k > 0) ? 1u : 0u);
                            mma_f16(tmem_d, d_al + koff, d_bh + koff, idesc, 1u);
                            mma_f16(tmem_d, d_ah + koff, d_bl + koff, idesc, 1u);
                        }
                        commit(&empty_bar[stage]);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    commit(&tfull_bar[acc]);
                    if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== promoted accumulation + store =====================
        const int ew = warp - 4;                        // == warp % 4: TMEM lane quadrant
        const int row_in_tile = ew * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        float r[BN];
#pragma unroll
        for (int j = 0; j < BN; ++j) r[j] = 0.f;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int g = item / p.npairs;
            const double unscale = ldexp(1.0, -2 * __ldg(p.exps + g));
            int ti, tj;
            decode_pair(item % p.npairs, p.nt, ti, tj);
            for (int g0 = 0; g0 < p.nkb; g0 += FLUSH_KB) {
                mbar_wait(&tfull_bar[acc], acc_phase);
                fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    ld32(taddr + (uint32_t)c0, v);
                    wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[c0 + j] = __fadd_rn(r[c0 + j], __uint_as_float(v[j]));
                }
                fence_before();
                mbar_arrive(&tempty_bar[acc]);
                if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
            }
            double *G = p.gram + (size_t)g * p.d * p.d;
            const int gm = ti * BM + row_in_tile, n0 = tj * BN;
            double *row = G + (size_t)gm * p.d + n0;
            double *col = G + (size_t)n0 * p.d + gm;                      // mirror image: coalesced over the lanes
            const size_t cstep = (size_t)p.d;
            if (ti != tj) {
#pragma unroll
                for (int j = 0; j < BN; j += 2) {
                    const double v0 = (double)r[j] * unscale, v1 = (double)r[j + 1] * unscale;
                    *reinterpret_cast<double2 *>(row) = make_double2(v0, v1);
                    col[0] = v0; col[cstep] = v1;
                    row += 2; col += 2 * cstep;
                    asm volatile("" : "+l"(row), "+l"(col));              // running pointers: no 128 precomputed addresses
                }
            } else {
                // diagonal tile: (i,j) and (j,i) come out of differently ordered MMA sums; keep the upper triangle and mirror
                // it, so the chain reads an exactly symmetric matrix
#pragma unroll
                for (int j = 0; j < BN; ++j) {
                    const double v0 = (double)r[j] * unscale;
                    if (j >= row_in_tile) { row[0] = v0; col[0] = v0; }
                    row += 1; col += cstep;
                    asm volatile("" : "+l"(row), "+l"(col));
                }
            }
#pragma unroll
            for (int j = 0; j < BN; ++j) r[j] = 0.f;
        }
    }

    fence_before();
    __syncthreads();
    if (warp == 2) {
        fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

struct WsView {
    double *sum;
    float *mean32;
    __half *xt_hi, *xt_lo;
    float *absmax;
    int *exps;
    size_t bytes;
};
static WsView carve(void *base, int n_groups, int64_t nb, int d) {
    WsView w;
    char *p = reinterpret_cast<char *>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { char *q = p + off; off += align_up(bytes, 1024); return q; };
    const int64_t nbp = (nb + 63) / 64 * 64;
    w.sum = (double *)take((size_t)n_groups * d * sizeof(double) + (size_t)n_groups * sizeof(float));   // sums, then absmax
    w.absmax = reinterpret_cast<float *>(w.sum + (size_t)n_groups * d);
    w.exps = (int *)take((size_t)n_groups * sizeof(int));
    w.mean32 = (float *)take((size_t)n_groups * d * sizeof(float));
    w.xt_hi = (__half *)take((size_t)n_groups * d * nbp * 2);
    w.xt_lo = (__half *)take((size_t)n_groups * d * nbp * 2);
    w.bytes = off;
    return w;
}

}  // namespace stc

bool stats_tc_supported(int64_t nb, int d) {
    static int mode = -1;
    if (mode == -1) {
        const char *env = getenv("GANSPACE_B200_STATS");
        mode = (env && strcmp(env, "simt") == 0) ? 0 : 1;
    }
    return mode == 1 && d % 128 == 0 && d >= 128 && d <= 1024 && nb >= 1;
}

size_t stats_tc_workspace_bytes(int n_groups, int64_t nb, int d) { return stc::carve(nullptr, n_groups, nb, d).bytes; }

// mean[G][d], gram[G][d][d] of G consecutive groups of nb rows of x (row stride ld)
int stats_tc(const float *x, int n_groups, int64_t nb, int d, int64_t ld, double *mean, double *gram, void *ws, cudaStream_t st) {
    using namespace stc;
    WsView w = carve(ws, n_groups, nb, d);
    const int64_t nbp = (nb + 63) / 64 * 64;
    static bool attr_set = false;
    if (!attr_set) {
        GSB_CHECK_CUDA(cudaFuncSetAttribute(gram_groups_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        attr_set = true;
    }
    GSB_CHECK_CUDA(cudaMemsetAsync(w.sum, 0, (size_t)n_groups * d * sizeof(double) + (size_t)n_groups * sizeof(float), st));
    {
        const int rows = 256;
        dim3 grid((d + 127) / 128, (unsigned)((nb + rows - 1) / rows), (unsigned)n_groups);
        colsum_groups_kernel<<<grid, 128, 0, st>>>(x, nb, d, ld, rows, w.sum, w.absmax);
        GSB_CHECK_LAUNCH();
        mean_groups_kernel<<<(n_groups * d + 255) / 256, 256, 0, st>>>(w.sum, n_groups, d, (double)nb, mean, w.mean32, w.absmax, w.exps);
        GSB_CHECK_LAUNCH();
    }
    {
        dim3 grid((unsigned)(nbp / 64), (unsigned)(d / 64), (unsigned)n_groups);
        center_split_t_kernel<<<grid, 256, 0, st>>>(x, nb, d, ld, nbp, w.mean32, w.exps, w.xt_hi, w.xt_lo);
        GSB_CHECK_LAUNCH();
    }
    CUtensorMap tm_hi, tm_lo;
    const uint64_t dims[3] = {(uint64_t)nb, (uint64_t)d, (uint64_t)n_groups};
    const uint64_t strides[2] = {(uint64_t)nbp * 2, (uint64_t)d * nbp * 2};
    const uint32_t box[3] = {(uint32_t)BK, (uint32_t)BM, 1};
    if (int r = tc_make_tmap_f16(&tm_hi, w.xt_hi, 3, dims, strides, box)) return r;
    if (int r = tc_make_tmap_f16(&tm_lo, w.xt_lo, 3, dims, strides, box)) return r;
    Params p;
    p.gram = gram; p.exps = w.exps; p.d = d; p.nt = d / BM; p.npairs = p.nt * (p.nt + 1) / 2;
    p.n_groups = n_groups; p.nkb = (int)(nbp / BK);
    const int items = p.n_groups * p.npairs;
    const int grid = items < num_sms() ? items : num_sms();
    gram_groups_tc_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(tm_hi, tm_lo, p);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

}  // namespace gsb
