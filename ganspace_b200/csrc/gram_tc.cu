// Small-side Gram  T = M M^T  of the large-d IPCA engine on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), with a
// PROMOTED accumulator.
//
// Part of the replacement of sklearn IncrementalPCA.partial_fit (_incremental_pca.py:254-380) for conv feature maps
// (bigd.cu): M is sklearn's stacked matrix [S*Vt; X - mean_b; correction] ([n_s <= 4096, d ~ 5e5] fp32 in HBM).
//
// Precision.  As in mapping_tc.cu every fp32 operand is split x 2^e = hi + lo into two fp16 numbers (22 significant
// bits; e = a per-row power-of-two exponent that puts the row maximum into [8192, 16384)), and a product is three MMAs
// (hi hi, lo hi, hi lo) accumulated in fp32 in TMEM.  TMEM accumulation TRUNCATES when aligning addends (measured on the
// mapping network: -1.6e-6 relative per 96 accumulation steps), which is harmless over K = 512 but not over K = 524288.
// So the MMA warp accumulates only FLUSH_KB * 64 = 256 values of K (48 MMA steps) into one of two TMEM accumulators;
// the epilogue warps then add that partial into fp32 REGISTERS with round-to-nearest (tcgen05.ld + FADD) while the MMA
// warp fills the other accumulator, and after a chunk of 8192 columns they scale by 2^-(e_i + e_j) and add into the fp64
// matrix T (RED.ADD.F64), mirrored across the diagonal.  Error budget per entry: <= 48 x 2^-24 truncation inside a flush,
// round-to-nearest fp32 across the 32 flushes of a chunk, fp64 across chunks.
//
// Kernel: persistent, one CTA per SM, 256 threads, 128 x 128 x 64 tiles (upper tile pairs only), 3 smem stages x 64 KB
// (A_hi, A_lo, B_hi, B_lo; both operands are row blocks of the same two fp16 matrices), SWIZZLE_128B K-major, 2 TMEM
// accumulators x 128 columns:
//     warp 0  TMA producer          warp 1  MMA issuer (tcgen05.mma.cta_group::1.kind::f16, M128 N128 K16)
//     warp 2  TMEM allocator        warps 4-7  promoted accumulation + fp64 reduction into T
// Work items are (chunk of d, tile pair) in chunk-major order, so the CTAs running at the same time read the same
// 2112 x 8192 slab of M (69 MB as hi+lo) out of the L2.
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>

namespace gsb {
namespace gtc {

constexpr int BM = 128, BN = 128, BK = 64, UK = 16;
constexpr int STAGES = 3, ACC = 2, THREADS = 256;
constexpr int FLUSH_KB = 4;                 // K-blocks (of 64) accumulated in TMEM before the promotion into registers
constexpr int CHUNK_KB = 128;               // K-blocks per work item (8192 columns of d)
constexpr uint32_t TILE_BYTES = BM * BK * 2;               // 16 KB
constexpr uint32_t STAGE_BYTES = 4 * TILE_BYTES;           // 64 KB
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "GT_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra GT_DONE;\n\t"
        "bra GT_WAIT_LOOP;\n\t"
        "GT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (same encoding as mapping_tc.cu)
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Params {
    double *T;
    const int *rexp;        // [n_pad] per-row scale exponents of the split operands
    int ldt, n_rows;
    int npairs, nt;         // upper tile pairs of the nt x nt grid of 128-row blocks
    int total_kb;           // d / 64
    int nchunks;
};

__device__ __forceinline__ void decode_pair(int pair, int nt, int &ti, int &tj) {
    ti = 0;
    while (pair >= nt - ti) { pair -= nt - ti; ++ti; }
    tj = ti + pair;
}

__global__ void __launch_bounds__(THREADS, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *full_bar = bars;                     // [STAGES]
    uint64_t *empty_bar = bars + STAGES;           // [STAGES]
    uint64_t *tfull_bar = bars + 2 * STAGES;       // [ACC]
    uint64_t *tempty_bar = tfull_bar + ACC;        // [ACC]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tempty_bar + ACC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_items = p.nchunks * p.npairs;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_hi); tma_prefetch_desc(&tm_lo); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < ACC; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int chunk = item / p.npairs;
                int ti, tj;
                decode_pair(item % p.npairs, p.nt, ti, tj);
                const int m0 = ti * BM, n0 = tj * BN;
                const int kb0 = chunk * CHUNK_KB, kb1 = (kb0 + CHUNK_KB < p.total_kb) ? kb0 + CHUNK_KB : p.total_kb;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t *st = smem + stage * STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                    tma_load_2d(&tm_hi, &full_bar[stage], st, kb * BK, m0);
                    tma_load_2d(&tm_lo, &full_bar[stage], st + TILE_BYTES, kb * BK, m0);
                    tma_load_2d(&tm_hi, &full_bar[stage], st + 2 * TILE_BYTES, kb * BK, n0);
                    tma_load_2d(&tm_lo, &full_bar[stage], st + 3 * TILE_BYTES, kb * BK, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(BM, BN);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int chunk = item / p.npairs;
                const int kb0 = chunk * CHUNK_KB, kb1 = (kb0 + CHUNK_KB < p.total_kb) ? kb0 + CHUNK_KB : p.total_kb;
                const int nkb = kb1 - kb0;
                for (int g0 = 0; g0 < nkb; g0 += FLUSH_KB) {
                    const int g1 = (g0 + FLUSH_KB < nkb) ? g0 + FLUSH_KB : nkb;
                    mbar_wait(&tempty_bar[acc], acc_phase ^ 1);       // the promotion of this accumulator's last partial is done
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                    for (int kbl = g0; kbl < g1; ++kbl) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t st = smem_u32(smem + stage * STAGE_BYTES);
                        const uint64_t d_ah = make_sw128_kmajor_desc(st);
                        const uint64_t d_al = make_sw128_kmajor_desc(st + TILE_BYTES);
                        const uint64_t d_bh = make_sw128_kmajor_desc(st + 2 * TILE_BYTES);
                        const uint64_t d_bl = make_sw128_kmajor_desc(st + 3 * TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k) {
                            const uint64_t koff = (uint64_t)((k * UK * 2) >> 4);
                            tc_mma_f16(tmem_d, d_ah + koff, d_bh + koff, idesc, (kbl > g0 || k > 0) ? 1u : 0u);
                            tc_mma_f16(tmem_d, d_al + koff, d_bh + koff, idesc, 1u);
                            tc_mma_f16(tmem_d, d_ah + koff, d_bl + koff, idesc, 1u);
                        }
                        tc_commit(&empty_bar[stage]);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(&tfull_bar[acc]);
                    if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== promoted accumulation + reduction into T =====================
        const int ew = warp - 4;                        // == warp % 4: TMEM lane quadrant
        const int row_in_tile = ew * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        float r[BN];
#pragma unroll
        for (int j = 0; j < BN; ++j) r[j] = 0.f;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int chunk = item / p.npairs;
            int ti, tj;
            decode_pair(item % p.npairs, p.nt, ti, tj);
            const int m0 = ti * BM, n0 = tj * BN;
            const int kb0 = chunk * CHUNK_KB, kb1 = (kb0 + CHUNK_KB < p.total_kb) ? kb0 + CHUNK_KB : p.total_kb;
            const int nkb = kb1 - kb0;
            for (int g0 = 0; g0 < nkb; g0 += FLUSH_KB) {
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tc_ld32(taddr + (uint32_t)c0, v);
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[c0 + j] = __fadd_rn(r[c0 + j], __uint_as_float(v[j]));
                }
                tc_fence_before();
                mbar_arrive(&tempty_bar[acc]);
                if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
            }
            const int gm = m0 + row_in_tile;
            if (gm < p.n_rows) {
                const int em = p.rexp[gm];
                double *Trow = p.T + (size_t)gm * p.ldt;
#pragma unroll
                for (int j = 0; j < BN; ++j) {
                    const int gn = n0 + j;
                    if (gn < p.n_rows && gn >= gm) {
                        const double val = ldexp((double)r[j], -(em + __ldg(&p.rexp[gn])));
                        atomicAdd(&Trow[gn], val);
                        if (gn > gm) atomicAdd(&p.T[(size_t)gn * p.ldt + gm], val);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < BN; ++j) r[j] = 0.f;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
    }
}

// ---- operand preparation ----------------------------------------------------------------------------------
// per-row exponent e: the row maximum times 2^e lies in [8192, 16384) (hi keeps 11 bits, lo stays a normal fp16)
__global__ void __launch_bounds__(1024)
row_exponent_kernel(const float *__restrict__ M, int64_t d, int n_rows, int *__restrict__ rexp) {
    __shared__ float red[32];
    const int r = blockIdx.x, tid = threadIdx.x;
    float m = 0.f;
    if (r < n_rows) {
        const float4 *row = reinterpret_cast<const float4 *>(M + (size_t)r * d);
        for (int64_t i = tid; i < d / 4; i += 1024) {
            const float4 v = row[i];
            m = fmaxf(fmaxf(m, fabsf(v.x)), fmaxf(fabsf(v.y), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) red[tid >> 5] = m;
    __syncthreads();
    if (tid < 32) {
        m = red[tid];
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (tid == 0) {
            int e = 0;
            if (m > 0.f && m < 3.0e38f) { int ex; frexpf(m, &ex); e = 14 - ex; }
            rexp[r] = e;
        }
    }
}
__global__ void __launch_bounds__(256)
row_split_kernel(const float *__restrict__ M, int64_t d, int n_rows, const int *__restrict__ rexp, __half *__restrict__ hi,
                 __half *__restrict__ lo) {
    const int r = blockIdx.y;
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;           // float4 index within the row
    if (q >= d / 4) return;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < n_rows) v = reinterpret_cast<const float4 *>(M + (size_t)r * d)[q];
    const float sc = ldexpf(1.f, rexp[r]);
    const float f[4] = {v.x * sc, v.y * sc, v.z * sc, v.w * sc};              // power-of-two scale: exact
    __half h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        h[k] = __float2half_rn(f[k]);
        l[k] = __float2half_rn(f[k] - __half2float(h[k]));
    }
    uint2 ph, pl;
    ph.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
    ph.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
    pl.x = (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16);
    pl.y = (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16);
    reinterpret_cast<uint2 *>(hi + (size_t)r * d)[q] = ph;
    reinterpret_cast<uint2 *>(lo + (size_t)r * d)[q] = pl;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}
static int make_tmap(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return GSB_ERR_CUDA; }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return GSB_ERR_CUDA; }
    return GSB_OK;
}

}  // namespace gtc

size_t gram_tc_workspace_bytes(int n_pad, int64_t d) {
    return 2 * align_up((size_t)n_pad * d * 2, 256) + align_up((size_t)n_pad * sizeof(int), 256);
}
bool gram_tc_supported(int64_t d) { return d % 64 == 0; }

// T[n_pad, n_pad] (fp64, zeroed by the caller) += M[0:n_rows] M[0:n_rows]^T.  ws: gram_tc_workspace_bytes(n_pad, d).
int gram_tc(const float *M, int n_rows, int n_pad, int64_t d, void *ws, double *T, cudaStream_t st) {
    using namespace gtc;
    GSB_CHECK_ARG(gram_tc_supported(d) && n_rows <= n_pad, "gram_tc: d %% 64 != 0");
    const size_t hb = align_up((size_t)n_pad * d * 2, 256);
    __half *hi = reinterpret_cast<__half *>(ws);
    __half *lo = reinterpret_cast<__half *>(reinterpret_cast<char *>(ws) + hb);
    int *rexp = reinterpret_cast<int *>(reinterpret_cast<char *>(ws) + 2 * hb);
    static bool attr_set = false;
    if (!attr_set) {
        GSB_CHECK_CUDA(cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        attr_set = true;
    }
    row_exponent_kernel<<<n_pad, 1024, 0, st>>>(M, d, n_rows, rexp);
    GSB_CHECK_LAUNCH();
    dim3 sg((unsigned)((d / 4 + 255) / 256), (unsigned)n_pad);
    row_split_kernel<<<sg, 256, 0, st>>>(M, d, n_rows, rexp, hi, lo);
    GSB_CHECK_LAUNCH();
    CUtensorMap tm_hi, tm_lo;
    if (int r = make_tmap(&tm_hi, hi, (uint64_t)n_pad, (uint64_t)d)) return r;
    if (int r = make_tmap(&tm_lo, lo, (uint64_t)n_pad, (uint64_t)d)) return r;
    Params p;
    p.T = T; p.rexp = rexp; p.ldt = n_pad; p.n_rows = n_rows;
    p.nt = (n_rows + BM - 1) / BM;
    p.npairs = p.nt * (p.nt + 1) / 2;
    p.total_kb = (int)(d / BK);
    p.nchunks = (p.total_kb + CHUNK_KB - 1) / CHUNK_KB;
    const int items = p.nchunks * p.npairs;
    const int grid = items < num_sms() ? items : num_sms();
    gram_tc_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(tm_hi, tm_lo, p);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

}  // namespace gsb
