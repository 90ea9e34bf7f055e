// StyleGAN2 mapping network on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), fp32-grade accuracy.
//
// Replaces models/stylegan2/stylegan2-pytorch/model.py:151-161 (EqualLinear.forward: F.linear(x, W*scale)
// then fused bias + leaky-ReLU(0.2)*sqrt2, op/fused_act.py:86-92) for the 8 layers of Generator.style.
//
// Precision.  The parity target (principal directions vs sklearn on the same seeds, cos >= 0.999 at relative
// eigen-gaps of 0.27 %) needs ~1e-6 relative accuracy per layer: single-pass bf16/fp16 (2^-9 / 2^-12) fails
// it (SURVEY.md section 7 hard part 3).  Each fp32 operand is split into two fp16 numbers, x = hi + lo with
// hi = fp16(x), lo = fp16(x - hi)  (22 significant bits), and the product is formed from three MMAs
//     D += A_hi W_hi ;  D += A_lo W_hi ;  D += A_hi W_lo          (fp32 accumulation in TMEM)
// dropping only the lo*lo term (2^-24).  Weights are pre-scaled by a per-layer power of two so that W_lo
// stays in fp16's normal range; the scale is undone exactly in the epilogue.  fp16 (11-bit significand) at
// the bf16 MMA rate gives this accuracy in 3 MMAs; TF32 would need 3 MMAs at half the rate.
//
// Kernel (one launch per layer; activations travel between layers as fp16 hi/lo pairs = the same 4 B/element
// as fp32): persistent, one CTA per SM, warp-specialised:
//     warp 0   TMA producer   (A_hi, A_lo: 128 x 64 boxes; W_hi, W_lo: 256 x 64 boxes; SWIZZLE_128B, K-major)
//     warp 1   MMA issuer     (one thread: tcgen05.mma.cta_group::1.kind::f16, M128 N256 K16, 12 per K-block)
//     warp 2   TMEM allocator (512 columns = two 128x256 fp32 accumulators, so the epilogue of tile i
//                              overlaps the MMAs of tile i+1)
//     warps 4-7 epilogue      (tcgen05.ld 32x32b -> 2^-s, +bias, leaky-ReLU, *sqrt2 -> split to fp16 hi/lo
//                              for the next layer, or fp32 for the last layer)
// smem ring: 2 stages x 96 KB (A_hi 16K, A_lo 16K, W_hi 32K, W_lo 32K), mbarrier full/empty pairs.
//
// Round 2: the kernel was L2-bandwidth bound (768 KB of operands per 128 x 256 x 512 tile, two thirds of it weights: 12 GB per
// layer over 1.01M rows against ~12 TB/s of L2 -> tensor pipe 52 %).  CTAs now run in thread-block clusters of TC_CLUSTER (4)
// along M: the four CTAs work on four different row tiles and the same weight tile, each loads a quarter of the weight box and
// TMA-multicasts it into all four shared memories (weight traffic / 4: 6 GB per layer), a stage is recycled when the MMAs of
// all four CTAs have retired (tcgen05.commit multicast onto every CTA's empty barrier).  A CTA walks through all N tiles of its
// row tile before moving on, so the second pass over its A rows hits L2.
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace gsb {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_N = 256;
constexpr int TC_BLOCK_K = 64;           // 64 fp16 = one 128-byte swizzle row
constexpr int TC_UMMA_K = 16;
constexpr int TC_STAGES = 2;
constexpr int TC_ACC_STAGES = 2;
constexpr int TC_THREADS = 256;
constexpr uint32_t TC_A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;   // 16 KB
constexpr uint32_t TC_W_BYTES = TC_BLOCK_N * TC_BLOCK_K * 2;   // 32 KB
constexpr uint32_t TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_W_BYTES;   // 96 KB
constexpr uint32_t TC_STAGING_BYTES = 2 * TC_A_BYTES + 1024;   // epilogue: two 16 KB output boxes + 256 bias floats
constexpr uint32_t TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align slack*/ + 1024 /*barriers*/ + TC_STAGING_BYTES;

// ---- PTX wrappers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
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
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_commit2_mc(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_mma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of this cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
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

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (8 rows * 128 B = 1024)
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) @4, a/b_format F16 (0) @7/@10,
// a/b K-major (0) @15/@16, N>>3 @17, M>>4 @24.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct TcParams {
    const float *bias;      // [N_total] pre-multiplied by lr_mul
    __half *out_hi;         // next layer's A_hi [M, N_total] (nullptr on the last layer)
    __half *out_lo;
    float *out_f32;         // fp32 output [M, N_total] (last layer)
    unsigned *overflow;     // set to 1 when an activation leaves fp16's range
    const float *inv_wscale;   // device pointer to 2^-s of this layer
    int M, N_total, K;
    int mode;               // 0: (acc 2^-s + bias) -> leaky-ReLU * sqrt2 (EqualLinear);  1: plain acc 2^-s (tc_gemm_plain);  2: acc 2^-s + bias
    int n_groups;           // work units per cluster tile (1 or N_total / 256)
    int pair;               // 1: CTA pairs (cta_group::2): one 256 x 256 MMA tile per pair, each CTA holds half of the weight tile
    int dbg;                // profiling experiments (GANSPACE_B200_MAPPING_DBG): 1 no stores, 2 no W loads, 4 no A loads, 8 no MMAs
};

template <bool PAIR>
__global__ void __launch_bounds__(TC_THREADS, 1)
mapping_layer_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                        const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                        const __grid_constant__ CUtensorMap tm_o0, const __grid_constant__ CUtensorMap tm_o1,
                        const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // pair mode (cta_group::2): each CTA stages its own 128 A rows and HALF of the weight tile (128 of the 256 N rows): 64 KB per
    // stage, three stages; otherwise 96 KB per stage, two stages -- the same 192 KB ring either way
    // (PAIR is a template parameter: a kernel that contains cta_group::2 instructions can only be launched with an even cluster
    // size -- "cluster misconfiguration" otherwise -- so the ordinary instantiation must not contain them)
    constexpr bool pair = PAIR;
    const int n_stages = pair ? 3 : TC_STAGES;
    const uint32_t w_tile_bytes = pair ? TC_W_BYTES / 2 : TC_W_BYTES;
    const uint32_t stage_bytes = 2 * TC_A_BYTES + 2 * w_tile_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t *full_bar = bars;                         // [3]
    uint64_t *empty_bar = bars + 3;                    // [3]
    uint64_t *tfull_bar = bars + 6;                    // [TC_ACC_STAGES]
    uint64_t *tempty_bar = tfull_bar + TC_ACC_STAGES;  // [TC_ACC_STAGES]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tempty_bar + TC_ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m_tiles = (p.M + TC_BLOCK_M - 1) / TC_BLOCK_M;
    // N and K need not be multiples of the tile: the TMA unit zero-fills operand rows / columns past the tensor's extent and
    // clips stores (the 128/64/32-channel StyledConv blocks: N = 9 cout = 1152 / 576 / 288, K = cin down to 32)
    const int num_n_tiles = (p.N_total + TC_BLOCK_N - 1) / TC_BLOCK_N;
    const int num_k_blocks = (p.K + TC_BLOCK_K - 1) / TC_BLOCK_K;
    // cluster of cs CTAs along M (cs = 1: no cluster): CTA `crank` of cluster `cluster_id` owns row tile ct * cs + crank of every
    // cluster tile ct it visits; all CTAs of a cluster run the same number of pipeline steps (row tiles past the end are
    // zero-filled by the TMA unit and never stored)
    const uint32_t cs = cluster_nctarank(), crank = cluster_ctarank();
    const int cluster_id = blockIdx.x / cs, num_clusters = gridDim.x / cs;
    const int num_ct = (num_m_tiles + (int)cs - 1) / (int)cs;
    // work unit = (cluster tile, group of consecutive N tiles); n_groups = 1: a CTA walks through all N tiles of its row tile
    // (its A rows are re-read out of L2), n_groups = num_n_tiles: one output tile per unit (few row tiles: more parallelism)
    const int n_groups = p.n_groups, tiles_per_group = num_n_tiles / n_groups;
    const int num_units = num_ct * n_groups;
    const uint16_t mc_mask = (uint16_t)((1u << cs) - 1u);
    const uint32_t w_slice_rows = TC_BLOCK_N / cs, w_slice_bytes = TC_W_BYTES / cs;
    const bool leader = (crank == 0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a_hi); tma_prefetch_desc(&tm_a_lo);
        tma_prefetch_desc(&tm_w_hi); tma_prefetch_desc(&tm_w_lo);
        tma_prefetch_desc(&tm_o0); tma_prefetch_desc(&tm_o1);
    }
    if (warp == 1 && lane == 0) {
        // pair mode: the leader's full barrier also takes the peer's "my half has landed" arrival; a stage is released by ONE
        // commit of the leader (multicast to both CTAs); the leader's accumulator-empty barrier takes both CTAs' epilogue threads
        for (int s = 0; s < 3; ++s) { mbar_init(&full_bar[s], (pair && leader) ? 2 : 1); mbar_init(&empty_bar[s], pair ? 1 : cs); }
        for (int s = 0; s < TC_ACC_STAGES; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], pair ? 256 : 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if constexpr (pair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (cs > 1) cluster_sync_all();                    // every CTA's barriers exist before a peer multicasts into them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int u = cluster_id; u < num_units; u += num_clusters) {
                const int ct = u / n_groups, ng = u % n_groups;
                const int m0 = (ct * (int)cs + (int)crank) * TC_BLOCK_M;
                for (int nt = ng * tiles_per_group; nt < (ng + 1) * tiles_per_group; ++nt) {
                    const int n0 = nt * TC_BLOCK_N;
                    for (int kb = 0; kb < num_k_blocks; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);          // the MMAs of every CTA of the cluster have left this stage
                        uint8_t *st = smem + stage * stage_bytes;
                        mbar_arrive_expect_tx(&full_bar[stage], ((p.dbg & 4) ? 0u : 2 * TC_A_BYTES) + ((p.dbg & 2) ? 0u : 2 * w_tile_bytes));
                        if (!(p.dbg & 4)) {
                            tma_load_2d(&tm_a_hi, &full_bar[stage], st, kb * TC_BLOCK_K, m0);
                            tma_load_2d(&tm_a_lo, &full_bar[stage], st + TC_A_BYTES, kb * TC_BLOCK_K, m0);
                        }
                        uint8_t *wh = st + 2 * TC_A_BYTES + (pair ? 0u : crank * w_slice_bytes), *wl = wh + w_tile_bytes;
                        const int wrow = n0 + (int)(crank * w_slice_rows);          // pair: rows [n0 + crank * 128, + 128)
                        if (p.dbg & 2) {
                        } else if (cs > 1 && !pair) {                                      // this CTA's slice of the weight box, to all peers
                            tma_load_2d_mc(&tm_w_hi, &full_bar[stage], wh, kb * TC_BLOCK_K, wrow, mc_mask);
                            tma_load_2d_mc(&tm_w_lo, &full_bar[stage], wl, kb * TC_BLOCK_K, wrow, mc_mask);
                        } else {
                            tma_load_2d(&tm_w_hi, &full_bar[stage], wh, kb * TC_BLOCK_K, wrow);
                            tma_load_2d(&tm_w_lo, &full_bar[stage], wl, kb * TC_BLOCK_K, wrow);
                        }
                        if (++stage == n_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1 && pair && !leader) {
        // ===================== pair mode, peer CTA: tell the leader when this CTA's half of a stage has landed =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const int nkb = ((num_units - cluster_id + num_clusters - 1) / num_clusters) * tiles_per_group * num_k_blocks;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(&full_bar[stage], phase);
                mbar_arrive_remote(&full_bar[stage], 0);
                if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = pair ? make_idesc_f16(2 * TC_BLOCK_M, TC_BLOCK_N) : make_idesc_f16(TC_BLOCK_M, TC_BLOCK_N);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = 0, ntile = ((num_units - cluster_id + num_clusters - 1) / num_clusters) * tiles_per_group; tile < ntile; ++tile) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);      // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * TC_BLOCK_N);
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);           // TMA bytes have landed
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + stage * stage_bytes);
                    const uint64_t d_ah = make_sw128_kmajor_desc(st);
                    const uint64_t d_al = make_sw128_kmajor_desc(st + TC_A_BYTES);
                    const uint64_t d_wh = make_sw128_kmajor_desc(st + 2 * TC_A_BYTES);
                    const uint64_t d_wl = make_sw128_kmajor_desc(st + 2 * TC_A_BYTES + w_tile_bytes);
#pragma unroll
                    for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
                        if (p.dbg & 8) break;
                        const uint64_t koff = (uint64_t)((k * TC_UMMA_K * 2) >> 4);   // +32 B per K step
                        if constexpr (pair) { // 256 x 256 x 16 over the pair: A rows and B (weight) rows are split between the CTAs
                            tc_mma2_f16(tmem_d, d_ah + koff, d_wh + koff, idesc, (kb | k) ? 1u : 0u);
                            tc_mma2_f16(tmem_d, d_al + koff, d_wh + koff, idesc, 1u);
                            tc_mma2_f16(tmem_d, d_ah + koff, d_wl + koff, idesc, 1u);
                        } else {
                            tc_mma_f16(tmem_d, d_ah + koff, d_wh + koff, idesc, (kb | k) ? 1u : 0u);
                            tc_mma_f16(tmem_d, d_al + koff, d_wh + koff, idesc, 1u);
                            tc_mma_f16(tmem_d, d_ah + koff, d_wl + koff, idesc, 1u);
                        }
                    }
                    if constexpr (pair) tc_commit2_mc(&empty_bar[stage], 3);     // both CTAs' stage is free when the pair's MMAs retire
                    else if (cs > 1) tc_commit_mc(&empty_bar[stage], mc_mask);   // ... on every CTA of the cluster (their boxes land here too)
                    else tc_commit(&empty_bar[stage]);            // frees the smem slot when the MMAs retire
                    if (++stage == n_stages) { stage = 0; phase ^= 1; }
                }
                if constexpr (pair) tc_commit2_mc(&tfull_bar[acc], 3);   // accumulators (one half in each CTA's TMEM) complete -> both epilogues
                else tc_commit(&tfull_bar[acc]);                  // accumulator complete -> epilogue
                if (++acc == TC_ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (TMEM -> registers -> swizzled smem staging -> TMA store) =====================
        // Thread = output row.  Row-strided 16-byte global stores cost 6 of the 14.6 ms of the 8-layer network (round-2
        // experiment: 8.4 ms without them), so the 128 x 64 chunk is staged in shared memory in the SWIZZLE_128B image of the
        // output box (conflict-free: thread r writes 16-byte piece j of its row to slot j ^ (r & 7)) and one thread hands
        // it to the TMA unit: full 128-byte lines, rows past M clipped by the tensor map.
        const int ew = warp - 4;                                   // == warp % 4: TMEM lane quadrant
        const int row_in_tile = ew * 32 + lane;
        const int et = threadIdx.x - 128;                          // 0..127
        const bool has_bias = (p.mode != 1);
        const float slope = (p.mode == 0) ? 0.2f : 1.0f, gain = (p.mode == 0) ? 1.41421356237309515f : 1.0f;
        const float inv_wscale = __ldg(p.inv_wscale);
        uint8_t *stg = smem + TC_STAGES * TC_STAGE_BYTES + 1024;   // [2][16 KB]: (hi, lo) of 64 columns, or 2 x 32 fp32 columns
        float *bias_s = reinterpret_cast<float *>(stg + 2 * TC_A_BYTES);   // [256]
        uint8_t *my0 = stg + row_in_tile * 128, *my1 = my0 + TC_A_BYTES;
        const uint32_t sw = (uint32_t)(row_in_tile & 7);
        const bool storer = (et == 0);
        const bool to_f32 = (p.out_f32 != nullptr);
        int acc = 0; uint32_t acc_phase = 0;
        bool ovf = false;
        for (int u = cluster_id; u < num_units; u += num_clusters)
        for (int nt = (u % n_groups) * tiles_per_group; nt < (u % n_groups + 1) * tiles_per_group; ++nt) {
            const int m0 = ((u / n_groups) * (int)cs + (int)crank) * TC_BLOCK_M, n0 = nt * TC_BLOCK_N;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * TC_BLOCK_N);
#pragma unroll 1
            for (int c0 = 0; c0 < TC_BLOCK_N; c0 += 64) {
                uint32_t v[64];
                tc_ld32(taddr + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                tc_ld32(taddr + (uint32_t)(c0 + 32), *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
                tc_wait_ld();
                if (c0 + 64 == TC_BLOCK_N) {                       // accumulator drained: the MMA warp may refill it
                    tc_fence_before();
                    if (pair && !leader) mbar_arrive_remote(&tempty_bar[acc], 0);     // the leader issues for both CTAs
                    else mbar_arrive(&tempty_bar[acc]);
                }
                if (storer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging has been read out
                if (c0 == 0) {                                     // (safe: the previous tile's readers passed a barrier below)
                    bias_s[et] = (has_bias && n0 + et < p.N_total) ? __ldg(&p.bias[n0 + et]) : 0.f;
                    bias_s[et + 128] = (has_bias && n0 + et + 128 < p.N_total) ? __ldg(&p.bias[n0 + et + 128]) : 0.f;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                float f[64];
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    // one instruction stream for the three modes (a per-element branch on p.mode cost 40 % of the kernel):
                    // no bias = + 0, no activation = slope 1, gain 1 -- exact in fp32
                    float x = __fmul_rn(__uint_as_float(v[j]), inv_wscale) + bias_s[c0 + j];
                    x = (x >= 0.f) ? x : __fmul_rn(x, slope);
                    f[j] = __fmul_rn(gain, x);
                }
                if (to_f32) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {                  // columns [0,32) -> box 0, [32,64) -> box 1; 4 floats per piece
                        *reinterpret_cast<float4 *>(my0 + (((uint32_t)j ^ sw) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                        *reinterpret_cast<float4 *>(my1 + (((uint32_t)j ^ sw) << 4)) =
                            make_float4(f[32 + 4 * j], f[33 + 4 * j], f[34 + 4 * j], f[35 + 4 * j]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {                  // 8 fp16 per 16-byte piece
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float a = f[8 * j + 2 * q], b = f[8 * j + 2 * q + 1];
                            const __half h0 = __float2half_rn(a), h1 = __float2half_rn(b);
                            const __half l0 = __float2half_rn(a - __half2float(h0)), l1 = __float2half_rn(b - __half2float(h1));
                            ovf |= (fabsf(a) > 60000.f) | (fabsf(b) > 60000.f);
                            hi[q] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                            lo[q] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
                        }
                        *reinterpret_cast<uint4 *>(my0 + (((uint32_t)j ^ sw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4 *>(my1 + (((uint32_t)j ^ sw) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the TMA unit
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (storer && !(p.dbg & 1) && n0 + c0 < p.N_total) {
                    if (to_f32) {
                        tma_store_2d(&tm_o0, stg, n0 + c0, m0);
                        if (n0 + c0 + 32 < p.N_total) tma_store_2d(&tm_o0, stg + TC_A_BYTES, n0 + c0 + 32, m0);
                    } else {
                        tma_store_2d(&tm_o0, stg, n0 + c0, m0);
                        tma_store_2d(&tm_o1, stg + TC_A_BYTES, n0 + c0, m0);
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            if (++acc == TC_ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
        if (ovf) atomicOr(p.overflow, 1u);
        if (storer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // all output boxes are in global memory
    }

    tc_fence_before();
    __syncthreads();
    if (cs > 1) cluster_sync_all();                    // no CTA leaves while a peer can still signal its barriers
    if (warp == 2) {
        tc_fence_after();
        if constexpr (pair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---- operand preparation ------------------------------------------------------------------------------
// PixelNorm (model.py:14-19) + split into fp16 hi/lo.  One warp per row.
__global__ void pixelnorm_split_kernel(const float *__restrict__ x, __half *__restrict__ hi, __half *__restrict__ lo,
                                       int64_t n, int dim, int do_norm, unsigned *overflow) {
    int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const int lane = threadIdx.x & 31;
    const float4 *xr = reinterpret_cast<const float4 *>(x + row * dim);
    float r = 1.f;
    if (do_norm) {
        float s = 0.f;
        for (int i = lane; i < dim / 4; i += 32) {
            float4 v = xr[i];
            s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        s = warp_sum(s);
        r = 1.0f / sqrtf(s / (float)dim + 1e-8f);
    }
    bool ovf = false;
    for (int i = lane; i < dim / 4; i += 32) {
        float4 v = xr[i];
        float f[4] = {v.x * r, v.y * r, v.z * r, v.w * r};
        __half h[4], l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            h[q] = __float2half_rn(f[q]);
            l[q] = __float2half_rn(f[q] - __half2float(h[q]));
            ovf |= fabsf(f[q]) > 60000.f;
        }
        uint2 ph, pl;
        ph.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
        ph.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
        pl.x = (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16);
        pl.y = (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16);
        reinterpret_cast<uint2 *>(hi + row * dim)[i] = ph;
        reinterpret_cast<uint2 *>(lo + row * dim)[i] = pl;
    }
    if (ovf) atomicOr(overflow, 1u);
}

// per-layer power-of-two weight scale: the largest |w'| = |w| 2^s lands in [8192, 16384), so that hi keeps
// its 11 bits and lo (<= 2^-12 |w'|) stays a normal fp16 number.  scales[0] = 2^s, scales[1] = 2^-s.
__global__ void pick_wscale_kernel(const float *__restrict__ absmax, float *__restrict__ wscale,
                                   float *__restrict__ inv_wscale) {
    float m = *absmax;
    if (!(m > 0.f)) m = 1.f;
    int e = 0;
    frexpf(m, &e);                         // m = f * 2^e, f in [0.5, 1)
    const int s = 14 - e;
    *wscale = ldexpf(1.f, s);
    *inv_wscale = ldexpf(1.f, -s);
}

// weights: w' = (w*scale) * 2^s  ->  fp16 hi / lo
__global__ void weight_split_kernel(const float *__restrict__ pw, int64_t count, const float *__restrict__ wscale_p,
                                    __half *__restrict__ hi, __half *__restrict__ lo) {
    const float wscale = *wscale_p;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        float w = pw[i] * wscale;          // power-of-two scale: exact
        __half h = __float2half_rn(w);
        hi[i] = h;
        lo[i] = __float2half_rn(w - __half2float(h));
    }
}

__global__ void absmax_kernel(const float *__restrict__ x, int64_t count, float *__restrict__ out) {
    float m = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(x[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int *>(out), __float_as_int(m));   // m >= 0: int order == float order
}

// ---- host side ------------------------------------------------------------------------------------------
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

// 2-D fp16 row-major [rows, cols] tensor, box = box_rows x 64 columns, 128B swizzle (rows / cols smaller than the box are fine:
// the TMA unit fills the rest of the box with zeros)
static int make_tmap_f16(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return GSB_ERR_CUDA; }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BLOCK_K, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return GSB_ERR_CUDA; }
    return GSB_OK;
}

// 2-D fp32 row-major [rows, cols] tensor, box = 128 rows x 32 columns (128 bytes), 128B swizzle (epilogue output boxes)
static int make_tmap_f32_out(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return GSB_ERR_CUDA; }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)TC_BLOCK_M};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (fp32 output) failed (%d)", (int)r); return GSB_ERR_CUDA; }
    return GSB_OK;
}

// Tensor-core packed layout (appended after the fp32 SIMT pack inside the same allocation):
//   [n_layers][dim*dim] fp16 W_hi | [n_layers][dim*dim] fp16 W_lo | [3][n_layers] float {inv_wscale, wscale, absmax} | flag
size_t mapping_tc_packed_bytes(int n_layers, int dim) {
    return align_up((size_t)n_layers * dim * dim * 2, 256) * 2 + align_up((size_t)3 * n_layers * sizeof(float), 256) + 256;
}

struct TcPackView {
    __half *w_hi, *w_lo;
    float *inv_wscale, *wscale, *absmax;
    unsigned *overflow;
};
static TcPackView tc_pack_view(void *base, int n_layers, int dim) {
    TcPackView v;
    char *p = reinterpret_cast<char *>(base);
    size_t wb = align_up((size_t)n_layers * dim * dim * 2, 256);
    v.w_hi = reinterpret_cast<__half *>(p);
    v.w_lo = reinterpret_cast<__half *>(p + wb);
    v.inv_wscale = reinterpret_cast<float *>(p + 2 * wb);
    v.wscale = v.inv_wscale + n_layers;
    v.absmax = v.wscale + n_layers;
    v.overflow = reinterpret_cast<unsigned *>(p + 2 * wb + align_up((size_t)3 * n_layers * sizeof(float), 256));
    return v;
}

// Splits the already scale-multiplied fp32 weights `pw` ([n_layers][dim*dim]); stream-ordered, no host sync.
int mapping_tc_pack(const float *pw, int n_layers, int dim, void *tc_base, cudaStream_t st) {
    TcPackView v = tc_pack_view(tc_base, n_layers, dim);
    GSB_CHECK_CUDA(cudaMemsetAsync(v.inv_wscale, 0, (size_t)3 * n_layers * sizeof(float), st));
    GSB_CHECK_CUDA(cudaMemsetAsync(v.overflow, 0, sizeof(unsigned), st));
    const int64_t per = (int64_t)dim * dim;
    for (int l = 0; l < n_layers; ++l) {
        absmax_kernel<<<64, 256, 0, st>>>(pw + l * per, per, v.absmax + l);
        GSB_CHECK_LAUNCH();
        pick_wscale_kernel<<<1, 1, 0, st>>>(v.absmax + l, v.wscale + l, v.inv_wscale + l);
        GSB_CHECK_LAUNCH();
        weight_split_kernel<<<256, 256, 0, st>>>(pw + l * per, per, v.wscale + l, v.w_hi + l * per, v.w_lo + l * per);
        GSB_CHECK_LAUNCH();
    }
    return GSB_OK;
}

static int tc_ensure_attr() {
    static bool attr_set = false;
    if (!attr_set) {
        GSB_CHECK_CUDA(cudaFuncSetAttribute(mapping_layer_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)TC_SMEM_BYTES));
        GSB_CHECK_CUDA(cudaFuncSetAttribute(mapping_layer_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)TC_SMEM_BYTES));
        attr_set = true;
    }
    return GSB_OK;
}

// cluster size of the layer kernel: GANSPACE_B200_MAPPING_CLUSTER = 1 | 2 | 4.  Default 1: weight multicast measured no gain
// (11.5 ms without, 12.2 ms with 4-CTA clusters for 1.01M rows x 8 layers -- the kernel is bound by shared-memory bandwidth,
// which multicast does not relieve: every CTA still receives the whole weight tile).
static int tc_cluster_size() {
    static int cs = 0;
    if (!cs) {
        const char *e = getenv("GANSPACE_B200_MAPPING_CLUSTER");
        cs = e ? atoi(e) : 1;
        if (cs != 1 && cs != 2 && cs != 4) cs = 1;
    }
    return cs;
}

// clusters of `cs` CTAs that can be resident at once (the kernel is persistent: one wave)
static int tc_max_clusters(int cs) {
    static int cache[5] = {0, 0, 0, 0, 0};
    if (!cache[cs]) {
        int n = num_sms() / cs;
        if (cs > 1) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)(num_sms() / cs * cs)); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = TC_SMEM_BYTES;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int q = 0;
            if (cudaOccupancyMaxActiveClusters(&q, mapping_layer_tc_kernel<false>, &cfg) == cudaSuccess && q > 0 && q < n) n = q;
        }
        cache[cs] = n;
    }
    return cache[cs];
}

// one launch of the layer kernel over `m_tiles` row tiles; W tensor maps must have been built with box rows 256 / cs
static int tc_launch_layer(const CUtensorMap &tm_ah, const CUtensorMap &tm_al, const CUtensorMap &tm_wh, const CUtensorMap &tm_wl,
                           TcParams p, int cs, int leave_free_sms, cudaStream_t st, int pair = 0) {
    p.pair = pair;                       // pair mode needs cs == 2 and W tensor maps with 128-row boxes
    if (pair && cs != 2) { set_error("tc_launch_layer: pair mode needs clusters of 2"); return GSB_ERR_ARG; }
    // output boxes of the epilogue's TMA stores: fp16 hi / lo [M, N] (the next layer's A operand) or fp32 [M, N]
    CUtensorMap tm_o0, tm_o1;
    if (p.out_f32) {
        if (int r = make_tmap_f32_out(&tm_o0, p.out_f32, (uint64_t)p.M, (uint64_t)p.N_total)) return r;
        tm_o1 = tm_o0;
    } else {
        if (int r = make_tmap_f16(&tm_o0, p.out_hi, (uint64_t)p.M, (uint64_t)p.N_total, TC_BLOCK_M)) return r;
        if (int r = make_tmap_f16(&tm_o1, p.out_lo, (uint64_t)p.M, (uint64_t)p.N_total, TC_BLOCK_M)) return r;
    }
    const int m_tiles = (p.M + TC_BLOCK_M - 1) / TC_BLOCK_M, n_tiles = (p.N_total + TC_BLOCK_N - 1) / TC_BLOCK_N;
    const int num_ct = (m_tiles + cs - 1) / cs;
    int avail = (num_sms() - leave_free_sms) / cs;
    if (avail < 16 / cs) avail = 16 / cs;
    const int cap = tc_max_clusters(cs);
    if (avail > cap) avail = cap;
    p.n_groups = (num_ct >= 2 * avail) ? 1 : n_tiles;
    {
        static int dbg = -1;
        if (dbg < 0) { const char *e = getenv("GANSPACE_B200_MAPPING_DBG"); dbg = e ? atoi(e) : 0; }
        p.dbg = dbg;
    }
    const int units = num_ct * p.n_groups;
    const int clusters = units < avail ? units : avail;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(clusters * cs)); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = TC_SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (pair) GSB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, mapping_layer_tc_kernel<true>, tm_ah, tm_al, tm_wh, tm_wl, tm_o0, tm_o1, p));
    else GSB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, mapping_layer_tc_kernel<false>, tm_ah, tm_al, tm_wh, tm_wl, tm_o0, tm_o1, p));
    return GSB_OK;
}

// out[M, N] (fp32, row-major) = (A_hi + A_lo)[M, K] * (W_hi + W_lo)[N, K]^T * inv_wscale   -- the same persistent
// tcgen05 kernel with the plain epilogue.  Both operands are K-major fp16 hi/lo pairs; K % 64 == 0, N % 256 == 0.
// Used by the modulated-convolution path (synthesis.cu): one dense contraction per 3x3 tap.
int tc_gemm_plain(const __half *a_hi, const __half *a_lo, int64_t M, int K, const __half *w_hi, const __half *w_lo, int N,
                  const float *inv_wscale, float *out, unsigned *overflow, int leave_free_sms, cudaStream_t st) {
    GSB_CHECK_ARG(N % 32 == 0 && N >= 32 && K % 8 == 0 && K >= 8 && M > 0 && M < (1ll << 31),
                  "tc_gemm_plain: need N%%32==0, K%%8==0 (M=%lld N=%d K=%d)", (long long)M, N, K);
    if (int r = tc_ensure_attr()) return r;
    const int cs = 1;                  // (no weight multicast for the tap GEMMs: see tc_cluster_size)
    CUtensorMap tm_ah, tm_al, tm_wh, tm_wl;
    if (int r = make_tmap_f16(&tm_ah, a_hi, (uint64_t)M, (uint64_t)K, TC_BLOCK_M)) return r;
    if (int r = make_tmap_f16(&tm_al, a_lo, (uint64_t)M, (uint64_t)K, TC_BLOCK_M)) return r;
    if (int r = make_tmap_f16(&tm_wh, w_hi, (uint64_t)N, (uint64_t)K, TC_BLOCK_N / cs)) return r;
    if (int r = make_tmap_f16(&tm_wl, w_lo, (uint64_t)N, (uint64_t)K, TC_BLOCK_N / cs)) return r;
    TcParams p;
    p.bias = nullptr; p.out_hi = nullptr; p.out_lo = nullptr; p.out_f32 = out; p.overflow = overflow;
    p.inv_wscale = inv_wscale; p.M = (int)M; p.N_total = N; p.K = K; p.mode = 1; p.n_groups = 1;
    return tc_launch_layer(tm_ah, tm_al, tm_wh, tm_wl, p, cs, leave_free_sms, st);
}

// y[n, N] = x[n, K] W[N, K]^T + bias (optionally sqrt2 * lrelu) on the tensor cores, fp32-grade (hi/lo split of both operands
// done here: W per call -- N*K elements, negligible next to the n*N*K product for n >= 128).  N % 256 == 0, K % 64 == 0.
// Used for BigGAN's generator.gen_z (biggan model.py:211-212,232): [B, 256] x [256, 32768].
// ws layout: x_hi, x_lo [n*K] fp16 | w_hi, w_lo [N*K] fp16 | {inv_wscale, wscale, absmax} | overflow flag
size_t tc_linear_workspace_bytes(int64_t n, int N, int K) {
    return 2 * align_up((size_t)n * K * 2, 256) + 2 * align_up((size_t)N * K * 2, 256) + 512;
}

int tc_linear(const float *x, const float *w, const float *bias, float *y, int64_t n, int N, int K, bool lrelu, void *ws,
              cudaStream_t st) {
    GSB_CHECK_ARG(N % TC_BLOCK_N == 0 && K % TC_BLOCK_K == 0 && n > 0 && n < (1ll << 31) && bias, "tc_linear: need N%%256==0, K%%64==0, bias");
    if (int r = tc_ensure_attr()) return r;
    char *p0 = reinterpret_cast<char *>(ws);
    const size_t xb = align_up((size_t)n * K * 2, 256), wb = align_up((size_t)N * K * 2, 256);
    __half *x_hi = (__half *)p0, *x_lo = (__half *)(p0 + xb), *w_hi = (__half *)(p0 + 2 * xb), *w_lo = (__half *)(p0 + 2 * xb + wb);
    float *scal = (float *)(p0 + 2 * xb + 2 * wb);              // inv_wscale, wscale, absmax
    unsigned *overflow = (unsigned *)(p0 + 2 * xb + 2 * wb + 256);
    GSB_CHECK_CUDA(cudaMemsetAsync(scal, 0, 512, st));
    absmax_kernel<<<256, 256, 0, st>>>(w, (int64_t)N * K, scal + 2);
    GSB_CHECK_LAUNCH();
    pick_wscale_kernel<<<1, 1, 0, st>>>(scal + 2, scal + 1, scal);
    GSB_CHECK_LAUNCH();
    weight_split_kernel<<<1024, 256, 0, st>>>(w, (int64_t)N * K, scal + 1, w_hi, w_lo);
    GSB_CHECK_LAUNCH();
    pixelnorm_split_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(x, x_hi, x_lo, n, K, 0, overflow);
    GSB_CHECK_LAUNCH();
    CUtensorMap tm_ah, tm_al, tm_wh, tm_wl;
    if (int r = make_tmap_f16(&tm_ah, x_hi, (uint64_t)n, (uint64_t)K, TC_BLOCK_M)) return r;
    if (int r = make_tmap_f16(&tm_al, x_lo, (uint64_t)n, (uint64_t)K, TC_BLOCK_M)) return r;
    if (int r = make_tmap_f16(&tm_wh, w_hi, (uint64_t)N, (uint64_t)K, TC_BLOCK_N)) return r;
    if (int r = make_tmap_f16(&tm_wl, w_lo, (uint64_t)N, (uint64_t)K, TC_BLOCK_N)) return r;
    TcParams p;
    p.bias = bias; p.out_hi = nullptr; p.out_lo = nullptr; p.out_f32 = y; p.overflow = overflow;
    p.inv_wscale = scal; p.M = (int)n; p.N_total = N; p.K = K; p.mode = lrelu ? 0 : 2; p.n_groups = 1;
    return tc_launch_layer(tm_ah, tm_al, tm_wh, tm_wl, p, 1, 0, st);
}

// Full mapping network on the tensor cores.  ws: 4 fp16 buffers of n*dim (two hi/lo ping-pong pairs).
int mapping_forward_tc(const float *pb, void *tc_base, int n_layers, int dim,
                       const float *d_z, float *d_w, int64_t n, bool pixelnorm, void *ws, int leave_free_sms,
                       cudaStream_t st) {
    GSB_CHECK_ARG(dim % TC_BLOCK_N == 0 && dim % TC_BLOCK_K == 0, "mapping_forward_tc: dim must be a multiple of 256");
    GSB_CHECK_ARG(n < (1ll << 31), "mapping_forward_tc: too many rows");
    TcPackView v = tc_pack_view(tc_base, n_layers, dim);
    const size_t buf = align_up((size_t)n * dim * 2, 256);
    __half *a_hi[2] = {reinterpret_cast<__half *>(ws), reinterpret_cast<__half *>((char *)ws + 2 * buf)};
    __half *a_lo[2] = {reinterpret_cast<__half *>((char *)ws + buf), reinterpret_cast<__half *>((char *)ws + 3 * buf)};

    if (int r = tc_ensure_attr()) return r;
    pixelnorm_split_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(d_z, a_hi[0], a_lo[0], n, dim, pixelnorm ? 1 : 0,
                                                                  v.overflow);
    GSB_CHECK_LAUNCH();
    // persistent CTAs in clusters, one CTA per SM; `leave_free_sms` keeps some SMs idle for a concurrent latency-critical
    // stream (the IPCA chain's 16-CTA cluster kernels), which otherwise wait for a whole layer launch to drain
    const int m_tiles = (int)((n + TC_BLOCK_M - 1) / TC_BLOCK_M);
    // GANSPACE_B200_MAPPING_PAIR=1: CTA pairs issuing tcgen05.mma.cta_group::2 (256 x 256 tiles, each CTA stages half of the
    // weight tile: 64 B/clk of MMA operand reads and 42 B/clk of TMA fills per SM instead of 96 + 62).  Measured round 2: correct
    // (1.4e-5 vs fp64, as the 1-CTA form) but SLOWER, 15.9 ms against 11.5 ms for 1.01M rows x 8 layers -- the peer's "my half
    // has landed" relay through a remote mbarrier arrive sits on the critical path of every stage.  Kept opt-in.
    static int pair_mode = -1;
    if (pair_mode < 0) { const char *e = getenv("GANSPACE_B200_MAPPING_PAIR"); pair_mode = (e && atoi(e) == 1) ? 1 : 0; }
    const int pair = (pair_mode && m_tiles >= 64) ? 1 : 0;
    const int cs = pair ? 2 : ((m_tiles >= 64) ? tc_cluster_size() : 1);
    const int64_t per = (int64_t)dim * dim;
    for (int l = 0; l < n_layers; ++l) {
        const int src = l & 1, dst = src ^ 1;
        CUtensorMap tm_ah, tm_al, tm_wh, tm_wl;
        if (int r = make_tmap_f16(&tm_ah, a_hi[src], (uint64_t)n, dim, TC_BLOCK_M)) return r;
        if (int r = make_tmap_f16(&tm_al, a_lo[src], (uint64_t)n, dim, TC_BLOCK_M)) return r;
        if (int r = make_tmap_f16(&tm_wh, v.w_hi + l * per, dim, dim, TC_BLOCK_N / cs)) return r;
        if (int r = make_tmap_f16(&tm_wl, v.w_lo + l * per, dim, dim, TC_BLOCK_N / cs)) return r;
        TcParams p;
        p.bias = pb + (int64_t)l * dim;
        const bool last = (l == n_layers - 1);
        p.out_hi = last ? nullptr : a_hi[dst];
        p.out_lo = last ? nullptr : a_lo[dst];
        p.out_f32 = last ? d_w : nullptr;
        p.overflow = v.overflow;
        p.inv_wscale = v.inv_wscale + l;
        p.M = (int)n; p.N_total = dim; p.K = dim; p.mode = 0; p.n_groups = 1;
        if (int r = tc_launch_layer(tm_ah, tm_al, tm_wh, tm_wl, p, cs, leave_free_sms, st, pair)) return r;
    }
    return GSB_OK;
}

size_t mapping_tc_workspace_bytes(int64_t n, int dim) { return 4 * align_up((size_t)n * dim * 2, 256); }

unsigned *mapping_tc_overflow_flag(void *tc_base, int n_layers, int dim) {
    return tc_pack_view(tc_base, n_layers, dim).overflow;
}

}  // namespace gsb
