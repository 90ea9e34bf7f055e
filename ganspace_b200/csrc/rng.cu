// Bit-exact NumPy legacy RandomState streams on the device.
//
// Replaces the host RNG of the reference's latent samplers:
//   models/wrappers.py:167-175 (StyleGAN2.sample_latent: RandomState(seed).standard_normal(512*n))
//   models/biggan/pytorch_biggan/pytorch_pretrained_biggan/utils.py:21-33 (truncnorm.rvs(-2,2,...))
//
// Design: MT19937 is a lag-(624,397) linear recurrence, so a single stream is inherently sequential
// (at most 227 new words per dependent step).  The reference draws ONE stream per sample_latent() call
// (~5.1e6 normals, ~1.3e7 words for a 10k batch) and ~100 independent streams per run, so the mapping is
// one CTA per stream, all streams of a run in one launch (101 CTAs on 148 SMs for config 2):
//   1. the CTA regenerates 16 x 624 state words (3 barrier-separated phases per 624, ping-pong in smem),
//      tempering them into a 9984-word smem buffer;
//   2. 1024 threads turn those words into 2496 polar-method attempts (fp64, no FMA contraction so the
//      accept/reject decisions are those of the C code NumPy runs), thread t owning 3 consecutive
//      attempts so that output order == thread order;
//   3. a block scan of the accept counts gives each thread its output offset (order-preserving
//      compaction of the rejection sampler); log/div/sqrt run only for accepted pairs, whose results
//      are staged in smem and streamed out coalesced.
// HBM traffic = the fp32 outputs only (4 B per normal).
#include "common.cuh"

namespace gsb {

constexpr int MT_N = 624;
constexpr int MT_M = 397;
constexpr int RNG_THREADS = 1024;   // 32 warps: the fp64 log/div/sqrt chains of the transform need the TLP
constexpr int BLOCKS_PER_SUPER = 16;
constexpr int WORDS_PER_SUPER = MT_N * BLOCKS_PER_SUPER;   // 9984
constexpr int ATT_PER_SUPER = WORDS_PER_SUPER / 4;         // 2496 polar attempts
constexpr int ATT_PER_THREAD = 3;                          // 1024*3 >= 2496

enum { MODE_RAW = 0, MODE_NORMAL = 1, MODE_TRUNCNORM = 2 };

__device__ __forceinline__ uint32_t mt_twist(uint32_t u, uint32_t v, uint32_t m) {
    uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
    return m ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
// random_sample(): 53-bit double from two consecutive outputs.
__device__ __forceinline__ double mt_double(uint32_t a, uint32_t b) {
    // (a>>5)*2^26 + (b>>6) is an exact integer < 2^53; the scaling by 2^-53 is exact.
    double hi = (double)(a >> 5), lo = (double)(b >> 6);
    return __dmul_rn(__dadd_rn(__dmul_rn(hi, 67108864.0), lo), 1.0 / 9007199254740992.0);
}

struct RngSmem {
    uint32_t state[2][MT_N];
    uint32_t words[WORDS_PER_SUPER];
    float outbuf[2 * ATT_PER_SUPER];   // MODE_NORMAL staging (also reused for truncnorm: 4992 floats)
    int warp_tot[RNG_THREADS / 32];
};

template <int MODE>
__global__ void __launch_bounds__(RNG_THREADS, 1)
mt_stream_kernel(const uint32_t *__restrict__ seeds, int64_t n_per_stream, void *__restrict__ out_v,
                 int64_t out_stride, double tn_pa, double tn_pw, float tn_scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RngSmem &sm = *reinterpret_cast<RngSmem *>(smem_raw);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    // init_genrand(seed): sequential Knuth LCG, 624 steps (negligible next to 1e7 outputs)
    if (tid == 0) {
        uint32_t x = seeds[blockIdx.x];
        sm.state[0][0] = x;
        for (int i = 1; i < MT_N; ++i) {
            x = 1812433253u * (x ^ (x >> 30)) + (uint32_t)i;
            sm.state[0][i] = x;
        }
    }
    __syncthreads();

    int cur = 0;
    int64_t produced = 0;   // outputs written so far (uniform across the CTA)

    while (produced < n_per_stream) {
        // ---- 1. regenerate 16 x 624 words ----------------------------------------------------
        for (int b = 0; b < BLOCKS_PER_SUPER; ++b) {
            const uint32_t *o = sm.state[cur];
            uint32_t *n = sm.state[cur ^ 1];
            uint32_t *w = sm.words + b * MT_N;
            if (tid < MT_N - MT_M) {   // i in [0,227)
                uint32_t v = mt_twist(o[tid], o[tid + 1], o[tid + MT_M]);
                n[tid] = v;
                w[tid] = mt_temper(v);
            }
            __syncthreads();
            if (tid < MT_N - MT_M) {   // i in [227,454)
                int i = tid + (MT_N - MT_M);
                uint32_t v = mt_twist(o[i], o[i + 1], n[tid]);
                n[i] = v;
                w[i] = mt_temper(v);
            }
            __syncthreads();
            if (tid < MT_N - 2 * (MT_N - MT_M)) {   // i in [454,624): 170 words
                int i = tid + 2 * (MT_N - MT_M);
                uint32_t nxt = (i == MT_N - 1) ? n[0] : o[i + 1];
                uint32_t v = mt_twist(o[i], nxt, n[i - (MT_N - MT_M)]);
                n[i] = v;
                w[i] = mt_temper(v);
            }
            __syncthreads();
            cur ^= 1;
        }

        if (MODE == MODE_RAW) {
            uint32_t *out = reinterpret_cast<uint32_t *>(out_v) + (int64_t)blockIdx.x * out_stride;
            int64_t rem = n_per_stream - produced;
            int cnt = rem < WORDS_PER_SUPER ? (int)rem : WORDS_PER_SUPER;
            for (int i = tid; i < cnt; i += RNG_THREADS) out[produced + i] = sm.words[i];
            produced += cnt;
            __syncthreads();
            continue;
        }
        if (MODE == MODE_TRUNCNORM) {
            // scipy.stats.truncnorm.rvs: u = RandomState.uniform() ; x = ndtri(Phi(a) + u*(Phi(b)-Phi(a)))
            float *out = reinterpret_cast<float *>(out_v) + (int64_t)blockIdx.x * out_stride;
            int64_t rem = n_per_stream - produced;
            int cnt = rem < WORDS_PER_SUPER / 2 ? (int)rem : WORDS_PER_SUPER / 2;
            for (int i = tid; i < cnt; i += RNG_THREADS) {
                double u = mt_double(sm.words[2 * i], sm.words[2 * i + 1]);
                double q = __dadd_rn(tn_pa, __dmul_rn(u, tn_pw));
                out[produced + i] = __fmul_rn(__double2float_rn(normcdfinv(q)), tn_scale);
            }
            produced += cnt;
            __syncthreads();
            continue;
        }

        // ---- 2. accept/reject this thread's consecutive attempts (cheap part only) -------------
        unsigned accept = 0;
#pragma unroll
        for (int j = 0; j < ATT_PER_THREAD; ++j) {
            int a = tid * ATT_PER_THREAD + j;
            if (a < ATT_PER_SUPER) {
                uint4 wv = *reinterpret_cast<const uint4 *>(&sm.words[4 * a]);
                double x1 = __dadd_rn(__dmul_rn(2.0, mt_double(wv.x, wv.y)), -1.0);
                double x2 = __dadd_rn(__dmul_rn(2.0, mt_double(wv.z, wv.w)), -1.0);
                double r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
                if (r2 < 1.0 && r2 != 0.0) accept |= 1u << j;
            }
        }
        const int cnt = __popc(accept);
        // ---- 3. order-preserving compaction: block scan of the accept counts ---------------------
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) sm.warp_tot[warp] = incl;
        __syncthreads();
        int base = 0, total = 0;
#pragma unroll
        for (int wq = 0; wq < RNG_THREADS / 32; ++wq) {
            int t = sm.warp_tot[wq];
            if (wq < warp) base += t;
            total += t;
        }
        int off = 2 * (base + incl - cnt);
        // ---- 4. the expensive part (log, div, sqrt) only for accepted pairs, straight to its slot --
#pragma unroll
        for (int j = 0; j < ATT_PER_THREAD; ++j) {
            if (accept & (1u << j)) {
                int a = tid * ATT_PER_THREAD + j;
                uint4 wv = *reinterpret_cast<const uint4 *>(&sm.words[4 * a]);
                double x1 = __dadd_rn(__dmul_rn(2.0, mt_double(wv.x, wv.y)), -1.0);
                double x2 = __dadd_rn(__dmul_rn(2.0, mt_double(wv.z, wv.w)), -1.0);
                double r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
                double f = __dsqrt_rn(__ddiv_rn(__dmul_rn(-2.0, log(r2)), r2));
                // legacy_gauss returns f*x2 first and caches f*x1 for the next call
                sm.outbuf[off] = __double2float_rn(__dmul_rn(f, x2));
                sm.outbuf[off + 1] = __double2float_rn(__dmul_rn(f, x1));
                off += 2;
            }
        }
        __syncthreads();
        {
            float *out = reinterpret_cast<float *>(out_v) + (int64_t)blockIdx.x * out_stride;
            int64_t rem = n_per_stream - produced;
            int nout = 2 * total;
            if ((int64_t)nout > rem) nout = (int)rem;
            for (int i = tid; i < nout; i += RNG_THREADS) out[produced + i] = sm.outbuf[i];
            produced += nout;
        }
        __syncthreads();
    }
}

template <int MODE>
static int launch_stream(const uint32_t *d_seeds, int n_streams, int64_t n_per_stream, void *d_out,
                         int64_t out_stride, double pa, double pw, float scale, gsb_stream_t stream) {
    GSB_CHECK_ARG(d_seeds && d_out, "rng: null pointer");
    GSB_CHECK_ARG(n_streams >= 0 && n_per_stream >= 0 && out_stride >= n_per_stream,
                  "rng: bad sizes (n_streams=%d n_per_stream=%lld stride=%lld)", n_streams,
                  (long long)n_per_stream, (long long)out_stride);
    if (n_streams == 0 || n_per_stream == 0) return GSB_OK;
    static bool attr_set = false;
    if (!attr_set) {
        GSB_CHECK_CUDA(cudaFuncSetAttribute(mt_stream_kernel<MODE>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)sizeof(RngSmem)));
        attr_set = true;
    }
    mt_stream_kernel<MODE><<<n_streams, RNG_THREADS, sizeof(RngSmem), (cudaStream_t)stream>>>(
        d_seeds, n_per_stream, d_out, out_stride, pa, pw, scale);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

}  // namespace gsb

extern "C" int gsb_legacy_normal_f32(const uint32_t *d_seeds, int n_streams, int64_t n_per_stream,
                                     float *d_out, int64_t out_stride, gsb_stream_t stream) {
    return gsb::launch_stream<gsb::MODE_NORMAL>(d_seeds, n_streams, n_per_stream, d_out, out_stride,
                                                0.0, 0.0, 1.0f, stream);
}

extern "C" int gsb_mt19937_raw_u32(const uint32_t *d_seeds, int n_streams, int64_t n_per_stream,
                                   uint32_t *d_out, int64_t out_stride, gsb_stream_t stream) {
    return gsb::launch_stream<gsb::MODE_RAW>(d_seeds, n_streams, n_per_stream, d_out, out_stride, 0.0,
                                             0.0, 1.0f, stream);
}

extern "C" int gsb_legacy_truncnorm_f32(const uint32_t *d_seeds, int n_streams, int64_t n_per_stream,
                                        double lo, double hi, float scale, float *d_out,
                                        int64_t out_stride, gsb_stream_t stream) {
    GSB_CHECK_ARG(lo < hi, "truncnorm: lo >= hi");
    // Phi(lo), Phi(hi)-Phi(lo) in fp64 on the host (erfc-based, same as scipy.special.ndtr)
    double pa = 0.5 * erfc(-lo / 1.4142135623730951);
    double pb = 0.5 * erfc(-hi / 1.4142135623730951);
    return gsb::launch_stream<gsb::MODE_TRUNCNORM>(d_seeds, n_streams, n_per_stream, d_out,
                                                   out_stride, pa, pb - pa, scale, stream);
}
