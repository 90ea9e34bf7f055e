// Bit-exact NumPy legacy RandomState streams on the device.
//
// Replaces the host RNG of the reference's latent samplers:
//   models/wrappers.py:167-175 (StyleGAN2.sample_latent: RandomState(seed).standard_normal(512*n))
//   models/biggan/pytorch_biggan/pytorch_pretrained_biggan/utils.py:21-33 (truncnorm.rvs(-2,2,...))
//
// Design: MT19937 is a lag-(624,397) linear recurrence, so a single stream is inherently sequential
// (at most 227 new words per dependent step).  The reference draws ONE stream per sample_latent() call
// (~5.1e6 normals, ~1.3e7 words for a 10k batch) and ~100 independent streams per run, so the mapping is
// one CTA per stream, all streams of a run in one launch (101 CTAs on 148 SMs for config 2).  Inside the
// CTA the work is warp-specialised and double-buffered through shared memory:
//   producer warps 0-7   regenerate 16 x 624 state words per buffer (3 barrier-separated phases per 624,
//                        ping-pong state, named barrier among the 256 producer threads only) and temper
//                        them into a 9984-word buffer;
//   consumer warps 8-31  turn the previous buffer into 2496 polar-method attempts (fp64, no FMA
//                        contraction so the accept/reject decisions are those of the C code NumPy runs),
//                        thread t owning 4 consecutive attempts so that output order == thread order;
//                        a scan of the accept counts gives each thread its output offset
//                        (order-preserving compaction of the rejection sampler); log/div/sqrt run only for
//                        accepted pairs, staged in smem and streamed out coalesced.
// The two groups hand buffers over with full/empty named barriers, so the latency-bound state recurrence
// and the fp64 transform overlap.  HBM traffic = the fp32 outputs only (4 B per normal).
#include "common.cuh"

namespace gsb {

constexpr int MT_N = 624;
constexpr int MT_M = 397;
constexpr int RNG_THREADS = 1024;
constexpr int RNG_PRODUCERS = 256;                         // warps 0-7
constexpr int RNG_CONSUMERS = RNG_THREADS - RNG_PRODUCERS; // warps 8-31
constexpr int BLOCKS_PER_SUPER = 16;
constexpr int WORDS_PER_SUPER = MT_N * BLOCKS_PER_SUPER;   // 9984
constexpr int ATT_PER_SUPER = WORDS_PER_SUPER / 4;         // 2496 polar attempts
constexpr int ATT_PER_THREAD = 4;                          // 768*4 >= 2496

enum { MODE_RAW = 0, MODE_NORMAL = 1, MODE_TRUNCNORM = 2 };
enum { BAR_FULL0 = 1, BAR_FULL1 = 2, BAR_EMPTY0 = 3, BAR_EMPTY1 = 4, BAR_PROD = 5, BAR_CONS = 6 };

// bar.sync / bar.arrive are warp-aligned instructions: re-converge the warp first (the lanes come out of
// divergent `if (lane == ...)` blocks that the compiler does not know must re-join before inline PTX).
__device__ __forceinline__ void named_sync(int id, int count) {
    __syncwarp();
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int count) {
    __syncwarp();
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

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
    uint32_t words[2][WORDS_PER_SUPER];
    float outbuf[2 * ATT_PER_SUPER];   // normals of one buffer (also 4992 truncnorm outputs)
    int warp_tot[RNG_CONSUMERS / 32];
    int done;
};

template <int MODE>
__global__ void __launch_bounds__(RNG_THREADS, 1)
mt_stream_kernel(const uint32_t *__restrict__ seeds, int64_t n_per_stream, void *__restrict__ out_v,
                 int64_t out_stride, double tn_pa, double tn_pw, float tn_scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RngSmem &sm = *reinterpret_cast<RngSmem *>(smem_raw);
    const int tid = threadIdx.x;

    // init_genrand(seed): sequential Knuth LCG, 624 steps (negligible next to 1e7 outputs)
    if (tid == 0) {
        uint32_t x = seeds[blockIdx.x];
        sm.state[0][0] = x;
        for (int i = 1; i < MT_N; ++i) {
            x = 1812433253u * (x ^ (x >> 30)) + (uint32_t)i;
            sm.state[0][i] = x;
        }
        sm.done = 0;
    }
    __syncthreads();

    if (tid < RNG_PRODUCERS) {
        // ===================== producers: MT19937 state regeneration =====================
        int cur = 0;
        for (int t = 0;; ++t) {
            const int b = t & 1;
            if (t >= 2) named_sync(b ? BAR_EMPTY1 : BAR_EMPTY0, RNG_THREADS);   // consumers released buffer b
            if (*reinterpret_cast<volatile int *>(&sm.done)) break;
            for (int blk = 0; blk < BLOCKS_PER_SUPER; ++blk) {
                const uint32_t *o = sm.state[cur];
                uint32_t *n = sm.state[cur ^ 1];
                uint32_t *w = sm.words[b] + blk * MT_N;
                if (tid < MT_N - MT_M) {   // i in [0,227)
                    uint32_t v = mt_twist(o[tid], o[tid + 1], o[tid + MT_M]);
                    n[tid] = v;
                    w[tid] = mt_temper(v);
                }
                named_sync(BAR_PROD, RNG_PRODUCERS);
                if (tid < MT_N - MT_M) {   // i in [227,454)
                    int i = tid + (MT_N - MT_M);
                    uint32_t v = mt_twist(o[i], o[i + 1], n[tid]);
                    n[i] = v;
                    w[i] = mt_temper(v);
                }
                named_sync(BAR_PROD, RNG_PRODUCERS);
                if (tid < MT_N - 2 * (MT_N - MT_M)) {   // i in [454,624): 170 words
                    int i = tid + 2 * (MT_N - MT_M);
                    uint32_t nxt = (i == MT_N - 1) ? n[0] : o[i + 1];
                    uint32_t v = mt_twist(o[i], nxt, n[i - (MT_N - MT_M)]);
                    n[i] = v;
                    w[i] = mt_temper(v);
                }
                named_sync(BAR_PROD, RNG_PRODUCERS);
                cur ^= 1;
            }
            named_arrive(b ? BAR_FULL1 : BAR_FULL0, RNG_THREADS);               // buffer b is ready
        }
        return;
    }

    // ===================== consumers: words -> outputs =====================
    const int ct = tid - RNG_PRODUCERS;
    const int lane = ct & 31, warp = ct >> 5;
    int64_t produced = 0;   // outputs written so far (uniform across the consumers)
    for (int t = 0;; ++t) {
        const int b = t & 1;
        named_sync(b ? BAR_FULL1 : BAR_FULL0, RNG_THREADS);
        const uint32_t *words = sm.words[b];
        if (MODE == MODE_RAW) {
            uint32_t *out = reinterpret_cast<uint32_t *>(out_v) + (int64_t)blockIdx.x * out_stride;
            int64_t rem = n_per_stream - produced;
            int cnt = rem < WORDS_PER_SUPER ? (int)rem : WORDS_PER_SUPER;
            for (int i = ct; i < cnt; i += RNG_CONSUMERS) out[produced + i] = words[i];
            produced += cnt;
        } else if (MODE == MODE_TRUNCNORM) {
            // scipy.stats.truncnorm.rvs: u = RandomState.uniform() ; x = ndtri(Phi(a) + u*(Phi(b)-Phi(a)))
            float *out = reinterpret_cast<float *>(out_v) + (int64_t)blockIdx.x * out_stride;
            int64_t rem = n_per_stream - produced;
            int cnt = rem < WORDS_PER_SUPER / 2 ? (int)rem : WORDS_PER_SUPER / 2;
            for (int i = ct; i < cnt; i += RNG_CONSUMERS) {
                double u = mt_double(words[2 * i], words[2 * i + 1]);
                double q = __dadd_rn(tn_pa, __dmul_rn(u, tn_pw));
                out[produced + i] = __fmul_rn(__double2float_rn(normcdfinv(q)), tn_scale);
            }
            produced += cnt;
        } else {
            // ---- accept/reject this thread's consecutive attempts (cheap part only) ---------------
            unsigned accept = 0;
#pragma unroll
            for (int j = 0; j < ATT_PER_THREAD; ++j) {
                int a = ct * ATT_PER_THREAD + j;
                if (a < ATT_PER_SUPER) {
                    uint4 wv = *reinterpret_cast<const uint4 *>(&words[4 * a]);
                    double x1 = __dadd_rn(__dmul_rn(2.0, mt_double(wv.x, wv.y)), -1.0);
                    double x2 = __dadd_rn(__dmul_rn(2.0, mt_double(wv.z, wv.w)), -1.0);
                    double r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
                    if (r2 < 1.0 && r2 != 0.0) accept |= 1u << j;
                }
            }
            const int cnt = __popc(accept);
            // ---- order-preserving compaction: scan of the accept counts over the consumer threads ----
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) sm.warp_tot[warp] = incl;
            named_sync(BAR_CONS, RNG_CONSUMERS);
            int base = 0, total = 0;
#pragma unroll
            for (int wq = 0; wq < RNG_CONSUMERS / 32; ++wq) {
                int v = sm.warp_tot[wq];
                if (wq < warp) base += v;
                total += v;
            }
            int off = 2 * (base + incl - cnt);
            // ---- the expensive part (log, div, sqrt) only for accepted pairs, straight to its slot ----
#pragma unroll
            for (int j = 0; j < ATT_PER_THREAD; ++j) {
                if (accept & (1u << j)) {
                    int a = ct * ATT_PER_THREAD + j;
                    uint4 wv = *reinterpret_cast<const uint4 *>(&words[4 * a]);
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
            named_sync(BAR_CONS, RNG_CONSUMERS);
            float *out = reinterpret_cast<float *>(out_v) + (int64_t)blockIdx.x * out_stride;
            int64_t rem = n_per_stream - produced;
            int nout = 2 * total;
            if ((int64_t)nout > rem) nout = (int)rem;
            for (int i = ct; i < nout; i += RNG_CONSUMERS) out[produced + i] = sm.outbuf[i];
            produced += nout;
            named_sync(BAR_CONS, RNG_CONSUMERS);      // outbuf / warp_tot are reused by the next buffer
        }
        const bool finished = produced >= n_per_stream;
        if (finished && ct == 0) *reinterpret_cast<volatile int *>(&sm.done) = 1;
        __threadfence_block();
        named_arrive(b ? BAR_EMPTY1 : BAR_EMPTY0, RNG_THREADS);                // buffer b may be refilled
        if (finished) break;
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
