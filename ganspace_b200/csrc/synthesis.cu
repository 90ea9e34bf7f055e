// StyleGAN2 synthesis blocks up to a hooked StyledConv (BASELINE config 5 family: layer = conv1 | convs.k).
//
// Replaces models/stylegan2/stylegan2-pytorch/model.py:181-277 (ModulatedConv2d), :280-291 (NoiseInjection),
// :294-304 (ConstantInput), :307-341 (StyledConv), op/fused_act.py:86-92 (FusedLeakyReLU), model.py:75-91 +
// op/upfirdn2d.py:144-198 (Blur after the stride-2 transposed conv), as driven by models/wrappers.py:224-255.
//
// The reference builds per-sample weights  w[b] = scale * W * s[b,ci] * demod[b,co]  ([B,co,ci,3,3] = 9.4 MB per
// sample and layer) and runs a grouped conv.  Here the SHARED weights are used:
//     xs[b,p,ci]      = x[b,p,ci] * s[b,ci]                               (modulation moved to the input)
//     Y[b,p,tap,co]   = sum_ci (scale W)[co,ci,tap] xs[b,p,ci]            one dense contraction, K = ci, N = 9 co
//                                                                        -> tcgen05 GEMM (mapping_tc.cu, fp16 hi/lo split)
//     stride 1 :  out[b,y,x,co] = sum_tap Y[b,(y+ky-1,x+kx-1),tap,co]
//     upsample :  T[b,u,v,co]   = sum_{tap: u-ky, v-kx even} Y[b,((u-ky)/2,(v-kx)/2),tap,co]      (2H+1 grid)
//                 out[b,y,x,co] = sum_{i,j<4} k[i]k[j]/16 * T[b,y+i-1,x+j-1,co],  k = [1,3,3,1]     (Blur, pad (1,1))
//     out = out * demod[b,co] + noise_w * noise[y,x] + bias[co];  out = sqrt2 * leaky_relu_0.2(out)
//     demod[b,co] = rsqrt( sum_ci s[b,ci]^2 * sum_tap (scale W)[co,ci,tap]^2 + 1e-8 )
// which is algebraically the reference's computation (oracle: styled_conv_taps == styled_conv_forward).
//
// Layout: activations are NHWC ([b, y, x, c]) so that the GEMM operand is K-major and the gathers are coalesced
// over channels; between layers they travel as fp16 hi/lo pairs already multiplied by the NEXT layer's style.
// The hooked layer's activation is written as fp32 NHWC rows of length res*res*co with a caller-given row stride
// (directly into the large-d IPCA batch buffer).  Samples are processed in chunks whose tap planes (Y) fit the L2.
#include "common.cuh"
#include <cuda_fp16.h>
#include <math.h>

namespace gsb {

int tc_gemm_plain(const __half *a_hi, const __half *a_lo, int64_t M, int K, const __half *w_hi, const __half *w_lo, int N,
                  const float *inv_wscale, float *out, unsigned *overflow, int leave_free_sms, cudaStream_t st);

constexpr int SY_MAX_LAYERS = 24;
constexpr int SY_CHUNK_ROWS = 4096;       // GEMM rows per launch: 4096 x 9*512 fp32 tap planes = 75 MB (L2-resident)

// ---- packed layout ----------------------------------------------------------------------------------------
struct SynthLayerView {
    __half *w_hi, *w_lo;      // [9*cout, cin]   row = tap*cout + co
    float *scal;              // [4]: inv_wscale, wscale, absmax
    float *wsq;               // [cout, cin]  sum_tap (scale W)^2
    float *modw;              // [cin, style_dim]  modulation.weight * (1/sqrt(style_dim))
    float *modb;              // [cin]
    float *actb;              // [cout]
    float *noise;             // [res_out^2]  noise_weight * noise
};
struct SynthView {
    float *const_nhwc;        // [16, c0]
    float *zeros;             // [max channels]
    unsigned *overflow;
    SynthLayerView L[SY_MAX_LAYERS];
    size_t bytes;
};
static int res_out_of(const gsb_styled_conv &l) { return l.upsample ? 2 * l.res_in : l.res_in; }

static SynthView synth_view(void *base, const gsb_styled_conv *layers, int n_layers, int style_dim) {
    SynthView v;
    char *p = reinterpret_cast<char *>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { char *q = p + off; off += align_up(bytes, 256); return q; };
    int cmax = 0;
    for (int l = 0; l < n_layers; ++l) { cmax = cmax > layers[l].cin ? cmax : layers[l].cin; cmax = cmax > layers[l].cout ? cmax : layers[l].cout; }
    v.const_nhwc = (float *)take((size_t)16 * (n_layers ? layers[0].cin : 0) * 4);
    v.zeros = (float *)take((size_t)cmax * 4);
    v.overflow = (unsigned *)take(256);
    for (int l = 0; l < n_layers; ++l) {
        const gsb_styled_conv &c = layers[l];
        const int ro = res_out_of(c);
        v.L[l].w_hi = (__half *)take((size_t)9 * c.cout * c.cin * 2);
        v.L[l].w_lo = (__half *)take((size_t)9 * c.cout * c.cin * 2);
        v.L[l].scal = (float *)take(16);
        v.L[l].wsq = (float *)take((size_t)c.cout * c.cin * 4);
        v.L[l].modw = (float *)take((size_t)c.cin * style_dim * 4);
        v.L[l].modb = (float *)take((size_t)c.cin * 4);
        v.L[l].actb = (float *)take((size_t)c.cout * 4);
        v.L[l].noise = (float *)take((size_t)ro * ro * 4);
    }
    v.bytes = off;
    return v;
}

static int check_layers(const gsb_styled_conv *layers, int n_layers, int style_dim) {
    GSB_CHECK_ARG(layers && n_layers >= 1 && n_layers <= SY_MAX_LAYERS, "synthesis: need 1..%d layers", SY_MAX_LAYERS);
    GSB_CHECK_ARG(style_dim > 0 && style_dim % 16 == 0, "synthesis: style_dim %% 16");
    for (int l = 0; l < n_layers; ++l) {
        const gsb_styled_conv &c = layers[l];
        GSB_CHECK_ARG(c.cin % 32 == 0 && c.cout % 32 == 0 && c.cin >= 32 && c.cout >= 32,
                      "synthesis: layer %d needs cin%%32==0, cout%%32==0 (cin=%d cout=%d)", l, c.cin, c.cout);
        GSB_CHECK_ARG(c.res_in >= 4 && c.res_in <= 1024, "synthesis: layer %d bad res_in", l);
        if (l == 0) GSB_CHECK_ARG(c.res_in == 4 && !c.upsample, "synthesis: layer 0 is conv1 on the 4x4 constant");
        else GSB_CHECK_ARG(c.cin == layers[l - 1].cout && c.res_in == res_out_of(layers[l - 1]), "synthesis: layer %d does not chain", l);
    }
    return GSB_OK;
}

// ---- pack kernels ---------------------------------------------------------------------------------------
__global__ void sy_absmax_kernel(const float *__restrict__ x, int64_t count, float scale, float *__restrict__ out) {
    float m = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(x[i] * scale));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int *>(out), __float_as_int(m));
}
// scal[2] = absmax -> scal[1] = 2^s, scal[0] = 2^-s with the largest |w 2^s| in [8192, 16384)
__global__ void sy_pick_scale_kernel(float *__restrict__ scal) {
    float m = scal[2];
    if (!(m > 0.f)) m = 1.f;
    int e = 0;
    frexpf(m, &e);
    scal[1] = ldexpf(1.f, 14 - e);
    scal[0] = ldexpf(1.f, e - 14);
}
// W[co,ci,ky,kx] -> rows (tap, co), K-major over ci, times scale*2^s, split into fp16 hi/lo; and wsq[co,ci]
__global__ void sy_weight_pack_kernel(const float *__restrict__ W, int cout, int cin, float scale, const float *__restrict__ scal,
                                      __half *__restrict__ hi, __half *__restrict__ lo, float *__restrict__ wsq) {
    const float ws = scal[1];
    const int64_t total = (int64_t)cout * cin;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int co = (int)(idx / cin), ci = (int)(idx % cin);
        float sq = 0.f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const float w = W[idx * 9 + tap] * scale;
            sq = fmaf(w, w, sq);
            const float ww = w * ws;
            const __half h = __float2half_rn(ww);
            const int64_t o = ((int64_t)tap * cout + co) * cin + ci;
            hi[o] = h;
            lo[o] = __float2half_rn(ww - __half2float(h));
        }
        wsq[idx] = sq;
    }
}
__global__ void sy_scale_copy_kernel(const float *__restrict__ src, int64_t count, float scale, const float *__restrict__ dev_scale,
                                     float *__restrict__ dst) {
    const float s = dev_scale ? scale * dev_scale[0] : scale;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[i] * s;
}
// const[c, 4, 4] -> [16, c]
__global__ void sy_const_nhwc_kernel(const float *__restrict__ src, int c, float *__restrict__ dst) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 16 * c) { const int p = idx / c, ch = idx % c; dst[idx] = src[ch * 16 + p]; }
}

// ---- forward kernels ------------------------------------------------------------------------------------
__global__ void sy_square_kernel(const float *__restrict__ x, int64_t count, float *__restrict__ y) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < count) { const float v = x[i]; y[i] = v * v; }
}
__global__ void sy_rsqrt_eps_kernel(float *__restrict__ x, int64_t count) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < count) x[i] = 1.0f / sqrtf(x[i] + 1e-8f);
}

__device__ __forceinline__ void store_split4(const float (&f)[4], __half *hi, __half *lo, int64_t off, bool &ovf) {
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
    *reinterpret_cast<uint2 *>(hi + off) = ph;
    *reinterpret_cast<uint2 *>(lo + off) = pl;
}

// ConstantInput (model.py:300-304) times the first layer's style: out[b,p,c] = const[p,c] * s[b,c]  -> hi/lo
__global__ void sy_const_modulate_kernel(const float *__restrict__ cst, const float *__restrict__ s, int64_t n, int hw, int c,
                                         __half *__restrict__ hi, __half *__restrict__ lo, unsigned *overflow) {
    const int cq = c >> 2;
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= n * hw * cq) return;
    const int q = (int)(idx % cq);
    const int64_t pix = idx / cq;
    const int p = (int)(pix % hw);
    const int64_t b = pix / hw;
    const float4 cv = *reinterpret_cast<const float4 *>(cst + (int64_t)p * c + 4 * q);
    const float4 sv = *reinterpret_cast<const float4 *>(s + b * c + 4 * q);
    const float f[4] = {cv.x * sv.x, cv.y * sv.y, cv.z * sv.z, cv.w * sv.w};
    bool ovf = false;
    store_split4(f, hi, lo, pix * c + 4 * q, ovf);
    if (ovf) atomicOr(overflow, 1u);
}

// y[n, N] = x[n, K] W[N, K]^T + bias: one thread per output, for the style / demodulation products of blocks whose channel
// count is below the GEMM kernels' 128-wide tiles (64- and 32-channel blocks at 512^2 / 1024^2)
__global__ void sy_small_linear_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
                                       float *__restrict__ y, int64_t n, int N, int K) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= n * N) return;
    const int j = (int)(idx % N);
    const int64_t r = idx / N;
    const float *xr = x + r * K, *wr = w + (int64_t)j * K;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int k = 0;
    for (; k + 3 < K; k += 4) {
        const float4 xv = *reinterpret_cast<const float4 *>(xr + k), wv = *reinterpret_cast<const float4 *>(wr + k);
        a0 = fmaf(xv.x, wv.x, a0); a1 = fmaf(xv.y, wv.y, a1); a2 = fmaf(xv.z, wv.z, a2); a3 = fmaf(xv.w, wv.w, a3);
    }
    for (; k < K; ++k) a0 = fmaf(xr[k], wr[k], a0);
    y[idx] = (a0 + a1) + (a2 + a3) + bias[j];
}
static int sy_linear(const float *x, const float *w, const float *bias, float *y, int64_t n, int N, int K, gsb_stream_t stream) {
    if (N % 128 == 0 && K % 16 == 0) return gsb_linear_forward(x, w, bias, y, n, N, K, 0, nullptr, 0, stream);
    const int64_t total = n * N;
    sy_small_linear_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, w, bias, y, n, N, K);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

struct EpiParams {
    const float *demod;     // [nb, c]  (chunk-local rows)
    const float *noise;     // [H*W] pre-multiplied by NoiseInjection.weight
    const float *bias;      // [c]
    const float *s_next;    // [nb, c] next layer's style, or nullptr on the hooked layer
    __half *out_hi, *out_lo;   // [nb, H*W, c] (chunk-local)
    float *out_f32;         // hooked layer: row b at out_f32 + b*ld
    int64_t ld;
    unsigned *overflow;
    // ToRGB of this resolution (model.py:344-363), fused: rgb[b,pix,o] += sum_c (scale W)[o,c] s_rgb[b,c] f[b,pix,c]
    const float *rgb_w;     // [3, c] ToRGB.conv.weight (unscaled), or nullptr
    const float *rgb_s;     // [nb, c] ToRGB style (chunk-local rows)
    float *rgb_out;         // [nb, H*W, 3] (chunk-local), pre-initialised with bias + up-sampled skip
    float rgb_scale;        // 1 / sqrt(c)
};
// conv result (4 channels) -> demod, noise, bias, leaky-ReLU * sqrt2 -> next layer's operand or the fp32 activation
__device__ __forceinline__ void sy_epilogue(const EpiParams &e, float4 acc, int64_t b, int pix, int hw, int c, int q) {
    const float4 dm = *reinterpret_cast<const float4 *>(e.demod + b * c + 4 * q);
    const float4 bs = *reinterpret_cast<const float4 *>(e.bias + 4 * q);
    const float nz = e.noise[pix];
    float f[4] = {fmaf(acc.x, dm.x, nz) + bs.x, fmaf(acc.y, dm.y, nz) + bs.y, fmaf(acc.z, dm.z, nz) + bs.z,
                  fmaf(acc.w, dm.w, nz) + bs.w};
    const float sqrt2 = 1.41421356237309515f;
#pragma unroll
    for (int k = 0; k < 4; ++k) f[k] = sqrt2 * ((f[k] >= 0.f) ? f[k] : 0.2f * f[k]);
    if (e.rgb_w) {
        // the channel quads of one pixel sit in consecutive lanes (c/4 of them, a power of two; >= 32: whole warps)
        const float4 sr = *reinterpret_cast<const float4 *>(e.rgb_s + b * c + 4 * q);
        const float g[4] = {f[0] * sr.x, f[1] * sr.y, f[2] * sr.z, f[3] * sr.w};
        float r[3];
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            const float4 wv = *reinterpret_cast<const float4 *>(e.rgb_w + (int64_t)o * c + 4 * q);
            r[o] = e.rgb_scale * (g[0] * wv.x + g[1] * wv.y + g[2] * wv.z + g[3] * wv.w);
        }
        const int cq = c >> 2, span = cq < 32 ? cq : 32;
        for (int off = span >> 1; off > 0; off >>= 1) {
#pragma unroll
            for (int o = 0; o < 3; ++o) r[o] += __shfl_xor_sync(0xffffffffu, r[o], off);
        }
        float *dst = e.rgb_out + (b * hw + pix) * 3;
        if (cq <= 32) {
            // one lane per pixel holds the sum: a plain read-modify-write, deterministic
            if ((q & (span - 1)) == 0) { dst[0] += r[0]; dst[1] += r[1]; dst[2] += r[2]; }
        } else {
            // 256 / 512 channels: the pixel's quads span 2 / 4 warps of this block (blocks hold whole pixels: 256 % cq == 0 and
            // the launch has no partial blocks); combine them through shared memory in a fixed order
            __shared__ float rgb_red[8][3];
            const int wib = threadIdx.x >> 5, wpp = cq >> 5;                    // warp in block, warps per pixel
            if ((threadIdx.x & 31) == 0) { rgb_red[wib][0] = r[0]; rgb_red[wib][1] = r[1]; rgb_red[wib][2] = r[2]; }
            __syncthreads();
            if ((threadIdx.x & 31) == 0 && (wib % wpp) == 0) {
                float t0 = 0.f, t1 = 0.f, t2 = 0.f;
                for (int k = 0; k < wpp; ++k) { t0 += rgb_red[wib + k][0]; t1 += rgb_red[wib + k][1]; t2 += rgb_red[wib + k][2]; }
                dst[0] += t0; dst[1] += t1; dst[2] += t2;
            }
            __syncthreads();
        }
    }
    if (e.s_next) {
        const float4 sn = *reinterpret_cast<const float4 *>(e.s_next + b * c + 4 * q);
        f[0] *= sn.x; f[1] *= sn.y; f[2] *= sn.z; f[3] *= sn.w;
        bool ovf = false;
        store_split4(f, e.out_hi, e.out_lo, (b * hw + pix) * (int64_t)c + 4 * q, ovf);
        if (ovf) atomicOr(e.overflow, 1u);
    } else if (e.out_f32) {
        *reinterpret_cast<float4 *>(e.out_f32 + b * e.ld + (int64_t)pix * c + 4 * q) = make_float4(f[0], f[1], f[2], f[3]);
    }
}

// rgb[b, y, x, o] = bias[o] (+ Upsample(prev)[b, y, x, o]):  upfirdn2d(prev, [1,3,3,1] outer * 4 / 64, up = 2, pad = (2, 1))
// (model.py:33-51, op/upfirdn2d.py:157-198): zero-insertion puts prev[i] at 2i; out[y] = sum_i k[i] up[y + i - 2].
__global__ void sy_rgb_init_kernel(const float *__restrict__ bias, const float *__restrict__ prev, int64_t n, int R, float *__restrict__ rgb) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= n * R * R) return;
    const int x = (int)(idx % R), y = (int)((idx / R) % R);
    const int64_t b = idx / ((int64_t)R * R);
    float acc[3] = {bias[0], bias[1], bias[2]};
    if (prev) {
        const int Rp = R >> 1;
        const float k1[4] = {1.f, 3.f, 3.f, 1.f};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int u = y + i - 2;
            if (u < 0 || (u & 1) || (u >> 1) >= Rp) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int v = x + j - 2;
                if (v < 0 || (v & 1) || (v >> 1) >= Rp) continue;
                const float kw = k1[i] * k1[j] * 0.0625f;
                const float *pp = prev + ((b * Rp + (u >> 1)) * Rp + (v >> 1)) * 3;
                acc[0] = fmaf(kw, pp[0], acc[0]); acc[1] = fmaf(kw, pp[1], acc[1]); acc[2] = fmaf(kw, pp[2], acc[2]);
            }
        }
    }
    float *o = rgb + idx * 3;
    o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2];
}

// stride-1 3x3: gather the nine tap planes.  Y [nb*H*W, 9*c]
__global__ void __launch_bounds__(256)
sy_conv_gather_kernel(const float *__restrict__ Y, int64_t nb, int H, int W, int c, EpiParams e) {
    const int cq = c >> 2;
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= nb * H * W * cq) return;
    const int q = (int)(idx % cq);
    const int64_t pixg = idx / cq;
    const int x = (int)(pixg % W), y = (int)((pixg / W) % H);
    const int64_t b = pixg / ((int64_t)W * H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 1;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int xx = x + kx - 1;
            if (xx < 0 || xx >= W) continue;
            const float4 v = *reinterpret_cast<const float4 *>(Y + ((b * H + yy) * W + xx) * (int64_t)(9 * c) + (ky * 3 + kx) * c + 4 * q);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    sy_epilogue(e, acc, b, y * W + x, H * W, c, q);
}

// stride-2 transposed conv: T[b,u,v,:] on the (2H+1) x (2W+1) grid = sum of the tap planes that land on (u,v)
__global__ void __launch_bounds__(256)
sy_upconv_scatter_kernel(const float *__restrict__ Y, int64_t nb, int H, int W, int c, float *__restrict__ T) {
    const int cq = c >> 2, TH = 2 * H + 1, TW = 2 * W + 1;
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= nb * TH * TW * cq) return;
    const int q = (int)(idx % cq);
    const int64_t pixg = idx / cq;
    const int v = (int)(pixg % TW), u = (int)((pixg / TW) % TH);
    const int64_t b = pixg / ((int64_t)TW * TH);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int t = u - ky;
        if (t < 0 || (t & 1) || (t >> 1) >= H) continue;
        const int yy = t >> 1;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int s = v - kx;
            if (s < 0 || (s & 1) || (s >> 1) >= W) continue;
            const int xx = s >> 1;
            const float4 w = *reinterpret_cast<const float4 *>(Y + ((b * H + yy) * W + xx) * (int64_t)(9 * c) + (ky * 3 + kx) * c + 4 * q);
            acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
        }
    }
    *reinterpret_cast<float4 *>(T + pixg * c + 4 * q) = acc;
}

// Blur([1,3,3,1] outer / 64 * 4, pad (1,1)) of T -> 2H x 2W, then the shared epilogue
__global__ void __launch_bounds__(256)
sy_blur_epilogue_kernel(const float *__restrict__ T, int64_t nb, int H2, int W2, int c, EpiParams e) {
    const int cq = c >> 2, TH = H2 + 1, TW = W2 + 1;
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= nb * H2 * W2 * cq) return;
    const int q = (int)(idx % cq);
    const int64_t pixg = idx / cq;
    const int x = (int)(pixg % W2), y = (int)((pixg / W2) % H2);
    const int64_t b = pixg / ((int64_t)W2 * H2);
    const float k1[4] = {1.f, 3.f, 3.f, 1.f};
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int u = y + i - 1;
        if (u < 0 || u >= TH) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int v = x + j - 1;
            if (v < 0 || v >= TW) continue;
            const float kw = k1[i] * k1[j] * 0.0625f;
            const float4 t = *reinterpret_cast<const float4 *>(T + ((b * TH + u) * TW + v) * (int64_t)c + 4 * q);
            acc.x = fmaf(kw, t.x, acc.x); acc.y = fmaf(kw, t.y, acc.y); acc.z = fmaf(kw, t.z, acc.z); acc.w = fmaf(kw, t.w, acc.w);
        }
    }
    sy_epilogue(e, acc, b, y * W2 + x, H2 * W2, c, q);
}

// ---- workspace --------------------------------------------------------------------------------------------
struct SynthWs {
    float *S[SY_MAX_LAYERS], *D[SY_MAX_LAYERS];
    float *s2;
    __half *act[2][2];     // [ping-pong][hi/lo]
    float *Y, *T;
    float *rgb[2], *rgb_s, *rgb_modw;      // render path: skip images (ping-pong), ToRGB style, scaled modulation weight
    size_t bytes;
};
static int chunk_samples(const gsb_styled_conv &c) {
    int spc = SY_CHUNK_ROWS / (c.res_in * c.res_in);
    return spc < 1 ? 1 : spc;
}
static SynthWs synth_ws(void *base, const gsb_styled_conv *layers, int n_run, int64_t n, bool with_rgb = false, int style_dim = 0) {
    SynthWs w;
    char *p = reinterpret_cast<char *>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { char *q = p + off; off += align_up(bytes, 256); return q; };
    size_t cmax = 0, act_elems = 0, y_elems = 0, t_elems = 0;
    for (int l = 0; l < n_run; ++l) {
        const gsb_styled_conv &c = layers[l];
        w.S[l] = (float *)take((size_t)n * c.cin * 4);
        w.D[l] = (float *)take((size_t)n * c.cout * 4);
        cmax = cmax > (size_t)c.cin ? cmax : (size_t)c.cin;
        const size_t in_elems = (size_t)n * c.res_in * c.res_in * c.cin;
        act_elems = act_elems > in_elems ? act_elems : in_elems;
        const size_t spc = (size_t)chunk_samples(c);
        const size_t ye = spc * c.res_in * c.res_in * 9 * c.cout;
        y_elems = y_elems > ye ? y_elems : ye;
        if (c.upsample) {
            const size_t te = spc * (2 * c.res_in + 1) * (2 * c.res_in + 1) * c.cout;
            t_elems = t_elems > te ? t_elems : te;
        }
    }
    w.s2 = (float *)take((size_t)n * cmax * 4);
    for (int a = 0; a < 2; ++a)
        for (int h = 0; h < 2; ++h) w.act[a][h] = (__half *)take(act_elems * 2);
    w.Y = (float *)take(y_elems * 4);
    w.T = (float *)take((t_elems ? t_elems : 64) * 4);
    w.rgb[0] = w.rgb[1] = w.rgb_s = w.rgb_modw = nullptr;
    if (with_rgb) {
        const int ro = res_out_of(layers[n_run - 1]);
        size_t cm = 0;
        for (int l = 0; l < n_run; ++l) cm = cm > (size_t)layers[l].cout ? cm : (size_t)layers[l].cout;
        for (int a = 0; a < 2; ++a) w.rgb[a] = (float *)take((size_t)n * ro * ro * 3 * 4);
        w.rgb_s = (float *)take((size_t)n * cm * 4);
        w.rgb_modw = (float *)take(cm * (size_t)style_dim * 4);
    }
    w.bytes = off;
    return w;
}

}  // namespace gsb

extern "C" size_t gsb_synthesis_packed_bytes(const gsb_styled_conv *layers, int n_layers, int style_dim) {
    if (gsb::check_layers(layers, n_layers, style_dim)) return 0;
    return gsb::synth_view(nullptr, layers, n_layers, style_dim).bytes;
}

extern "C" int gsb_synthesis_pack(const gsb_styled_conv *layers, int n_layers, int style_dim, const float *d_const_input,
                                  void *d_packed, size_t packed_bytes, gsb_stream_t stream) {
    using namespace gsb;
    if (int r = check_layers(layers, n_layers, style_dim)) return r;
    GSB_CHECK_ARG(d_const_input && d_packed, "synthesis_pack: null pointer");
    SynthView v = synth_view(d_packed, layers, n_layers, style_dim);
    if (packed_bytes < v.bytes) { set_error("synthesis_pack: buffer too small (%zu < %zu)", packed_bytes, v.bytes); return GSB_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    GSB_CHECK_CUDA(cudaMemsetAsync(d_packed, 0, v.bytes, st));
    const int c0 = layers[0].cin;
    sy_const_nhwc_kernel<<<(16 * c0 + 255) / 256, 256, 0, st>>>(d_const_input, c0, v.const_nhwc);
    GSB_CHECK_LAUNCH();
    for (int l = 0; l < n_layers; ++l) {
        const gsb_styled_conv &c = layers[l];
        GSB_CHECK_ARG(c.conv_weight && c.mod_weight && c.mod_bias && c.act_bias && c.noise && c.noise_weight,
                      "synthesis_pack: layer %d has a null parameter pointer", l);
        const float scale = (float)(1.0 / sqrt((double)c.cin * 9.0));           // ModulatedConv2d.scale (model.py:219-220)
        const float mscale = (float)(1.0 / sqrt((double)style_dim));              // EqualLinear.scale, lr_mul = 1 (model.py:143)
        const int64_t wcount = (int64_t)c.cout * c.cin * 9;
        sy_absmax_kernel<<<128, 256, 0, st>>>(c.conv_weight, wcount, scale, v.L[l].scal + 2);
        GSB_CHECK_LAUNCH();
        sy_pick_scale_kernel<<<1, 1, 0, st>>>(v.L[l].scal);
        GSB_CHECK_LAUNCH();
        sy_weight_pack_kernel<<<256, 256, 0, st>>>(c.conv_weight, c.cout, c.cin, scale, v.L[l].scal, v.L[l].w_hi, v.L[l].w_lo, v.L[l].wsq);
        GSB_CHECK_LAUNCH();
        sy_scale_copy_kernel<<<128, 256, 0, st>>>(c.mod_weight, (int64_t)c.cin * style_dim, mscale, nullptr, v.L[l].modw);
        GSB_CHECK_LAUNCH();
        sy_scale_copy_kernel<<<4, 256, 0, st>>>(c.mod_bias, c.cin, 1.0f, nullptr, v.L[l].modb);
        GSB_CHECK_LAUNCH();
        sy_scale_copy_kernel<<<4, 256, 0, st>>>(c.act_bias, c.cout, 1.0f, nullptr, v.L[l].actb);
        GSB_CHECK_LAUNCH();
        const int ro = res_out_of(c);
        sy_scale_copy_kernel<<<64, 256, 0, st>>>(c.noise, (int64_t)ro * ro, 1.0f, c.noise_weight, v.L[l].noise);
        GSB_CHECK_LAUNCH();
    }
    return GSB_OK;
}

extern "C" size_t gsb_synthesis_workspace_bytes(const gsb_styled_conv *layers, int n_run, int64_t n) {
    if (!layers || n_run < 1 || n_run > gsb::SY_MAX_LAYERS || n < 1) return 0;
    return gsb::synth_ws(nullptr, layers, n_run, n).bytes;
}

namespace gsb {

// layers[0 .. n_run) on per-layer latents d_w [w_layers][n][style_dim] (layer l reads entry min(l, w_layers - 1)); optional fp32
// activation of the last layer (d_out), optional ToRGB chain: rgbs[j] follows layer 2j and reads latent entry 2j + 1.
static int synthesis_run(const void *d_packed, const gsb_styled_conv *layers, int n_layers, int n_run, int style_dim,
                         const gsb_to_rgb *rgbs, int n_rgb, const float *d_w, int w_layers, int64_t n, float *d_out, int64_t ld_out,
                         float *d_rgb_out, void *d_workspace, size_t workspace_bytes, gsb_stream_t stream) {
    if (int r = check_layers(layers, n_layers, style_dim)) return r;
    GSB_CHECK_ARG(d_packed && d_w && d_workspace && (d_out || d_rgb_out), "synthesis: null pointer");
    GSB_CHECK_ARG(n_run >= 1 && n_run <= n_layers && w_layers >= 1, "synthesis: n_run / w_layers out of range");
    GSB_CHECK_ARG(n_rgb >= 0 && (n_rgb == 0 || (rgbs && d_rgb_out && 2 * (n_rgb - 1) <= n_run - 1)), "synthesis: bad ToRGB list");
    if (n == 0) return GSB_OK;
    const gsb_styled_conv &last = layers[n_run - 1];
    const int ro_last = res_out_of(last);
    GSB_CHECK_ARG(n > 0 && (!d_out || (ld_out >= (int64_t)ro_last * ro_last * last.cout && ld_out % 4 == 0)), "synthesis: bad n / ld_out");
    SynthView v = synth_view(const_cast<void *>(d_packed), layers, n_layers, style_dim);
    SynthWs w = synth_ws(d_workspace, layers, n_run, n, n_rgb > 0, style_dim);
    if (workspace_bytes < w.bytes) { set_error("synthesis: workspace too small (%zu < %zu)", workspace_bytes, w.bytes); return GSB_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    auto latent = [&](int idx) { return d_w + (size_t)(idx < w_layers ? idx : w_layers - 1) * n * style_dim; };

    // styles and demodulation factors of every layer that runs (model.py:234,239)
    for (int l = 0; l < n_run; ++l) {
        const gsb_styled_conv &c = layers[l];
        if (int r = sy_linear(latent(l), v.L[l].modw, v.L[l].modb, w.S[l], n, c.cin, style_dim, stream)) return r;
        const int64_t cnt = n * c.cin;
        sy_square_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(w.S[l], cnt, w.s2);
        GSB_CHECK_LAUNCH();
        if (int r = sy_linear(w.s2, v.L[l].wsq, v.zeros, w.D[l], n, c.cout, c.cin, stream)) return r;
        const int64_t cnt2 = n * c.cout;
        sy_rsqrt_eps_kernel<<<(unsigned)((cnt2 + 255) / 256), 256, 0, st>>>(w.D[l], cnt2);
        GSB_CHECK_LAUNCH();
    }
    {   // ConstantInput * style of conv1
        const int64_t total = n * 16 * (layers[0].cin / 4);
        sy_const_modulate_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(v.const_nhwc, w.S[0], n, 16, layers[0].cin, w.act[0][0],
                                                                                w.act[0][1], v.overflow);
        GSB_CHECK_LAUNCH();
    }
    const float *rgb_prev = nullptr;
    for (int l = 0; l < n_run; ++l) {
        const gsb_styled_conv &c = layers[l];
        const int src = l & 1, dst = src ^ 1;
        const bool hooked = (l == n_run - 1);
        const int H = c.res_in, ro = res_out_of(c), hw_in = H * H, hw_out = ro * ro;
        const int spc = chunk_samples(c);
        // ToRGB after conv1 and after the second conv of every resolution (model.py:546-561)
        const int j = l / 2;
        const bool with_rgb = (l % 2 == 0) && j < n_rgb;
        float *rgb_cur = nullptr;
        if (with_rgb) {
            const gsb_to_rgb &t = rgbs[j];
            GSB_CHECK_ARG(t.conv_weight && t.mod_weight && t.mod_bias && t.bias && t.cin == c.cout && (c.cout & (c.cout - 1)) == 0,
                          "synthesis: ToRGB %d does not match layer %d (cin=%d, cout=%d)", j, l, t.cin, c.cout);
            rgb_cur = (j == n_rgb - 1) ? d_rgb_out : w.rgb[j & 1];
            const float mscale = (float)(1.0 / sqrt((double)style_dim));
            sy_scale_copy_kernel<<<64, 256, 0, st>>>(t.mod_weight, (int64_t)c.cout * style_dim, mscale, nullptr, w.rgb_modw);
            GSB_CHECK_LAUNCH();
            if (int r = sy_linear(latent(2 * j + 1), w.rgb_modw, t.mod_bias, w.rgb_s, n, c.cout, style_dim, stream)) return r;
            const int64_t tot = n * hw_out;
            sy_rgb_init_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(t.bias, rgb_prev, n, ro, rgb_cur);
            GSB_CHECK_LAUNCH();
        }
        for (int64_t b0 = 0; b0 < n; b0 += spc) {
            const int64_t nb = (b0 + spc <= n) ? spc : (n - b0);
            const int64_t rows = nb * hw_in;
            if (int r = tc_gemm_plain(w.act[src][0] + b0 * hw_in * c.cin, w.act[src][1] + b0 * hw_in * c.cin, rows, c.cin, v.L[l].w_hi,
                                      v.L[l].w_lo, 9 * c.cout, v.L[l].scal, w.Y, v.overflow, 0, st)) return r;
            EpiParams e;
            e.demod = w.D[l] + b0 * c.cout;
            e.noise = v.L[l].noise;
            e.bias = v.L[l].actb;
            e.s_next = hooked ? nullptr : w.S[l + 1] + b0 * c.cout;
            e.out_hi = hooked ? nullptr : w.act[dst][0] + b0 * hw_out * c.cout;
            e.out_lo = hooked ? nullptr : w.act[dst][1] + b0 * hw_out * c.cout;
            e.out_f32 = (hooked && d_out) ? d_out + b0 * ld_out : nullptr;
            e.ld = ld_out;
            e.overflow = v.overflow;
            e.rgb_w = with_rgb ? rgbs[j].conv_weight : nullptr;
            e.rgb_s = with_rgb ? w.rgb_s + b0 * c.cout : nullptr;
            e.rgb_out = with_rgb ? rgb_cur + b0 * hw_out * 3 : nullptr;
            e.rgb_scale = (float)(1.0 / sqrt((double)c.cout));
            const int cq = c.cout / 4;
            if (c.upsample) {
                const int64_t t_total = nb * (2 * H + 1) * (2 * H + 1) * cq;
                sy_upconv_scatter_kernel<<<(unsigned)((t_total + 255) / 256), 256, 0, st>>>(w.Y, nb, H, H, c.cout, w.T);
                GSB_CHECK_LAUNCH();
                const int64_t o_total = nb * hw_out * cq;
                sy_blur_epilogue_kernel<<<(unsigned)((o_total + 255) / 256), 256, 0, st>>>(w.T, nb, ro, ro, c.cout, e);
                GSB_CHECK_LAUNCH();
            } else {
                const int64_t o_total = nb * hw_out * cq;
                sy_conv_gather_kernel<<<(unsigned)((o_total + 255) / 256), 256, 0, st>>>(w.Y, nb, H, H, c.cout, e);
                GSB_CHECK_LAUNCH();
            }
        }
        if (with_rgb) rgb_prev = rgb_cur;
    }
    return GSB_OK;
}

}  // namespace gsb

extern "C" int gsb_synthesis_forward(const void *d_packed, const gsb_styled_conv *layers, int n_layers, int n_run, int style_dim,
                                     const float *d_w, int64_t n, float *d_out, int64_t ld_out, void *d_workspace,
                                     size_t workspace_bytes, gsb_stream_t stream) {
    GSB_CHECK_ARG(d_out, "synthesis_forward: null output");
    return gsb::synthesis_run(d_packed, layers, n_layers, n_run, style_dim, nullptr, 0, d_w, 1, n, d_out, ld_out, nullptr, d_workspace,
                              workspace_bytes, stream);
}

extern "C" size_t gsb_synthesis_render_workspace_bytes(const gsb_styled_conv *layers, int n_run, int64_t n, int style_dim) {
    if (!layers || n_run < 1 || n_run > gsb::SY_MAX_LAYERS || n < 1) return 0;
    return gsb::synth_ws(nullptr, layers, n_run, n, true, style_dim).bytes;
}

// Render path (Generator.forward, model.py:493-571): per-layer latents + the ToRGB / skip chain.
extern "C" int gsb_synthesis_render(const void *d_packed, const gsb_styled_conv *layers, int n_layers, int n_run, int style_dim,
                                    const gsb_to_rgb *rgbs, int n_rgb, const float *d_w, int w_layers, int64_t n, float *d_act_out,
                                    int64_t ld_act, float *d_rgb_out, void *d_workspace, size_t workspace_bytes, gsb_stream_t stream) {
    return gsb::synthesis_run(d_packed, layers, n_layers, n_run, style_dim, rgbs, n_rgb, d_w, w_layers, n, d_act_out, ld_act, d_rgb_out,
                              d_workspace, workspace_bytes, stream);
}

extern "C" int gsb_synthesis_status(const void *d_packed, const gsb_styled_conv *layers, int n_layers, int style_dim, unsigned *h_flags) {
    using namespace gsb;
    if (int r = check_layers(layers, n_layers, style_dim)) return r;
    GSB_CHECK_ARG(d_packed && h_flags, "synthesis_status: null pointer");
    SynthView v = synth_view(const_cast<void *>(d_packed), layers, n_layers, style_dim);
    GSB_CHECK_CUDA(cudaMemcpy(h_flags, v.overflow, sizeof(unsigned), cudaMemcpyDeviceToHost));
    return GSB_OK;
}
