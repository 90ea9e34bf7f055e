// Incremental-PCA chain, small-d engine (d <= 1024): all state and arithmetic fp64, on device.
//
// Replaces estimators.py:55-81 (IPCAEstimator.fit_partial / get_components), i.e. scikit-learn's
// IncrementalPCA.partial_fit (_incremental_pca.py:254-380) in its Gram form:
//     G = V^T S^2 V + Xc^T Xc + m m^T,   m = sqrt(n_seen*n_b/n_tot) (mean - mean_b)
//     top-c eigenpairs of G  ->  components_ (svd_flip sign rule), singular_values_ = sqrt(lambda)
//     mean/var merge of extmath._incremental_mean_and_var (Chan et al.), batch variance = diag(Xc^T Xc)
//
// The symmetric eigensolver is the classical direct route, written for one B200:
//   1. tridiag_kernel   Householder tridiagonalisation spread over P = n/8 CTAs (column-cyclic, each CTA's
//                       columns resident in its shared memory), ONE grid barrier per reflector: the fused
//                       pass applies the pending rank-2 update, accumulates A v for the next reflector and
//                       extracts the next pivot row; every CTA rebuilds v / w redundantly from the two
//                       exchanged n-vectors, so nothing else crosses SMs.
//   2. bisect_kernel    top-c eigenvalues of T by 32-way multisection (one warp per eigenvalue, Sturm counts).
//   3. invit_kernel     eigenvectors of T by inverse iteration on the pivoted LU of T - lambda I.
//   4. backtransform_kernel  applies the reflectors (one warp per eigenvector) and the sign rule.
#include "common.cuh"
#include <math.h>

namespace gsb {

constexpr int ST_HDR = 4;   // state header doubles: [0]=n_seen, [1]=steps, [2..3] reserved

struct StateView {
    double *hdr, *mean, *unnorm, *S, *V;
};
__host__ __device__ inline StateView state_view(void *p, int d, int c) {
    StateView s;
    s.hdr = reinterpret_cast<double *>(p);
    s.mean = s.hdr + ST_HDR;
    s.unnorm = s.mean + d;
    s.S = s.unnorm + d;
    s.V = s.S + c;
    return s;
}

struct Workspace {
    double *A, *dg, *e, *beta, *Vh, *lam, *Z, *lu, *xch, *evecs;
    unsigned *counter;
    unsigned char *swp;
    size_t bytes;
};
static Workspace carve(void *base, int d, int c) {
    Workspace w;
    char *p = reinterpret_cast<char *>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { char *q = p + off; off += align_up(bytes, 256); return q; };
    w.A = (double *)take((size_t)d * d * 8);
    w.dg = (double *)take((size_t)d * 8);
    w.e = (double *)take((size_t)d * 8);
    w.beta = (double *)take((size_t)d * 8);
    w.Vh = (double *)take((size_t)d * d * 8);
    w.lam = (double *)take((size_t)c * 8);
    w.Z = (double *)take((size_t)c * d * 8);
    w.evecs = (double *)take((size_t)c * d * 8);
    w.lu = (double *)take((size_t)5 * d * c * 8);
    w.swp = (unsigned char *)take((size_t)d * c);
    w.xch = (double *)take((size_t)4 * d * 8);
    w.counter = (unsigned *)take(256);
    w.bytes = off;
    return w;
}

// ---------------------------------------------------------------------------------------------
// G = gram_b + m m^T + sum_t S_t^2 v_t v_t^T      (first step: G = gram_b)
// ---------------------------------------------------------------------------------------------
constexpr int BG_T = 32;
__global__ void build_g_kernel(const double *__restrict__ gram_b, const double *__restrict__ mean_b,
                               const double *__restrict__ mean, const double *__restrict__ S,
                               const double *__restrict__ V, int d, int c, double n_seen, double n_b,
                               double *__restrict__ G) {
    __shared__ double Vi[BG_T][33], Vj[BG_T][33], s2[BG_T];
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    const int j = blockIdx.x * 32 + tx;
    const int i0 = blockIdx.y * 32;
    double acc[4] = {0, 0, 0, 0};
    if (n_seen > 0) {
        for (int t0 = 0; t0 < c; t0 += BG_T) {
            for (int tt = ty; tt < BG_T; tt += 8) {
                int t = t0 + tt;
                bool ok = t < c;
                Vi[tt][tx] = (ok && i0 + tx < d) ? V[(size_t)t * d + i0 + tx] : 0.0;
                Vj[tt][tx] = (ok && j < d) ? V[(size_t)t * d + j] : 0.0;
                if (tx == 0) { double s = ok ? S[t] : 0.0; s2[tt] = s * s; }
            }
            __syncthreads();
#pragma unroll 8
            for (int tt = 0; tt < BG_T; ++tt) {
                double vj = Vj[tt][tx] * s2[tt];
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r] += Vi[tt][ty + 8 * r] * vj;
            }
            __syncthreads();
        }
    }
    if (j >= d) return;
    const double f = (n_seen > 0) ? sqrt((n_seen / (n_seen + n_b)) * n_b) : 0.0;
    const double mj = (n_seen > 0) ? f * (mean[j] - mean_b[j]) : 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int i = i0 + ty + 8 * r;
        if (i >= d) continue;
        double mi = (n_seen > 0) ? f * (mean[i] - mean_b[i]) : 0.0;
        G[(size_t)i * d + j] = gram_b[(size_t)i * d + j] + mi * mj + acc[r];
    }
}

// ---------------------------------------------------------------------------------------------
// grid barrier (all CTAs of the launch are co-resident: grid <= #SMs, 1 CTA each fits trivially)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void grid_barrier(unsigned *counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (ld_acquire_u32(counter) < target) { }
        __threadfence();
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Householder tridiagonalisation  A = Q T Q^T,  Q = H_0 H_1 ... H_{n-3},  H_k = I - beta_k v_k v_k^T
// ---------------------------------------------------------------------------------------------
constexpr int TRI_THREADS = 256;

__global__ void __launch_bounds__(TRI_THREADS, 1)
tridiag_kernel(const double *__restrict__ A, int n, double *__restrict__ dg, double *__restrict__ e,
               double *__restrict__ beta, double *__restrict__ Vh, double *__restrict__ xch,
               unsigned *__restrict__ counter) {
    extern __shared__ double smd[];
    const int P = gridDim.x, me = blockIdx.x, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = TRI_THREADS / 32;
    const int ncl = n / P;
    double *Aloc = smd;               // [ncl][n]   owned columns j = me + P*l
    double *a = Aloc + (size_t)ncl * n;   // current pivot column (rows > k valid)
    double *v = a + n;
    double *w = v + n;
    double *pv = w + n;               // pending rank-2 update (v_{k-1}, w_{k-1})
    double *pw = pv + n;
    double *red = pw + n;             // [64]

    for (int l = 0; l < ncl; ++l) {
        const double *src = A + (size_t)(me + P * l) * n;   // row j == column j (symmetric)
        for (int i = tid; i < n; i += TRI_THREADS) Aloc[(size_t)l * n + i] = src[i];
    }
    for (int i = tid; i < n; i += TRI_THREADS) {
        a[i] = A[i];
        pv[i] = 0.0;
        pw[i] = 0.0;
        w[i] = 0.0;
    }
    if (me == 0 && tid == 0) dg[0] = A[0];
    __syncthreads();

    unsigned target = 0;
    for (int k = 0; k <= n - 3; ++k) {
        const int par = k & 1;
        double *Pbuf = xch + (size_t)par * 2 * n, *Rbuf = Pbuf + n;
        // ---- 1. reflector from a[k+1 .. n-1] (redundant in every CTA) ---------------------------
        const double x0 = a[k + 1];
        double part = 0.0;
        for (int i = k + 2 + tid; i < n; i += TRI_THREADS) part += a[i] * a[i];
        const double sigma = block_sum(part, red);
        double alpha, bk, v0;
        if (sigma == 0.0) {
            alpha = x0; bk = 0.0; v0 = 0.0;
        } else {
            const double nrm = sqrt(x0 * x0 + sigma);
            alpha = (x0 > 0.0) ? -nrm : nrm;
            v0 = x0 - alpha;
            bk = 1.0 / (nrm * (nrm + fabs(x0)));     // 2 / (v^T v)
        }
        for (int i = tid; i < n; i += TRI_THREADS)
            v[i] = (i <= k || bk == 0.0) ? 0.0 : ((i == k + 1) ? v0 : a[i]);
        __syncthreads();
        if (me == 0) {
            if (tid == 0) { e[k] = alpha; beta[k] = bk; }
            for (int i = tid; i < n; i += TRI_THREADS) Vh[(size_t)k * n + i] = v[i];
        }
        // ---- 2. fused local pass: pending update, p = A v, next pivot row ------------------------
        for (int l = warp; l < ncl; l += nwarps) {
            const int j = me + P * l;
            if (j <= k) continue;
            double *col = Aloc + (size_t)l * n;
            const double pvj = pv[j], pwj = pw[j];
            double acc = 0.0, rj = 0.0;
            for (int i = k + 1 + lane; i < n; i += 32) {
                double x = col[i] - pv[i] * pwj - pw[i] * pvj;
                col[i] = x;
                acc += x * v[i];
                if (i == k + 1) rj = x;
            }
            acc = warp_sum(acc);
            if (lane == 0) {
                __stcg(&Pbuf[j], bk * acc);
                __stcg(&Rbuf[j], rj);
            }
        }
        // ---- 3. exchange ---------------------------------------------------------------------
        target += (unsigned)P;
        grid_barrier(counter, target);
        // ---- 4. w, next pivot column (redundant in every CTA) -------------------------------------
        part = 0.0;
        for (int i = k + 1 + tid; i < n; i += TRI_THREADS) {
            double pi = __ldcg(&Pbuf[i]);
            w[i] = pi;
            a[i] = __ldcg(&Rbuf[i]);
            part += pi * v[i];
        }
        const double ptv = block_sum(part, red);
        const double K2 = 0.5 * bk * ptv;
        for (int i = k + 1 + tid; i < n; i += TRI_THREADS) w[i] -= K2 * v[i];
        __syncthreads();
        const double vk1 = v[k + 1], wk1 = w[k + 1];
        for (int i = k + 1 + tid; i < n; i += TRI_THREADS) a[i] -= vk1 * w[i] + wk1 * v[i];
        __syncthreads();
        if (me == 0 && tid == 0) dg[k + 1] = a[k + 1];
        double *t = pv; pv = v; v = t;
        t = pw; pw = w; w = t;
    }
    // last 2x2 block: e[n-2] = A[n-1,n-2] (held in a[n-1]); dg[n-1] needs the pending update
    if (me == 0 && tid == 0) { e[n - 2] = a[n - 1]; e[n - 1] = 0.0; beta[n - 2] = 0.0; beta[n - 1] = 0.0; }
    if (me == (n - 1) % P && tid == 0) {
        int l = (n - 1) / P;
        dg[n - 1] = Aloc[(size_t)l * n + (n - 1)] - 2.0 * pv[n - 1] * pw[n - 1];
    }
}

// ---------------------------------------------------------------------------------------------
// top-c eigenvalues of the tridiagonal T by 32-way multisection on Sturm counts
// ---------------------------------------------------------------------------------------------
constexpr int BIS_THREADS = 256;
__global__ void __launch_bounds__(BIS_THREADS)
bisect_kernel(const double *__restrict__ dg, const double *__restrict__ e, int n, int c,
              double *__restrict__ lam) {
    extern __shared__ double smd[];
    double *sd = smd, *se2 = smd + n, *red = se2 + n;   // red[64]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double gl = 1e300, gu = -1e300, emax = 0.0;
    for (int i = tid; i < n; i += BIS_THREADS) {
        double di = dg[i];
        double el = (i > 0) ? fabs(e[i - 1]) : 0.0, er = (i < n - 1) ? fabs(e[i]) : 0.0;
        sd[i] = di;
        se2[i] = (i < n - 1) ? e[i] * e[i] : 0.0;
        gl = fmin(gl, di - el - er);
        gu = fmax(gu, di + el + er);
        emax = fmax(emax, er * er);
    }
    // block min / max via shared scratch
    for (int o = 16; o > 0; o >>= 1) {
        gl = fmin(gl, __shfl_xor_sync(0xffffffffu, gl, o));
        gu = fmax(gu, __shfl_xor_sync(0xffffffffu, gu, o));
        emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    }
    if (lane == 0) { red[warp] = gl; red[8 + warp] = gu; red[16 + warp] = emax; }
    __syncthreads();
    gl = red[0]; gu = red[8]; emax = red[16];
    for (int q = 1; q < BIS_THREADS / 32; ++q) {
        gl = fmin(gl, red[q]); gu = fmax(gu, red[8 + q]); emax = fmax(emax, red[16 + q]);
    }
    const double eps = 2.220446049250313e-16, safemin = 2.2250738585072014e-308;
    const double pivmin = safemin * fmax(1.0, emax);
    const double tnorm = fmax(fabs(gl), fabs(gu));
    const double margin = 2.0 * tnorm * eps * n + 2.0 * pivmin;
    const int t = blockIdx.x * (BIS_THREADS / 32) + warp;   // t-th largest
    if (t >= c) return;
    const int m = n - 1 - t;                                 // ascending index
    double lo = gl - margin, hi = gu + margin;
    for (int it = 0; it < 16; ++it) {
        const double width = hi - lo;
        const double x = lo + width * ((double)(lane + 1) / 33.0);
        // Sturm count: number of eigenvalues < x
        int cnt = 0;
        double q = sd[0] - x;
        if (fabs(q) <= pivmin) q = -pivmin;
        cnt += (q < 0.0);
        for (int i = 1; i < n; ++i) {
            q = sd[i] - x - se2[i - 1] / q;
            if (fabs(q) <= pivmin) q = -pivmin;
            cnt += (q < 0.0);
        }
        unsigned mask = __ballot_sync(0xffffffffu, cnt >= m + 1);
        int f = mask ? (__ffs(mask) - 1) : 32;
        double xhi = __shfl_sync(0xffffffffu, x, f & 31);
        double xlo = __shfl_sync(0xffffffffu, x, (f > 0 ? f - 1 : 0));
        double nhi = (f < 32) ? xhi : hi;
        double nlo = (f > 0) ? xlo : lo;
        hi = nhi; lo = nlo;
        if (hi - lo <= 2.0 * eps * fmax(fabs(lo), fabs(hi)) + 2.0 * pivmin || hi - lo >= width) break;
    }
    if (lane == 0) lam[t] = 0.5 * (lo + hi);
}

// ---------------------------------------------------------------------------------------------
// eigenvectors of T: inverse iteration on the partially pivoted LU of T - lambda I (one thread each)
// scratch arrays are [i][t] so that the threads of a warp touch consecutive addresses
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double hash_unit(unsigned i, unsigned t) {
    unsigned h = i * 2654435761u ^ (t + 1u) * 40503u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    return ((double)(h & 0xffffffu) / 8388608.0) - 1.0;   // [-1, 1)
}

__global__ void invit_kernel(const double *__restrict__ dg, const double *__restrict__ e,
                             const double *__restrict__ lam, int n, int c, double *__restrict__ lu,
                             unsigned char *__restrict__ swp, double *__restrict__ Z) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= c) return;
    const size_t cs = (size_t)c;
    double *u0 = lu, *u1 = lu + (size_t)n * cs, *u2 = lu + 2 * (size_t)n * cs, *ml = lu + 3 * (size_t)n * cs,
           *xb = lu + 4 * (size_t)n * cs;
    const double lambda = lam[t];
    // scale for tiny pivots: eps * ||T||_1-ish
    double tn = 0.0;
    for (int i = 0; i < n; ++i) tn = fmax(tn, fabs(dg[i]) + ((i < n - 1) ? fabs(e[i]) : 0.0) + ((i > 0) ? fabs(e[i - 1]) : 0.0));
    const double tiny = fmax(2.220446049250313e-16 * tn, 1e-300);

    double p = dg[0] - lambda, q = (n > 1) ? e[0] : 0.0;
    for (int i = 0; i < n - 1; ++i) {
        const double sub = e[i];
        const double dn = dg[i + 1] - lambda;
        const double sn = (i + 1 < n - 1) ? e[i + 1] : 0.0;
        const size_t o = (size_t)i * cs + t;
        if (fabs(p) >= fabs(sub)) {
            if (fabs(p) < tiny) p = (p < 0.0) ? -tiny : tiny;
            const double mult = sub / p;
            u0[o] = p; u1[o] = q; u2[o] = 0.0; ml[o] = mult; swp[o] = 0;
            p = dn - mult * q;
            q = sn;
        } else {
            const double mult = p / sub;
            u0[o] = sub; u1[o] = dn; u2[o] = sn; ml[o] = mult; swp[o] = 1;
            p = q - mult * dn;
            q = -mult * sn;
        }
    }
    if (fabs(p) < tiny) p = (p < 0.0) ? -tiny : tiny;
    u0[(size_t)(n - 1) * cs + t] = p;

    for (int i = 0; i < n; ++i) xb[(size_t)i * cs + t] = hash_unit((unsigned)i, (unsigned)t);
    for (int iter = 0; iter < 3; ++iter) {
        // forward: apply the row operations to b
        double bi = xb[t];
        for (int i = 0; i < n - 1; ++i) {
            const size_t o = (size_t)i * cs + t;
            double bn = xb[o + cs];
            if (swp[o]) { double tmp = bi; bi = bn; bn = tmp; }
            xb[o] = bi;
            bi = bn - ml[o] * bi;
        }
        xb[(size_t)(n - 1) * cs + t] = bi;
        // backward
        double x2 = 0.0, x1 = 0.0, amax = 0.0;
        for (int i = n - 1; i >= 0; --i) {
            const size_t o = (size_t)i * cs + t;
            double r = xb[o];
            if (i < n - 1) r -= u1[o] * x1;
            if (i < n - 2) r -= u2[o] * x2;
            double x = r / u0[o];
            xb[o] = x;
            x2 = x1; x1 = x;
            amax = fmax(amax, fabs(x));
        }
        // normalise: max-abs first (overflow guard), then 2-norm
        double inv = 1.0 / amax, ss = 0.0;
        for (int i = 0; i < n; ++i) { double x = xb[(size_t)i * cs + t] * inv; ss += x * x; }
        inv = inv / sqrt(ss);
        for (int i = 0; i < n; ++i) xb[(size_t)i * cs + t] *= inv;
    }
    for (int i = 0; i < n; ++i) Z[(size_t)t * n + i] = xb[(size_t)i * cs + t];
}

// ---------------------------------------------------------------------------------------------
// Re-orthogonalise eigenvectors of (numerically) repeated eigenvalues.  Inverse iteration gives
// orthogonality ~ eps*||T||/gap, so only clusters with gaps below 1e-7*||T|| need it (LAPACK dstein
// uses 1e-3; with distinct eigenvalues -- every GAN activation spectrum seen here -- this kernel
// finds no cluster and returns after one pass over lam).  Classical Gram-Schmidt applied twice.
// ---------------------------------------------------------------------------------------------
constexpr int CO_THREADS = 1024;
__global__ void __launch_bounds__(CO_THREADS)
cluster_orth_kernel(const double *__restrict__ lam, const double *__restrict__ dg, const double *__restrict__ e,
                    int n, int c, double *__restrict__ Z) {
    extern __shared__ double smd[];
    double *zt = smd;            // [n]
    double *dots = zt + n;       // [c]
    double *red = dots + c;      // [64]
    int *start = reinterpret_cast<int *>(red + 64);   // [c]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = CO_THREADS / 32;
    double tn = 0.0;
    for (int i = tid; i < n; i += CO_THREADS)
        tn = fmax(tn, fabs(dg[i]) + ((i < n - 1) ? fabs(e[i]) : 0.0) + ((i > 0) ? fabs(e[i - 1]) : 0.0));
    for (int o = 16; o > 0; o >>= 1) tn = fmax(tn, __shfl_xor_sync(0xffffffffu, tn, o));
    if (lane == 0) red[warp] = tn;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int q = 0; q < nw; ++q) t = fmax(t, red[q]);
        const double tol = 1e-7 * t;
        int any = 0;
        start[0] = 0;
        for (int k = 1; k < c; ++k) {
            start[k] = (fabs(lam[k - 1] - lam[k]) <= tol) ? start[k - 1] : k;
            any |= (start[k] != k);
        }
        red[32] = (double)any;
    }
    __syncthreads();
    if (red[32] == 0.0) return;
    for (int t = 0; t < c; ++t) {
        const int s0 = start[t];
        if (s0 == t) continue;
        for (int pass = 0; pass < 2; ++pass) {
            for (int i = tid; i < n; i += CO_THREADS) zt[i] = Z[(size_t)t * n + i];
            __syncthreads();
            for (int s = s0 + warp; s < t; s += nw) {
                double d = 0.0;
                for (int i = lane; i < n; i += 32) d += Z[(size_t)s * n + i] * zt[i];
                d = warp_sum(d);
                if (lane == 0) dots[s] = d;
            }
            __syncthreads();
            double nrm = 0.0;
            for (int i = tid; i < n; i += CO_THREADS) {
                double x = zt[i];
                for (int s = s0; s < t; ++s) x -= dots[s] * Z[(size_t)s * n + i];
                zt[i] = x;
                nrm += x * x;
            }
            nrm = block_sum(nrm, red);
            const double inv = (nrm > 0.0) ? 1.0 / sqrt(nrm) : 0.0;
            for (int i = tid; i < n; i += CO_THREADS) Z[(size_t)t * n + i] = zt[i] * inv;
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// eigenvectors of A:  x = H_0 H_1 ... H_{n-3} z ; then the svd_flip sign rule (largest |.| entry > 0)
// ---------------------------------------------------------------------------------------------
constexpr int BT_WARPS = 4;
__global__ void __launch_bounds__(BT_WARPS * 32)
backtransform_kernel(const double *__restrict__ Z, const double *__restrict__ Vh,
                     const double *__restrict__ beta, int n, int c, double *__restrict__ out) {
    extern __shared__ double smd[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x * BT_WARPS + warp;
    if (t >= c) return;
    double *z = smd + (size_t)warp * n;
    for (int i = lane; i < n; i += 32) z[i] = Z[(size_t)t * n + i];
    __syncwarp();
    for (int k = n - 3; k >= 0; --k) {
        const double bk = beta[k];
        if (bk == 0.0) continue;
        const double *vk = Vh + (size_t)k * n;
        double s = 0.0;
        for (int i = k + 1 + lane; i < n; i += 32) s += vk[i] * z[i];
        s = warp_sum(s) * bk;
        for (int i = k + 1 + lane; i < n; i += 32) z[i] -= s * vk[i];
        __syncwarp();
    }
    // argmax |z| (first index on ties, as np.argmax)
    double best = -1.0;
    int bi = 0;
    for (int i = lane; i < n; i += 32) {
        double az = fabs(z[i]);
        if (az > best) { best = az; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        double ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    const double sgn = (z[bi] < 0.0) ? -1.0 : 1.0;
    for (int i = lane; i < n; i += 32) out[(size_t)t * n + i] = sgn * z[i];
}

// ---------------------------------------------------------------------------------------------
// state update after the eigensolve
// ---------------------------------------------------------------------------------------------
__global__ void finalize_kernel(double *hdr, double *mean, double *unnorm, double *S, double *V,
                                const double *__restrict__ mean_b, const double *__restrict__ gram_b,
                                const double *__restrict__ lam, const double *__restrict__ evecs, int d,
                                int c, double n_seen, double n_b) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const double n_tot = n_seen + n_b;
    if (idx < (size_t)c * d) V[idx] = evecs[idx];
    if (idx < (size_t)c) S[idx] = sqrt(fmax(lam[idx], 0.0));
    if (idx < (size_t)d) {
        const double mb = mean_b[idx], vb = gram_b[idx * (size_t)d + idx];
        if (n_seen > 0) {
            const double mo = mean[idx];
            // extmath._incremental_mean_and_var: updated_mean = (last_sum + new_sum) / updated_count
            mean[idx] = (mo * n_seen + mb * n_b) / n_tot;
            // last_unnorm + new_unnorm + last_over_new/updated * (last_sum/last_over_new - new_sum)^2
            const double r = n_seen / n_b;
            const double tq = (mo * n_seen) / r - mb * n_b;
            unnorm[idx] = unnorm[idx] + vb + r / n_tot * tq * tq;
        } else {
            mean[idx] = mb;
            unnorm[idx] = vb;
        }
    }
    if (idx == 0) { hdr[0] = n_tot; hdr[1] += 1.0; }
}

__global__ void export_kernel(const double *hdr, const double *mean, const double *unnorm, const double *S,
                              const double *V, int d, int c, double n_seen, double *o_comp, double *o_sv,
                              double *o_mean, double *o_var, double *o_ev, double *o_evr) {
    __shared__ double red[64];
    double part = 0.0;
    for (int i = threadIdx.x; i < d; i += blockDim.x) part += unnorm[i];
    const double tot = block_sum(part, red);
    for (size_t i = threadIdx.x; i < (size_t)c * d; i += blockDim.x)
        if (o_comp) o_comp[i] = V[i];
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        if (o_mean) o_mean[i] = mean[i];
        if (o_var) o_var[i] = unnorm[i] / n_seen;
    }
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        const double s = S[i];
        if (o_sv) o_sv[i] = s;
        if (o_ev) o_ev[i] = s * s / (n_seen - 1.0);
        if (o_evr) o_evr[i] = s * s / tot;
    }
}

// ---------------------------------------------------------------------------------------------
static int eig_top(const Workspace &w, int d, int c, double *evals, double *evecs, cudaStream_t st) {
    // P CTAs, 8 columns each (n % 8 == 0); all must be co-resident for the grid barrier
    int P = d / 8;
    while (P > 128) P /= 2;
    GSB_CHECK_ARG(d % P == 0, "sym_eig: d=%d not divisible by P=%d", d, P);
    const int ncl = d / P;
    const size_t tri_smem = ((size_t)ncl * d + 5 * (size_t)d + 64) * sizeof(double);
    GSB_CHECK_ARG(tri_smem <= 200 * 1024, "sym_eig: d=%d too large for the small-d engine", d);
    static size_t tri_smem_set = 0;
    if (tri_smem > tri_smem_set) {
        GSB_CHECK_CUDA(cudaFuncSetAttribute(tridiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)tri_smem));
        tri_smem_set = tri_smem;
    }
    GSB_CHECK_CUDA(cudaMemsetAsync(w.counter, 0, 256, st));
    tridiag_kernel<<<P, TRI_THREADS, tri_smem, st>>>(w.A, d, w.dg, w.e, w.beta, w.Vh, w.xch, w.counter);
    GSB_CHECK_LAUNCH();
    const size_t bis_smem = (2 * (size_t)d + 64) * sizeof(double);
    bisect_kernel<<<(c + 7) / 8, BIS_THREADS, bis_smem, st>>>(w.dg, w.e, d, c, evals);
    GSB_CHECK_LAUNCH();
    invit_kernel<<<(c + 31) / 32, 32, 0, st>>>(w.dg, w.e, evals, d, c, w.lu, w.swp, w.Z);
    GSB_CHECK_LAUNCH();
    const size_t co_smem = ((size_t)d + c + 64) * sizeof(double) + (size_t)c * sizeof(int);
    cluster_orth_kernel<<<1, CO_THREADS, co_smem, st>>>(evals, w.dg, w.e, d, c, w.Z);
    GSB_CHECK_LAUNCH();
    const size_t bt_smem = (size_t)BT_WARPS * d * sizeof(double);
    backtransform_kernel<<<(c + BT_WARPS - 1) / BT_WARPS, BT_WARPS * 32, bt_smem, st>>>(w.Z, w.Vh, w.beta, d, c,
                                                                                      evecs);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

static int check_dims(int d, int c) {
    GSB_CHECK_ARG(d >= 32 && d <= 1024 && d % 32 == 0, "ipca: small-d engine needs 32 <= d <= 1024, d%%32==0 (d=%d)", d);
    GSB_CHECK_ARG(c >= 1 && c <= d, "ipca: need 1 <= c <= d (c=%d d=%d)", c, d);
    return GSB_OK;
}

}  // namespace gsb

extern "C" size_t gsb_ipca_state_bytes(int d, int c) {
    return (size_t)(gsb::ST_HDR + 2 * (size_t)d + c + (size_t)c * d) * sizeof(double);
}

extern "C" size_t gsb_ipca_workspace_bytes(int d, int c) {
    return gsb::carve(nullptr, d, c).bytes;
}

extern "C" int gsb_ipca_reset(void *d_state, int d, int c, gsb_stream_t stream) {
    GSB_CHECK_ARG(d_state, "ipca_reset: null state");
    if (int r = gsb::check_dims(d, c)) return r;
    GSB_CHECK_CUDA(cudaMemsetAsync(d_state, 0, gsb_ipca_state_bytes(d, c), (cudaStream_t)stream));
    return GSB_OK;
}

extern "C" int gsb_ipca_chain_step(void *d_state, int d, int c, int64_t n_seen, int64_t n_batch,
                                   const double *d_mean_b, const double *d_gram_b, void *d_workspace,
                                   size_t workspace_bytes, gsb_stream_t stream) {
    GSB_CHECK_ARG(d_state && d_mean_b && d_gram_b && d_workspace, "ipca_chain_step: null pointer");
    if (int r = gsb::check_dims(d, c)) return r;
    GSB_CHECK_ARG(n_seen >= 0 && n_batch > 0, "ipca_chain_step: bad counts");
    // sklearn: "n_components must be <= the batch number of samples for the first partial_fit call"
    GSB_CHECK_ARG(n_seen > 0 || c <= n_batch, "ipca_chain_step: n_components=%d > first batch size %lld", c,
                  (long long)n_batch);
    gsb::Workspace w = gsb::carve(d_workspace, d, c);
    if (workspace_bytes < w.bytes) {
        gsb::set_error("ipca_chain_step: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
        return GSB_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    gsb::StateView s = gsb::state_view(d_state, d, c);
    dim3 grid((d + 31) / 32, (d + 31) / 32), block(32, 8);
    gsb::build_g_kernel<<<grid, block, 0, st>>>(d_gram_b, d_mean_b, s.mean, s.S, s.V, d, c, (double)n_seen,
                                                (double)n_batch, w.A);
    GSB_CHECK_LAUNCH();
    if (int r = gsb::eig_top(w, d, c, w.lam, w.evecs, st)) return r;
    size_t tot = (size_t)c * d;
    gsb::finalize_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(
        s.hdr, s.mean, s.unnorm, s.S, s.V, d_mean_b, d_gram_b, w.lam, w.evecs, d, c, (double)n_seen,
        (double)n_batch);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

extern "C" int gsb_ipca_export(const void *d_state, int d, int c, int64_t n_seen, double *d_components,
                               double *d_singular_values, double *d_mean, double *d_var,
                               double *d_explained_variance, double *d_explained_variance_ratio,
                               gsb_stream_t stream) {
    GSB_CHECK_ARG(d_state, "ipca_export: null state");
    if (int r = gsb::check_dims(d, c)) return r;
    GSB_CHECK_ARG(n_seen > 1, "ipca_export: nothing fitted yet");
    gsb::StateView s = gsb::state_view(const_cast<void *>(d_state), d, c);
    gsb::export_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(s.hdr, s.mean, s.unnorm, s.S, s.V, d, c,
                                                            (double)n_seen, d_components, d_singular_values,
                                                            d_mean, d_var, d_explained_variance,
                                                            d_explained_variance_ratio);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

extern "C" int gsb_sym_eig_top(double *d_a, int d, int c, double *d_evals, double *d_evecs,
                               void *d_workspace, size_t workspace_bytes, gsb_stream_t stream) {
    GSB_CHECK_ARG(d_a && d_evals && d_evecs && d_workspace, "sym_eig_top: null pointer");
    if (int r = gsb::check_dims(d, c)) return r;
    gsb::Workspace w = gsb::carve(d_workspace, d, c);
    if (workspace_bytes < w.bytes) {
        gsb::set_error("sym_eig_top: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
        return GSB_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    GSB_CHECK_CUDA(cudaMemcpyAsync(w.A, d_a, (size_t)d * d * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return gsb::eig_top(w, d, c, d_evals, d_evecs, st);
}
