// fp64 / shuffle / cluster-exchange micro-benchmarks on one B200 (numbers quoted in DESIGN.md section 5b).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void dfma_chain(double *out, int iters, long long *clk) {
    double a = out[0], b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < iters; ++i) a = fma(a, b, c);
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void ffma_chain(float *out, int iters, long long *clk) {
    float a = out[0], b = 1.0000001f, c = 1e-9f;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < iters; ++i) a = fmaf(a, b, c);
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
// ILP independent chains per thread
template <int ILP>
__global__ void dfma_tput(double *out, int iters, long long *clk) {
    double a[ILP];
#pragma unroll
    for (int q = 0; q < ILP; ++q) a[q] = out[q] + q;
    const double b = 1.0000001, c = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < ILP; ++q) a[q] = fma(a[q], b, c);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int q = 0; q < ILP; ++q) s += a[q];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void dadd_tput(double *out, int iters, long long *clk) {
    double a[ILP];
#pragma unroll
    for (int q = 0; q < ILP; ++q) a[q] = out[q] + q;
    const double c = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < ILP; ++q) a[q] = a[q] + c;
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int q = 0; q < ILP; ++q) s += a[q];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// DMMA m8n8k4: D(8x8) += A(8x4) B(4x8); per thread: a 1 double, b 1 double, c/d 2 doubles
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// m16n8k16 f64 (sm_90+): a 8 doubles, b 4 doubles, c 4 doubles
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
template <int ILP>
__global__ void dmma884_tput(double *out, int iters, long long *clk) {
    double d0[ILP], d1[ILP];
#pragma unroll
    for (int q = 0; q < ILP; ++q) { d0[q] = out[q]; d1[q] = out[q + 1]; }
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-3;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < ILP; ++q) dmma884(d0[q], d1[q], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int q = 0; q < ILP; ++q) s += d0[q] + d1[q];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void dmma16816_tput(double *out, int iters, long long *clk) {
    double d[ILP][4];
#pragma unroll
    for (int q = 0; q < ILP; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r) d[q][r] = out[q + r];
    double a[8], b[4];
#pragma unroll
    for (int r = 0; r < 8; ++r) a[r] = 1.0 + 1e-9 * (threadIdx.x + r);
#pragma unroll
    for (int r = 0; r < 4; ++r) b[r] = 1e-3 * (r + 1);
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < ILP; ++q) dmma16816(d[q], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int q = 0; q < ILP; ++q) s += d[q][0] + d[q][1] + d[q][2] + d[q][3];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

__global__ void shfl_chain(double *out, int iters, long long *clk) {
    double a = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < iters; ++i) a += __shfl_xor_sync(0xffffffffu, a, 1 + (i & 15));
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void shfl32_chain(float *out, int iters, long long *clk) {
    float a = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < iters; ++i) a += __shfl_xor_sync(0xffffffffu, a, 1 + (i & 15));
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void bar_chain(double *out, int iters, long long *clk) {
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void lds_chain(double *out, int iters, long long *clk) {
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 33 + 7) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) p = idx[p];
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}

// ---- cluster exchange: ping-pong and all-to-all with st.async + mbarrier --------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_pair(uint32_t raddr, double a, double b, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];"
                 :: "r"(raddr), "d"(a), "d"(b), "r"(rbar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// every CTA pushes `msgs` 16-byte messages to every CTA per round (spread over the first warps), waits for all of them
__global__ void exchange_rounds(int csize, int msgs, int rounds, long long *clk) {
    __shared__ __align__(16) double2 buf[2][1024];
    __shared__ __align__(8) uint64_t bars[2];
    uint32_t me;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(me));
    const int tid = threadIdx.x;
    const uint32_t buf_s = smem_u32(buf), bar_s = smem_u32(bars);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_s));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_s + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_s), "r"(16u * msgs * csize) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_s + 8), "r"(16u * msgs * csize) : "memory");
    }
    __syncthreads();
    cluster_sync_all();
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        const int par = r & 1;
        // thread t sends message (t / csize) to CTA (t % csize)
        if (tid < msgs * csize) {
            const int m = tid / csize, d = tid % csize;
            const uint32_t dst = buf_s + (uint32_t)((par * 1024 + me * msgs + m) * 16);
            st_async_pair(mapa_u32(dst, d), (double)r, (double)tid, mapa_u32(bar_s + par * 8, d));
        }
        while (!mbar_try_wait(bar_s + par * 8, (r >> 1) & 1)) { }
        if (tid == 0 && r + 2 < rounds)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_s + par * 8), "r"(16u * msgs * csize) : "memory");
    }
    long long t1 = clock64();
    if (tid == 0) clk[me] = t1 - t0;
    __syncthreads();
    cluster_sync_all();
}
__global__ void cluster_barrier_rounds(int rounds, long long *clk) {
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) cluster_sync_all();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

int main() {
    double *out; long long *clk;
    CK(cudaMalloc(&out, 1 << 24)); CK(cudaMalloc(&clk, 4096 * 8));
    CK(cudaMemset(out, 0, 1 << 24));
    std::vector<long long> h(4096);
    auto rd = [&](int n) { cudaDeviceSynchronize(); cudaMemcpy(h.data(), clk, n * 8, cudaMemcpyDeviceToHost); };
    const int IT = 4096;
    dfma_chain<<<1, 32>>>(out, IT, clk); rd(1); printf("DFMA dependent latency        : %.1f clk\n", (double)h[0] / IT);
    ffma_chain<<<1, 32>>>((float *)out, IT, clk); rd(1); printf("FFMA dependent latency        : %.1f clk\n", (double)h[0] / IT);
    shfl_chain<<<1, 32>>>(out, IT, clk); rd(1); printf("SHFL.64 + DADD chain          : %.1f clk\n", (double)h[0] / IT);
    shfl32_chain<<<1, 32>>>((float *)out, IT, clk); rd(1); printf("SHFL.32 + FADD chain          : %.1f clk\n", (double)h[0] / IT);
    bar_chain<<<1, 256>>>(out, IT, clk); rd(1); printf("__syncthreads (256 thr)       : %.1f clk\n", (double)h[0] / IT);
    bar_chain<<<1, 1024>>>(out, IT, clk); rd(1); printf("__syncthreads (1024 thr)      : %.1f clk\n", (double)h[0] / IT);
    lds_chain<<<1, 32>>>(out, IT, clk); rd(1); printf("LDS dependent latency         : %.1f clk\n", (double)h[0] / IT);
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        dfma_tput<8><<<1, 32 * warps>>>(out, IT, clk); rd(1);
        const double fma_per_clk = (double)warps * 32 * 8 * IT / h[0];
        printf("DFMA throughput, %2d warps x ILP8: %.2f FMA/clk/SM  (%.1f clk per warp-instr per SMSP-equivalent)\n", warps, fma_per_clk,
               32.0 * 4 / fma_per_clk);
    }
    dadd_tput<8><<<1, 512>>>(out, IT, clk); rd(1); printf("DADD throughput, 16 warps      : %.2f op/clk/SM\n", 16.0 * 32 * 8 * IT / h[0]);
    for (int warps : {1, 4, 8, 16}) {
        dmma884_tput<4><<<1, 32 * warps>>>(out, IT, clk); rd(1);
        const double fma_per_clk = (double)warps * 4 * IT * 256 / h[0];     // m8n8k4 = 256 FMA
        printf("DMMA m8n8k4, %2d warps x ILP4    : %.1f FMA/clk/SM   (%.1f clk per mma per warp)\n", warps, fma_per_clk, (double)h[0] / (4.0 * IT));
    }
    dmma884_tput<1><<<1, 32>>>(out, IT, clk); rd(1); printf("DMMA m8n8k4 dependent latency : %.1f clk\n", (double)h[0] / IT);
    for (int warps : {1, 4, 8, 16}) {
        dmma16816_tput<2><<<1, 32 * warps>>>(out, IT, clk); rd(1);
        const double fma_per_clk = (double)warps * 2 * IT * 2048 / h[0];    // m16n8k16 = 2048 FMA
        printf("DMMA m16n8k16, %2d warps x ILP2  : %.1f FMA/clk/SM   (%.1f clk per mma per warp)\n", warps, fma_per_clk, (double)h[0] / (2.0 * IT));
    }
    dmma16816_tput<1><<<1, 32>>>(out, IT, clk); rd(1); printf("DMMA m16n8k16 dependent latency: %.1f clk\n", (double)h[0] / IT);
    // whole-GPU DFMA / DMMA rate
    {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        dfma_tput<8><<<148 * 2, 512>>>(out, IT, clk); cudaDeviceSynchronize();
        cudaEventRecord(e0); dfma_tput<8><<<148 * 2, 512>>>(out, IT * 4, clk); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("whole GPU DFMA: %.2f TFLOP/s\n", 2.0 * 148 * 2 * 512 * 8 * (double)IT * 4 / ms / 1e9);
        dmma16816_tput<2><<<148 * 2, 512>>>(out, IT, clk); cudaDeviceSynchronize();
        cudaEventRecord(e0); dmma16816_tput<2><<<148 * 2, 512>>>(out, IT * 4, clk); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("whole GPU DMMA m16n8k16: %.2f TFLOP/s\n", 2.0 * 148 * 2 * 16 * 2 * 2048 * (double)IT * 4 / ms / 1e9);
        dmma884_tput<4><<<148 * 2, 512>>>(out, IT, clk); cudaDeviceSynchronize();
        cudaEventRecord(e0); dmma884_tput<4><<<148 * 2, 512>>>(out, IT * 4, clk); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("whole GPU DMMA m8n8k4: %.2f TFLOP/s\n", 2.0 * 148 * 2 * 16 * 4 * 256 * (double)IT * 4 / ms / 1e9);
    }
    // cluster exchanges
    CK(cudaFuncSetAttribute(exchange_rounds, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CK(cudaFuncSetAttribute(cluster_barrier_rounds, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    for (int cs : {2, 4, 8, 16}) {
        for (int msgs : {1, 4, 16, 32}) {
            if (msgs * cs > 512) continue;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(cs); cfg.blockDim = dim3(512);
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            const int rounds = 2000;
            CK(cudaLaunchKernelEx(&cfg, exchange_rounds, cs, msgs, rounds, clk));
            rd(cs);
            printf("cluster %2d all-to-all, %2d x16B msgs per CTA pair: %.0f clk per round\n", cs, msgs, (double)h[0] / rounds);
        }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(cs); cfg.blockDim = dim3(256);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, cluster_barrier_rounds, 2000, clk));
        rd(1);
        printf("cluster %2d barrier.cluster arrive+wait: %.0f clk\n", cs, (double)h[0] / 2000);
    }
    return 0;
}
