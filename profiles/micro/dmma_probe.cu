// Micro-benchmark: FP64 DFMA vs DMMA (mma.sync m8n8k4 / m16n8k8 f64) dependent-issue latency and throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double *c, const double *a, const double *b)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int CHAINS>
__global__ void k_dmma884(int iters, double *out, long long *cyc)
{
    double c[CHAINS][2];
    for (int i = 0; i < CHAINS; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < CHAINS; i++) dmma884(c[i][0], c[i][1], a, b);
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CHAINS; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CHAINS>
__global__ void k_dmma1688(int iters, double *out, long long *cyc)
{
    double c[CHAINS][4];
    for (int i = 0; i < CHAINS; i++) for (int j = 0; j < 4; j++) c[i][j] = threadIdx.x + j;
    double a[4] = {1.0 + 1e-9 * threadIdx.x, 1.0, 0.5, 0.25}, b[2] = {1.0 - 1e-9 * threadIdx.x, 0.5};
    long long t0 = clock64();
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < CHAINS; i++) dmma1688(c[i], a, b);
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CHAINS; i++) for (int j = 0; j < 4; j++) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CHAINS>
__global__ void k_dfma(int iters, double *out, long long *cyc)
{
    double c[CHAINS];
    for (int i = 0; i < CHAINS; i++) c[i] = threadIdx.x + i;
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < CHAINS; i++) c[i] = fma(c[i], a, b);
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CHAINS; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_lds(int iters, double *out, long long *cyc)
{
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 7 + 1) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) p = idx[p];
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_shfl(int iters, double *out, long long *cyc)
{
    double v = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) v = __shfl_xor_sync(0xffffffffu, v, 1) + 1.0;
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_rsqrt(int iters, double *out, long long *cyc)
{
    double v = 2.0 + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) v = rsqrt(v) + 1.5;
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <class F>
void run(const char *name, F launch, int blocks, int threads, int iters, double ops_per_thread_iter, double flop_per_op)
{
    double *out; long long *cyc;
    cudaMalloc(&out, sizeof(double) * blocks * threads); cudaMalloc(&cyc, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(blocks, threads, iters, out, cyc); cudaDeviceSynchronize();
    cudaEventRecord(e0); launch(blocks, threads, iters, out, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double warps = (double)blocks * threads / 32.0;
    printf("%-28s blocks %5d thr %4d  cycles/iter(warp0) %8.2f  ms %8.3f  TFLOP/s %7.2f  err %s\n", name, blocks, threads, (double)h / iters, ms,
           warps * iters * ops_per_thread_iter * flop_per_op / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    const int it = 20000;
#define L(K) [](int b, int t, int i, double *o, long long *c) { K<<<b, t>>>(i, o, c); }
    // latency: 1 warp, 1 chain
    run("dfma chain1 1warp", L(k_dfma<1>), 1, 32, it, 1, 64);
    run("dfma chain8 1warp", L(k_dfma<8>), 1, 32, it, 8, 64);
    run("dmma884 chain1 1warp", L(k_dmma884<1>), 1, 32, it, 1, 512);
    run("dmma884 chain2 1warp", L(k_dmma884<2>), 1, 32, it, 2, 512);
    run("dmma884 chain4 1warp", L(k_dmma884<4>), 1, 32, it, 4, 512);
    run("dmma884 chain8 1warp", L(k_dmma884<8>), 1, 32, it, 8, 512);
    run("dmma1688 chain1 1warp", L(k_dmma1688<1>), 1, 32, it, 1, 2048);
    run("dmma1688 chain4 1warp", L(k_dmma1688<4>), 1, 32, it, 4, 2048);
    run("lds dependent", L(k_lds), 1, 32, it, 1, 0);
    run("shfl+dadd dependent", L(k_shfl), 1, 32, it, 1, 0);
    run("rsqrt+dadd dependent", L(k_rsqrt), 1, 32, it, 1, 0);
    // throughput: full chip
    run("dfma chain8 full", L(k_dfma<8>), 148 * 8, 256, it, 8, 64);
    run("dmma884 chain8 full", L(k_dmma884<8>), 148 * 8, 256, it, 8, 512);
    run("dmma1688 chain4 full", L(k_dmma1688<4>), 148 * 8, 256, it, 4, 2048);
    run("dmma884 chain4 4w/SM", L(k_dmma884<4>), 148, 128, it, 4, 512);
    run("dmma884 chain4 8w/SM", L(k_dmma884<4>), 148, 256, it, 4, 512);
    run("dmma884 chain4 16w/SM", L(k_dmma884<4>), 148, 512, it, 4, 512);
    return 0;
}
