// Latency / throughput micro-benchmarks for the FP64 and shared-memory paths on B200 (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench tools/microbench.cu && /tmp/microbench
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat_dfma(double* out, long long* cyc, int n, double a, double b) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) x = fma(x, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lat_dmul_dadd(double* out, long long* cyc, int n, double a, double b) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { x = x * a; x = x + b; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void thr_dfma(double* out, long long* cyc, int n, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = out[threadIdx.x] + k;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lat_lds(double* out, long long* cyc, int n) {
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 7 + 1) & 1023;
    __syncthreads();
    int j = threadIdx.x & 31;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) j = idx[j];
    }
    long long t1 = clock64();
    out[threadIdx.x] = j;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// broadcast LDS.128 throughput: all lanes read the same 16 B, many independent loads
template <int W>
__global__ void thr_lds_bcast(double* out, long long* cyc, int n) {
    __shared__ double2 buf[64];
    if (threadIdx.x < 64) buf[threadIdx.x] = make_double2(threadIdx.x, 1.0);
    __syncthreads();
    double acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            double vx, vy;
            const unsigned addr = (unsigned)__cvta_generic_to_shared(&buf[(u + i) & 63]);
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(vx), "=d"(vy) : "r"(addr));
            acc += vx + vy;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void thr_lds64_bcast(double* out, long long* cyc, int n) {
    __shared__ double buf[64];
    if (threadIdx.x < 64) buf[threadIdx.x] = threadIdx.x;
    __syncthreads();
    double acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) acc += *(volatile double*)&buf[(u + i) & 63];
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void thr_shfl(double* out, long long* cyc, int n) {
    double x = out[threadIdx.x];
    double acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) acc += __shfl_sync(0xffffffffu, x, (u + i) & 31);
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lat_rsqrt(double* out, long long* cyc, int n) {
    double x = out[threadIdx.x] + 1.5;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) x = rsqrt(x) + 1.25;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lat_div(double* out, long long* cyc, int n) {
    double x = out[threadIdx.x] + 1.5;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) x = 3.0 / x + 1.25;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
double run(F launch, int blocks, double ops_per_thread_iter_total, long long* dcyc) {
    launch();
    cudaDeviceSynchronize();
    launch();
    cudaDeviceSynchronize();
    long long h[1024];
    cudaMemcpy(h, dcyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < blocks; ++i) mx = h[i] > mx ? h[i] : mx;
    return (double)mx / ops_per_thread_iter_total;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMemset(out, 0, 1 << 20);
    cudaMalloc(&cyc, 8192);
    const int n = 2000;
    printf("dependent DFMA latency            : %.2f cycles\n", run([&] { lat_dfma<<<1, 32>>>(out, cyc, n, 0.999, 1e-3); }, 1, n * 16.0, cyc));
    printf("dependent DMUL+DADD pair latency   : %.2f cycles (per op)\n", run([&] { lat_dmul_dadd<<<1, 32>>>(out, cyc, n, 0.999, 1e-3); }, 1, n * 16.0, cyc));
    printf("DFMA issue, 1 warp, ILP 8          : %.2f cycles/inst\n", run([&] { thr_dfma<8><<<1, 32>>>(out, cyc, n, 0.999, 1e-3); }, 1, n * 32.0, cyc));
    printf("DFMA issue, 4 warps (1/SMSP) ILP 8 : %.2f cycles/inst/warp\n", run([&] { thr_dfma<8><<<1, 128>>>(out, cyc, n, 0.999, 1e-3); }, 1, n * 32.0, cyc));
    printf("DFMA issue, 16 warps ILP 8         : %.2f cycles/inst/warp (x16 warps)\n", run([&] { thr_dfma<8><<<1, 512>>>(out, cyc, n, 0.999, 1e-3); }, 1, n * 32.0, cyc));
    printf("dependent LDS.32 latency           : %.2f cycles\n", run([&] { lat_lds<<<1, 32>>>(out, cyc, n); }, 1, n * 16.0, cyc));
    printf("bcast LDS.128, 1 warp              : %.2f cycles/load\n", run([&] { thr_lds_bcast<1><<<1, 32>>>(out, cyc, n); }, 1, n * 16.0, cyc));
    printf("bcast LDS.128, 4 warps             : %.2f cycles/load/warp\n", run([&] { thr_lds_bcast<4><<<1, 128>>>(out, cyc, n); }, 1, n * 16.0, cyc));
    printf("bcast LDS.128, 16 warps            : %.2f cycles/load/warp -> %.2f cycles per warp-load per SM\n", run([&] { thr_lds_bcast<16><<<1, 512>>>(out, cyc, n); }, 1, n * 16.0, cyc), run([&] { thr_lds_bcast<16><<<1, 512>>>(out, cyc, n); }, 1, n * 16.0 * 16, cyc));
    printf("bcast LDS.64, 16 warps             : %.2f cycles per warp-load per SM\n", run([&] { thr_lds64_bcast<<<1, 512>>>(out, cyc, n); }, 1, n * 16.0 * 16, cyc));
    printf("SHFL.64 (2x32), 16 warps           : %.2f cycles per warp-shfl64 per SM\n", run([&] { thr_shfl<<<1, 512>>>(out, cyc, n); }, 1, n * 16.0 * 16, cyc));
    printf("dependent rsqrt(double)+add        : %.2f cycles\n", run([&] { lat_rsqrt<<<1, 32>>>(out, cyc, n); }, 1, n * 4.0, cyc));
    printf("dependent div(double)+add          : %.2f cycles\n", run([&] { lat_div<<<1, 32>>>(out, cyc, n); }, 1, n * 4.0, cyc));
    return 0;
}
