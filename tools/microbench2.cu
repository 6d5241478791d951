// Does the evolution inner loop (R=3 rows x D=21, z from shared memory) reach the FP64 pipe rate?
#include <cstdio>
#include <cuda_runtime.h>
constexpr int D = 21, R = 3;
template <int WARPS>
__global__ void __maxnreg__(255) evo(double* out, long long* cyc, int steps) {
    __shared__ double2 zb[WARPS][4 * D];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane / 8;
    double v[R][D];
#pragma unroll
    for (int k = 0; k < R; ++k)
#pragma unroll
        for (int i = 0; i < D; ++i) v[k][i] = out[(threadIdx.x + k * 7 + i) & 1023];
    for (int i = lane; i < 4 * D; i += 32) zb[warp][i] = make_double2(0.001 * i, 0.002 * i);
    __syncthreads();
    double tot = 0.0;
    long long t0 = clock64();
    for (int s = 0; s < steps; ++s) {
        const double2* z0 = &zb[warp][g * D];
        double pr[R][2], pi[R][2];
#pragma unroll
        for (int kk = 0; kk < R; ++kk) { pr[kk][0] = pr[kk][1] = 0.0; pi[kk][0] = pi[kk][1] = 0.0; }
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const double2 z = z0[i];
#pragma unroll
            for (int kk = 0; kk < R; ++kk) {
                pr[kk][i & 1] = fma(v[kk][i], z.x, pr[kk][i & 1]);
                pi[kk][i & 1] = fma(v[kk][i], z.y, pi[kk][i & 1]);
            }
        }
#pragma unroll
        for (int kk = 0; kk < R; ++kk) {
            const double re = pr[kk][0] + pr[kk][1], im = pi[kk][0] + pi[kk][1];
            tot = fma(re, re, fma(im, im, tot));
        }
        if (tot == 123.0) zb[warp][lane] = make_double2(tot, tot);   // keep the loads inside the loop
        __syncwarp();
    }
    long long t1 = clock64();
    out[threadIdx.x] = tot;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int WARPS>
void run(double* out, long long* cyc) {
    const int steps = 2000;
    evo<WARPS><<<1, WARPS * 32>>>(out, cyc, steps);
    cudaDeviceSynchronize();
    evo<WARPS><<<1, WARPS * 32>>>(out, cyc, steps);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_step = (double)h / steps;
    printf("%2d warps/SM: %.1f cycles per step per warp ; FP64 pipe demand %d cycles/step/warp -> pipe utilisation %.0f%%\n", WARPS,
           per_step, 2 * (126 + 9), 100.0 * (WARPS / 4.0 > 1 ? WARPS / 4.0 : 1.0) * 2 * 135 / per_step / (WARPS < 4 ? 1 : 1));
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1 << 16); cudaMemset(out, 0, 1 << 16); cudaMalloc(&cyc, 64);
    run<1>(out, cyc); run<4>(out, cyc); run<8>(out, cyc); run<12>(out, cyc); run<16>(out, cyc);
    return 0;
}
