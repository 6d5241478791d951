// What bounds the register-row rotation replay (qtet_bigreg.cu) and is DMMA worth using?  B200, sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench3 tools/microbench3.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int D = 66, DP = 68;

// MODE 0: LDS.128 (c, s) per plane + 4 FP64 (one chain)     MODE 1: two fused rounds (two chains)
// MODE 2: (c, s) from registers (no LDS)                       MODE 3: LDS only (results summed, 1 DADD per load)
template <int MODE>
__global__ void __launch_bounds__(96, 4) replay(double* out, int rounds) {
    __shared__ double2 cs[2][DP];
    for (int i = threadIdx.x; i < 2 * DP; i += blockDim.x) (&cs[0][0])[i] = make_double2(0.8, 0.6);
    __syncthreads();
    double v[D];
#pragma unroll
    for (int j = 0; j < D; ++j) v[j] = (j == threadIdx.x) ? 1.0 : 0.001 * j;
    double acc = 0.0;
    const double2 creg = make_double2(out[0] + 0.8, out[1] + 0.6);
    for (int r = 0; r < rounds; ++r) {
        const double2* c0 = cs[r & 1];
        const double2* c1 = cs[(r + 1) & 1];
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = D - 2; i >= 0; --i) {
                const double2 c = (MODE == 2) ? creg : c0[i];
                const double x = v[i], y = v[i + 1];
                const double t1 = c.x * x, t2 = c.y * x;
                v[i] = fma(-c.y, y, t1);
                v[i + 1] = fma(c.x, y, t2);
            }
        } else if (MODE == 1) {
#pragma unroll
            for (int i = D - 2; i >= -2; --i) {
                if (i >= 0) {
                    const double2 c = c0[i];
                    const double x = v[i], y = v[i + 1];
                    const double t1 = c.x * x, t2 = c.y * x;
                    v[i] = fma(-c.y, y, t1);
                    v[i + 1] = fma(c.x, y, t2);
                }
                if (i + 2 <= D - 2) {
                    const double2 c = c1[i + 2];
                    const double x = v[i + 2], y = v[i + 3];
                    const double t1 = c.x * x, t2 = c.y * x;
                    v[i + 2] = fma(-c.y, y, t1);
                    v[i + 3] = fma(c.x, y, t2);
                }
            }
            ++r;
        } else {
#pragma unroll
            for (int i = D - 2; i >= 0; --i) { const double2 c = c0[i]; acc += c.x; }
        }
        if (v[0] == 123.456) cs[0][threadIdx.x & 63] = make_double2(acc, v[1]);   // keep the loads in the loop
    }
    double s = acc;
#pragma unroll
    for (int j = 0; j < D; ++j) s += v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x + 2] = s;
}

// evolution-like: per i one LDS.128 z feeding 2 FMA (MODE 0) ; two time samples -> 2 LDS.128 + 4 FMA (MODE 1)
template <int MODE>
__global__ void __launch_bounds__(96, 4) evo(double* out, int steps) {
    __shared__ double2 zt[4][DP];
    for (int i = threadIdx.x; i < 4 * DP; i += blockDim.x) (&zt[0][0])[i] = make_double2(0.001 * i, 0.002);
    __syncthreads();
    double v[D];
#pragma unroll
    for (int j = 0; j < D; ++j) v[j] = 0.001 * j + threadIdx.x;
    double tot = 0.0;
    for (int s = 0; s < steps; ++s) {
        const double2* z0 = zt[s & 3];
        const double2* z1 = zt[(s + 1) & 3];
        double pr0 = 0, pi0 = 0, pr1 = 0, pi1 = 0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const double2 za = z0[i];
            pr0 = fma(v[i], za.x, pr0); pi0 = fma(v[i], za.y, pi0);
            if (MODE == 1) { const double2 zb = z1[i]; pr1 = fma(v[i], zb.x, pr1); pi1 = fma(v[i], zb.y, pi1); }
        }
        tot += pr0 * pr0 + pi0 * pi0 + pr1 * pr1 + pi1 * pi1;
        if (tot == 123.0) zt[0][threadIdx.x & 63] = make_double2(tot, tot);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x + 2] = tot;
}

// DMMA m8n8k4: NACC independent accumulators per warp
template <int NACC>
__global__ void dmma(double* out, int iters) {
    double a = out[threadIdx.x & 31] + 1.0, b = out[(threadIdx.x + 7) & 31] + 0.5;
    double c[NACC][2];
#pragma unroll
    for (int k = 0; k < NACC; ++k) { c[k][0] = k; c[k][1] = -k; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < NACC; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < NACC; ++k) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x + 2] = s;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    double* out; cudaMalloc(&out, 1 << 24); cudaMemset(out, 0, 1 << 24);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    const int rounds = 4000;
    const char* names[4] = {"LDS.128 + 4 FP64, 1 chain", "two fused rounds (2 chains)", "(c,s) in registers", "LDS.128 only + DADD"};
    for (int mode = 0; mode < 4; ++mode) {
        float ms = 0;
        if (mode == 0) ms = timeit([&] { replay<0><<<sms * 4, 96>>>(out, rounds); });
        if (mode == 1) ms = timeit([&] { replay<1><<<sms * 4, 96>>>(out, rounds); });
        if (mode == 2) ms = timeit([&] { replay<2><<<sms * 4, 96>>>(out, rounds); });
        if (mode == 3) ms = timeit([&] { replay<3><<<sms * 4, 96>>>(out, rounds); });
        const double planes = (double)rounds * (D - 1) * 12;       // warp-plane-steps per SM
        const double cyc = ms * 1e-3 * ghz * 1e9;
        printf("replay %-30s: %.2f SM-cycles per warp-plane  (FP64 pipe %.0f%% at 4 FP64/plane)\n", names[mode], cyc / planes, 100.0 * 2.0 / (cyc / planes));
    }
    for (int mode = 0; mode < 2; ++mode) {
        const int steps = 20000;
        float ms = mode ? timeit([&] { evo<1><<<sms * 4, 96>>>(out, steps); }) : timeit([&] { evo<0><<<sms * 4, 96>>>(out, steps); });
        const double fma_w = (double)steps * D * (mode ? 4 : 2) * 12;
        const double cyc = ms * 1e-3 * ghz * 1e9;
        printf("evolution %d sample(s)/pass: %.2f SM-cycles per warp-FMA (FP64 pipe %.0f%%)\n", mode + 1, cyc / fma_w, 100.0 * 0.5 / (cyc / fma_w));
    }
    {
        const int iters = 20000;
        float ms = timeit([&] { dmma<8><<<sms * 4, 128>>>(out, iters); });
        const double flops = 2.0 * 256 * 8 * (double)iters * 4 * 4 * sms;
        printf("DMMA m8n8k4, 16 warps/SM x 8 accumulators: %.2f TFLOP/s\n", flops / (ms * 1e-3) / 1e12);
        ms = timeit([&] { dmma<8><<<sms * 8, 256>>>(out, iters); });
        printf("DMMA m8n8k4, 64 warps/SM x 8 accumulators: %.2f TFLOP/s\n", 2.0 * 256 * 8 * (double)iters * 8 * 8 * sms / (ms * 1e-3) / 1e12);
    }
    return 0;
}
