// Shared declarations of libqtet_b200 (host handle, device views, small device helpers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/qtet.h"

namespace qtet {

// Device-side, chi-independent description of one problem (all pointers on the device).
struct ProblemView {
    int N, f, d, T;
    int tridiag;          // 1: H is tridiagonal in the reference's basis order (dimer)
    int offd_nonzero;     // tridiag only: every off-diagonal element is non-zero (the direct recurrences divide by them)
    double max_t;
    const double* lin;    // [d]     sum_k omega_k n_k(m)                 HamiltonianLoss.py:132-133
    const double* quad;   // [d, f]  n_k(m)^2 / 2 = dH_mm/dchi_k          HamiltonianLoss.py:134
    const double* occ;    // [d, f]  n_k(m)                               HamiltonianLoss.py:62-110
    const double* offd;   // tridiag: [d-1] H[m+1,m]; dense: [d, d] hopping part   :136-163
    const double* tgrid;  // [T]     linspace(0, max_t, T)                HamiltonianLoss.py:221
    const double* omegas; // [f]
    unsigned long long* stats;   // [4] cumulative counters: 0 direct-path points sent to the generic kernel, 1 rotation-stream
                                 //     overflows (split paths), 2 eigen-solves that hit the iteration limit, 3 reserved
};

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);

#define QTET_CUDA(call)                                             \
    do {                                                            \
        cudaError_t _e = (call);                                    \
        if (_e != cudaSuccess) return ::qtet::cuda_fail(_e, #call); \
    } while (0)

struct Profiler;

// Entry points switch to the handle's device and give the caller's current device back on return (a process may hold
// handles on several devices, or have torch's current device elsewhere).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) err = cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};
#define QTET_ON_DEVICE(dev)                                   \
    DeviceGuard _guard(dev);                                  \
    if (_guard.err != cudaSuccess) return cuda_fail(_guard.err, "cudaSetDevice")

// ---- launchers implemented in the kernel translation units -------------------------------
int launch_generic(const ProblemView& pv, const double* chis, int64_t B, int site, double* loss,
                   double* grad, int32_t* tstar, double* trace, cudaStream_t st, Profiler* prof = nullptr);
size_t generic_smem_bytes(int d, int nthreads);
int generic_threads(int d);
int launch_build_h(const ProblemView& pv, const double* chis, double* H, cudaStream_t st);

// Optional per-kernel timing with CUDA events on the launching streams (bench.py's roofline figures).
struct Profiler {
    static constexpr int kKinds = 3;          // 0: ql_scalar (K_A), 1: apply (K_B), 2: fused generic
    static constexpr int kMaxPairs = 4096;
    bool on = false;
    std::vector<cudaEvent_t> beg[kKinds], end[kKinds];
    std::vector<long long> points[kKinds];
    void begin(int kind, long long npoints, cudaStream_t st) {
        if (!on || (int)beg[kind].size() >= kMaxPairs) return;
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        beg[kind].push_back(a); end[kind].push_back(b); points[kind].push_back(npoints);
        cudaEventRecord(a, st);
    }
    void finish(int kind, cudaStream_t st) {
        if (!on || beg[kind].empty() || (int)beg[kind].size() > kMaxPairs) return;
        cudaEventRecord(end[kind].back(), st);
    }
    void clear() {
        for (int k = 0; k < kKinds; ++k) {
            for (auto e : beg[k]) cudaEventDestroy(e);
            for (auto e : end[k]) cudaEventDestroy(e);
            beg[k].clear(); end[k].clear(); points[k].clear();
        }
    }
};

// Optional per-chunk hooks of the split path (the host-buffer entry point hangs its H2D / D2H copies on them):
// `before` runs when chunk [lo, lo+n) is about to be queued (its inputs are first read on `aux`), `after` once everything
// that writes the chunk's outputs has been queued on `st`.
struct ChunkHook {
    void* ctx = nullptr;
    int (*before)(void* ctx, long long lo, long long n, cudaStream_t aux) = nullptr;
    int (*after)(void* ctx, long long lo, long long n, cudaStream_t st) = nullptr;
};

struct SplitWorkspace;   // rotation streams etc. (qtet_split.cu)
bool split_eligible(const ProblemView& pv);
int launch_split(const ProblemView& pv, SplitWorkspace** ws, const double* chis, int64_t B, int site,
                 double* loss, double* grad, int32_t* tstar, double* trace, cudaStream_t st, Profiler* prof = nullptr,
                 const ChunkHook* hook = nullptr);
void split_workspace_free(SplitWorkspace* ws);

struct DirectWorkspace;  // eigenvalues by QL + eigenvectors by two-sided recurrences, tridiagonal d <= 32 (qtet_direct.cu)
bool direct_eligible(const ProblemView& pv);
// n_dev (optional, device pointer): process min(B, *n_dev) points -- the optimiser's compacted live list, no host round trip
int launch_direct(const ProblemView& pv, DirectWorkspace** ws, const double* chis, int64_t B, const int* n_dev, int site,
                  double* loss, double* grad, int32_t* tstar, double* trace, cudaStream_t st, Profiler* prof = nullptr,
                  const ChunkHook* hook = nullptr);
int direct_reserve(const ProblemView& pv, DirectWorkspace** ws, int64_t B, bool single_chunk = false);
void direct_workspace_free(DirectWorkspace* ws);

struct BigWorkspace;     // large-d split path (qtet_big.cu)
bool big_eligible(const ProblemView& pv);
int launch_big(const ProblemView& pv, BigWorkspace** ws, const double* chis, int64_t B, int site, double* loss,
               double* grad, int32_t* tstar, double* trace, cudaStream_t st, Profiler* prof = nullptr);
void big_workspace_free(BigWorkspace* ws);
// large-d direct path (eigenvectors by recurrences, no rotation stream); shares the workspace type
bool big_direct_eligible(const ProblemView& pv);
int launch_big_direct(const ProblemView& pv, BigWorkspace** ws, const double* chis, int64_t B, int site, double* loss,
                      double* grad, int32_t* tstar, double* trace, cudaStream_t st, Profiler* prof = nullptr,
                      const ChunkHook* hook = nullptr);

int launch_argmin(const double* v, int64_t B, double* min_out, int64_t* idx_out, cudaStream_t st);
// locality order of a batch (qtet_adam.cu): coarse 2-D binning of the points, results scattered back to the caller's order
int launch_locality_order(const double* x, int64_t B, int f, int* order, int* keys, int* hist, unsigned long long* mm, cudaStream_t st);
int launch_sort_gather(const int* order, int64_t B, int f, const double* x, double* xs, cudaStream_t st);
int launch_sort_scatter(const int* order, int64_t B, int f, const double* loss_s, const double* grad_s, const int* ts_s,
                        double* loss, double* grad, int* ts, cudaStream_t st);

struct AdamState;        // per-guess optimiser state on the device (qtet_adam.cu)

}  // namespace qtet

struct qtet_problem {
    int N, f, d, T, device;
    double coupling, max_t;
    bool tridiag;
    int path;                               // QTET_PATH_*
    std::vector<double> omegas, states, tgrid;
    double* dev_blob = nullptr;             // one allocation holding every constant array
    qtet::ProblemView view{};
    qtet::SplitWorkspace* split_ws = nullptr;
    qtet::DirectWorkspace* direct_ws = nullptr;
    qtet::BigWorkspace* big_ws = nullptr;
    qtet::Profiler prof;
    // scratch for *_host convenience calls and the optimiser
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // qtet_set_locality_sort: evaluate batches in locality order (own buffers: the optimiser's scratch stays untouched)
    int locality_sort = 0;
    void* sort_ws = nullptr;
    size_t sort_bytes = 0;
    // copy/compute pipeline of qtet_loss_grad_host: H2D, kernels and D2H of consecutive slices overlap
    static constexpr int kMaxSlices = 64;
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[kMaxSlices] = {}, ev_run[kMaxSlices] = {};
    // qtet_adam_run: pinned words the live counts are copied to (one per epoch, ring), their events, the copy stream
    static constexpr int kPollRing = 16;
    int* h_poll = nullptr;
    cudaStream_t s_poll = nullptr;
    cudaEvent_t ev_cnt[kPollRing] = {}, ev_poll[kPollRing] = {};
};

namespace qtet {
int ensure_scratch(qtet_problem* p, size_t bytes);
int loss_grad_dispatch(qtet_problem* p, const double* chis, int64_t B, int site, double* loss,
                       double* grad, int32_t* tstar, double* trace, cudaStream_t st, const ChunkHook* hook = nullptr,
                       const int* n_dev = nullptr);
}

// ---- device helpers ---------------------------------------------------------------------
#ifdef __CUDACC__
namespace qtet {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// (exp(-i x) - 1) / x, accurate for |x| <= 1e-2 (series to x^7)
__device__ __forceinline__ void expm1_over_x(double x, double& re, double& im) {
    const double x2 = x * x;
    re = x * (-0.5 + x2 * (1.0 / 24.0 + x2 * (-1.0 / 720.0 + x2 * (1.0 / 40320.0))));
    im = -1.0 + x2 * (1.0 / 6.0 + x2 * (-1.0 / 120.0 + x2 * (1.0 / 5040.0)));
}

// Daleckii-Krein divided difference of exp(-i e t) between eigenvalues (ei, fi) and (ej, fj):
// (f_i - f_j)/(e_i - e_j), degeneracy-safe; fi/fj are the unit phases exp(-i e t).
__device__ __forceinline__ void dk_phi(double ei, double fir, double fii, double ej, double fjr,
                                       double fji, double t, double& pr, double& pi) {
    const double dl = ei - ej;
    const double x = dl * t;
    if (fabs(x) < 1e-2) {
        double gr, gi;
        expm1_over_x(x, gr, gi);
        pr = t * (fjr * gr - fji * gi);
        pi = t * (fjr * gi + fji * gr);
    } else {
        const double inv = 1.0 / dl;
        pr = (fir - fjr) * inv;
        pi = (fii - fji) * inv;
    }
}

}  // namespace qtet
#endif
