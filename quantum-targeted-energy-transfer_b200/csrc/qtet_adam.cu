// Batched per-guess optimiser (Optimizer.train, Optimizer.py:99-242), arg-min over guesses
// (solver_mp.py:182-184) and the FP64 peak micro-benchmark.
//
// One thread per initial guess carries the whole per-guess state of the reference loop: Keras Adam
// moments, best-so-far tracking, the step-revert rule, the lr drop at loss<=0.1, the TET break and the
// ">2 sub-tol moves" stop.  Guesses that stopped are compacted away so that the loss+grad kernels only
// ever see live guesses (ragged convergence).
#include <cmath>
#include <cstdlib>
#include <limits>

#include "qtet_common.cuh"

namespace qtet {

namespace {

struct AdamArgs {
    int f;
    long long n_active;       // upper bound known to the host ...
    const int* n_dev;         // ... and, when non-null, the exact count on the device
    int* cnt_next;            // counter the NEXT epoch's compaction accumulates into (zeroed here)
    const int* active;        // [n_active] guess ids
    const double* xc;         // [n_active, f] compacted variables the loss/grad was evaluated at
    const double* loss;       // [n_active]
    const double* grad;       // [n_active, f]
    double* x;                // [B, f]
    double* m; double* v; double* vhat;   // [B, f]
    double* lr;               // [B]
    double* last_loss;        // [B]  mylosses[epoch]
    double* min_loss;         // [B]  min(mylosses[:epoch+1])
    double* best_vars;        // [B, f]
    int* err_count;           // [B, f]
    int* alive;               // [B]
    int* epochs;              // [B]
    unsigned train_mask;
    double beta_1, beta_2, epsilon, tol, s2, s1;
    int amsgrad;
    int epoch;
    long long B;
    double* loss_hist;        // [iterations, B] or null
    double* var_hist;         // [iterations/10, B, f] or null
};

// Optimizer.py:146-224, one epoch for every live guess.
__global__ void adam_step_kernel(AdamArgs a) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i == 0 && a.cnt_next) *a.cnt_next = 0;
    const long long n = a.n_dev ? (long long)*a.n_dev : a.n_active;
    if (i >= n) return;
    const int b = a.active[i];
    const int f = a.f;
    const double L = a.loss[i];
    double lr = a.lr[b];
    if (L <= 0.1) lr = 0.0001;                               // :165-166
    a.lr[b] = lr;
    if (a.loss_hist) a.loss_hist[(long long)a.epoch * a.B + b] = L;      // :169
    const double last = a.last_loss[b];
    double mn = a.min_loss[b];
    if (L < mn) {                                            // :172-175 (before the step)
        for (int k = 0; k < f; ++k) a.best_vars[(long long)b * f + k] = a.xc[i * f + k];
        mn = L;
    }
    a.min_loss[b] = mn;
    a.last_loss[b] = L;
    // Explicitly rounded operations (no FMA contraction): the update is bit-identical to the same
    // expression evaluated in IEEE double on the host, which keeps the (chaotic) trajectories comparable.
    const double alpha = __ddiv_rn(__dmul_rn(lr, a.s2), a.s1);          // Keras Adam, t = epoch+1
    bool stop = false;
    const bool revert = (L - last >= 0.5) && (L >= 0.5);     // :184
    bool counted_stop = false;
    const double omb1 = __dsub_rn(1.0, a.beta_1), omb2 = __dsub_rn(1.0, a.beta_2);
    for (int k = 0; k < f; ++k) {
        const long long o = (long long)b * f + k;
        const double x0 = a.xc[i * f + k];
        double x1 = x0;
        if ((a.train_mask >> k) & 1u) {
            const double g = a.grad[i * f + k];
            double m = a.m[o], v = a.v[o];
            m = __dadd_rn(m, __dmul_rn(__dsub_rn(g, m), omb1));
            v = __dadd_rn(v, __dmul_rn(__dsub_rn(__dmul_rn(g, g), v), omb2));
            a.m[o] = m; a.v[o] = v;
            double denom_v = v;
            if (a.amsgrad) { double vh = fmax(a.vhat[o], v); a.vhat[o] = vh; denom_v = vh; }
            x1 = __dsub_rn(x0, __ddiv_rn(__dmul_rn(alpha, m), __dadd_rn(__dsqrt_rn(denom_v), a.epsilon)));
            const double err = fabs(__dsub_rn(x1, x0));      // :181
            if (!counted_stop && err < a.tol) {              // :202-206 (first offending site returns)
                const int c = a.err_count[o] + 1;
                a.err_count[o] = c;
                if (c > 2) counted_stop = true;
            }
        }
        a.x[o] = revert ? x0 : x1;                           // :184-186
    }
    if (a.var_hist && ((a.epoch + 1) % 10 == 0)) {           // :189-192
        const long long slot = (a.epoch + 1) / 10 - 1;
        for (int k = 0; k < f; ++k) a.var_hist[(slot * a.B + b) * f + k] = a.x[(long long)b * f + k];
    }
    if (fabs(L) < 0.1) stop = true;                          // :195-196
    else if (counted_stop) stop = true;
    a.epochs[b] = a.epoch + 1;
    if (stop) a.alive[b] = 0;
}

// Build the compacted list of live guesses and gather their variables.
__global__ void compact_kernel(const int* alive, long long B, int* active, int* n_active) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const bool live = (b < B) && alive[b];
    const unsigned ballot = __ballot_sync(0xffffffffu, live);
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && ballot) base = atomicAdd(n_active, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (live) active[base + __popc(ballot & ((1u << lane) - 1u))] = (int)b;
}

__global__ void gather_kernel(const int* active, long long n, int f, const double* x, double* xc) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n * f) return;
    const long long r = i / f;
    const int k = (int)(i - r * f);
    xc[i] = x[(long long)active[r] * f + k];
}

// Compaction and gather in one launch over the PREVIOUS live list (a guess can only leave it): new live list,
// its variables, and its length -- all on the device, the host never needs the count to size a launch.
//   prev == nullptr: first epoch, every guess 0 .. n_prev-1 is a candidate.
__global__ void compact_gather_kernel(const int* alive, const int* prev, long long n_prev_host, const int* n_prev_dev,
                                      int f, const double* x, int* active, double* xc, int* n_active) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long n_prev = n_prev_dev ? (long long)*n_prev_dev : n_prev_host;
    int b = -1;
    if (i < n_prev) b = prev ? prev[i] : (int)i;
    const bool live = (b >= 0) && alive[b];
    const unsigned ballot = __ballot_sync(0xffffffffu, live);
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && ballot) base = atomicAdd(n_active, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (live) {
        const int slot = base + __popc(ballot & ((1u << lane) - 1u));
        active[slot] = b;
        for (int k = 0; k < f; ++k) xc[(long long)slot * f + k] = x[(long long)b * f + k];
    }
}

// ---------------------------------------------------------------------------------- locality order of the live list
// The loss+grad kernels put 32 (K_E) / 4 (K_D) CONSECUTIVE points of the list into one warp, and a warp costs what its most
// expensive point costs (QL sweeps, columns that carry weight).  Random guesses evaluate 27 % slower than the same number of
// grid-ordered points (1.07e8 vs 1.37e8 evals/s, tools/solver_point_cost.py); ordering the initial list by a coarse 256 x 256
// binning of (donor chi, acceptor chi) puts neighbours of the landscape side by side.  Results are batch-invariant per point,
// so the order (including the arbitrary order inside a bin) cannot change a single bit of the output.
__device__ __forceinline__ unsigned long long ordered_bits(double x) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}
__device__ __forceinline__ double from_ordered_bits(unsigned long long u) {
    return __longlong_as_double((long long)((u >> 63) ? (u & 0x7fffffffffffffffULL) : ~u));
}
__device__ __forceinline__ long long bin_count(long long B, const int* n_dev) {
    if (n_dev == nullptr) return B;
    const long long n = *n_dev;
    return n < B ? n : B;
}
__global__ void bin_minmax_kernel(const double* x, long long B, const int* n_dev, int f, int k0, int k1, unsigned long long* mm) {
    const long long n = bin_count(B, n_dev);
    unsigned long long lo0 = ~0ULL, hi0 = 0ULL, lo1 = ~0ULL, hi1 = 0ULL;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long a = ordered_bits(x[i * f + k0]), b = ordered_bits(x[i * f + k1]);
        lo0 = a < lo0 ? a : lo0; hi0 = a > hi0 ? a : hi0; lo1 = b < lo1 ? b : lo1; hi1 = b > hi1 ? b : hi1;
    }
    // warp first (65536 threads hammering four words took 170 us), then one set of atomics per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a0 = __shfl_xor_sync(0xffffffffu, lo0, o), a1 = __shfl_xor_sync(0xffffffffu, hi0, o);
        const unsigned long long b0 = __shfl_xor_sync(0xffffffffu, lo1, o), b1 = __shfl_xor_sync(0xffffffffu, hi1, o);
        lo0 = a0 < lo0 ? a0 : lo0; hi0 = a1 > hi0 ? a1 : hi0; lo1 = b0 < lo1 ? b0 : lo1; hi1 = b1 > hi1 ? b1 : hi1;
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(mm + 0, lo0); atomicMax(mm + 1, hi0); atomicMin(mm + 2, lo1); atomicMax(mm + 3, hi1); }
}
__global__ void bin_hist_kernel(const double* x, long long B, const int* n_dev, int f, int k0, int k1, int nb,
                                const unsigned long long* mm, int* keys, int* hist) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= bin_count(B, n_dev)) return;
    const double lo0 = from_ordered_bits(mm[0]), hi0 = from_ordered_bits(mm[1]), lo1 = from_ordered_bits(mm[2]), hi1 = from_ordered_bits(mm[3]);
    const double top = (double)nb - 1e-3;
    const double s0 = (hi0 > lo0) ? top / (hi0 - lo0) : 0.0, s1 = (hi1 > lo1) ? top / (hi1 - lo1) : 0.0;
    const double a = (x[i * f + k0] - lo0) * s0, b = (x[i * f + k1] - lo1) * s1;
    int b0 = (a >= 0.0 && a < (double)nb) ? (int)a : 0, b1 = (b >= 0.0 && b < (double)nb) ? (int)b : 0;      // (NaN guesses land in bin 0)
    if (b0 & 1) b1 = nb - 1 - b1;                    // boustrophedon: consecutive bins stay neighbours across a row change
    const int key = b0 * nb + b1;
    keys[i] = key;
    atomicAdd(hist + key, 1);
}
__global__ void bin_scan_kernel(int* hist, int nbins) {      // exclusive scan of nbins (<= 65536, multiple of 1024) counters, one CTA
    __shared__ int part[1024];
    const int t = threadIdx.x, per = nbins / 1024;
    int sum = 0;
    for (int k = 0; k < per; ++k) sum += hist[t * per + k];
    part[t] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = (t >= o) ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int run = part[t] - sum;
    for (int k = 0; k < per; ++k) { const int c = hist[t * per + k]; hist[t * per + k] = run; run += c; }
}
__global__ void bin_scatter_kernel(const int* keys, long long B, const int* n_dev, int* cursor, int* order) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= bin_count(B, n_dev)) return;
    order[atomicAdd(cursor + keys[i], 1)] = (int)i;
}
// live list and its gathered variables in the new order (through a temporary, then back)
__global__ void bin_permute_kernel(const int* order, long long B, const int* n_dev, int f, const int* act, const double* xc,
                                   int* act_tmp, double* xc_tmp) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= bin_count(B, n_dev)) return;
    const int src = order[i];
    act_tmp[i] = act[src];
    for (int k = 0; k < f; ++k) xc_tmp[i * f + k] = xc[(long long)src * f + k];
}
__global__ void bin_copyback_kernel(long long B, const int* n_dev, int f, const int* act_tmp, const double* xc_tmp, int* act, double* xc) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= bin_count(B, n_dev)) return;
    act[i] = act_tmp[i];
    for (int k = 0; k < f; ++k) xc[i * f + k] = xc_tmp[i * f + k];
}

__global__ void adam_init_kernel(long long B, int f, double N, double lr, double* m, double* v, double* vhat,
                                 double* lrv, double* last, double* mn, double* best, int* err, int* alive, int* epochs) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int k = 0; k < f; ++k) {
        const long long o = b * f + k;
        m[o] = 0.0; v[o] = 0.0; vhat[o] = 0.0; best[o] = 0.0; err[o] = 0;      // Optimizer.py:133, 143
    }
    lrv[b] = lr; last[b] = N; mn[b] = N; alive[b] = 1; epochs[b] = 0;           // :135-138
}

__global__ void fill_nan_kernel(double* p, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = __longlong_as_double(0x7ff8000000000000LL);
}

// ---------------------------------------------------------------------------------- argmin
__global__ void argmin_kernel(const double* v, long long B, double* min_out, long long* idx_out) {
    __shared__ double sv[32];
    __shared__ long long si[32];
    double bv = __longlong_as_double(0x7ff0000000000000LL);
    long long bi = B;
    for (long long i = threadIdx.x; i < B; i += blockDim.x) {
        const double x = v[i];
        if (x < bv) { bv = x; bi = i; }         // strided scan keeps the smallest index per thread
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (sv[w] < bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
        *min_out = bv;
        *idx_out = (bi == B) ? 0 : bi;
    }
}

// ---------------------------------------------------------------------------------- FP64 peak
__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FP64 tensor path (mma.sync m8n8k4): 8 independent accumulator pairs per warp
__global__ void dmma_peak_kernel(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int u = 0; u < 8; ++u) { c[u][0] = threadIdx.x * 1e-3 + u; c[u][1] = c[u][0] + 0.5; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(c[u][0]), "+d"(c[u][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += c[u][0] + c[u][1];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void sort_gather_kernel(const int* order, long long B, int f, const double* x, double* xs) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= B) return;
    const long long src = order[i];
    for (int k = 0; k < f; ++k) xs[i * f + k] = x[src * f + k];
}
__global__ void sort_scatter_kernel(const int* order, long long B, int f, const double* loss_s, const double* grad_s, const int* ts_s,
                                    double* loss, double* grad, int* ts) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= B) return;
    const long long dst = order[i];
    loss[dst] = loss_s[i];
    if (grad) for (int k = 0; k < f; ++k) grad[dst * f + k] = grad_s[i * f + k];
    if (ts) ts[dst] = ts_s[i];
}

}  // namespace

// order[0 .. B) = the points of x [B, f] by a coarse 2-D binning of (first, last) parameter (see the bin_* kernels)
int launch_locality_order(const double* x, int64_t B, int f, int* order, int* keys, int* hist, unsigned long long* mm, cudaStream_t st) {
    int nb = 32;
    while (nb < 256 && (long long)nb * nb * 32 < B) nb *= 2;
    static const unsigned long long mm0[4] = {~0ULL, 0ULL, ~0ULL, 0ULL};
    const int TPB = 256;
    const unsigned gB = (unsigned)((B + TPB - 1) / TPB);
    QTET_CUDA(cudaMemcpyAsync(mm, mm0, sizeof(mm0), cudaMemcpyHostToDevice, st));
    QTET_CUDA(cudaMemsetAsync(hist, 0, (size_t)nb * nb * sizeof(int), st));
    bin_minmax_kernel<<<64, TPB, 0, st>>>(x, (long long)B, nullptr, f, 0, f - 1, mm);
    bin_hist_kernel<<<gB, TPB, 0, st>>>(x, (long long)B, nullptr, f, 0, f - 1, nb, mm, keys, hist);
    bin_scan_kernel<<<1, 1024, 0, st>>>(hist, nb * nb);
    bin_scatter_kernel<<<gB, TPB, 0, st>>>(keys, (long long)B, nullptr, hist, order);
    count_launch(4);
    QTET_CUDA(cudaGetLastError());
    return QTET_OK;
}
int launch_sort_gather(const int* order, int64_t B, int f, const double* x, double* xs, cudaStream_t st) {
    sort_gather_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(order, (long long)B, f, x, xs);
    count_launch();
    QTET_CUDA(cudaGetLastError());
    return QTET_OK;
}
int launch_sort_scatter(const int* order, int64_t B, int f, const double* loss_s, const double* grad_s, const int* ts_s,
                        double* loss, double* grad, int* ts, cudaStream_t st) {
    sort_scatter_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(order, (long long)B, f, loss_s, grad_s, ts_s, loss, grad, ts);
    count_launch();
    QTET_CUDA(cudaGetLastError());
    return QTET_OK;
}

namespace {
}  // namespace

int launch_argmin(const double* v, int64_t B, double* min_out, int64_t* idx_out, cudaStream_t st) {
    argmin_kernel<<<1, 1024, 0, st>>>(v, (long long)B, min_out, (long long*)idx_out);
    count_launch();
    QTET_CUDA(cudaGetLastError());
    return QTET_OK;
}

}  // namespace qtet

using namespace qtet;

extern "C" int qtet_adam_run(qtet_problem* p, double* chis_dev, int64_t B, int target_site,
                             const int32_t* train_mask_host, int iterations, double lr, double beta_1,
                             double beta_2, double epsilon, int amsgrad, double tol,
                             double* best_vars_dev, double* best_loss_dev, int32_t* epochs_dev,
                             double* loss_hist_dev, double* var_hist_dev, void* stream) {
    if (!p || !chis_dev || !train_mask_host || !best_vars_dev || !best_loss_dev) { set_error("null argument"); return QTET_EINVAL; }
    if (B <= 0 || B > 0x7fffffffLL) { set_error("batch must be in 1..2^31-1"); return QTET_EINVAL; }
    if (target_site < 0 || target_site >= p->f) { set_error("target_site out of range"); return QTET_EINVAL; }
    if (iterations < 0) { set_error("negative iterations"); return QTET_EINVAL; }
    QTET_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int f = p->f;
    unsigned mask = 0;
    for (int k = 0; k < f; ++k) if (train_mask_host[k]) mask |= 1u << k;

    // workspace layout (all on the handle's scratch allocation)
    const size_t Bf = (size_t)B * f;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_m = take(Bf * 8), o_v = take(Bf * 8), o_vh = take(Bf * 8), o_lr = take(B * 8), o_last = take(B * 8),
                 o_xc = take(Bf * 8), o_loss = take(B * 8), o_grad = take(Bf * 8), o_err = take(Bf * 4),
                 o_alive = take(B * 4), o_active = take(2 * (size_t)B * 4), o_ep = take(B * 4), o_cnt = take(256),
                 o_order = take(B * 4), o_keys = take(B * 4), o_hist = take(65536 * 4), o_mm = take(256),
                 o_acttmp = take(B * 4), o_xctmp = take(Bf * 8);
    int rc = ensure_scratch(p, off);
    if (rc) return rc;
    char* base = (char*)p->scratch;
    double *m = (double*)(base + o_m), *v = (double*)(base + o_v), *vh = (double*)(base + o_vh), *lrv = (double*)(base + o_lr),
           *last = (double*)(base + o_last), *xc = (double*)(base + o_xc), *lossc = (double*)(base + o_loss),
           *gradc = (double*)(base + o_grad);
    int *err = (int*)(base + o_err), *alive = (int*)(base + o_alive), *active = (int*)(base + o_active),
        *ep = (int*)(base + o_ep), *cnt = (int*)(base + o_cnt);
    // pinned host words for the live counts + their events: allocated once per handle
    constexpr int kRing = qtet_problem::kPollRing;
    if (!p->h_poll) {
        QTET_CUDA(cudaMallocHost(&p->h_poll, kRing * sizeof(int)));
        QTET_CUDA(cudaStreamCreateWithFlags(&p->s_poll, cudaStreamNonBlocking));
        for (int i = 0; i < kRing; ++i) {
            QTET_CUDA(cudaEventCreateWithFlags(&p->ev_cnt[i], cudaEventDisableTiming));
            QTET_CUDA(cudaEventCreateWithFlags(&p->ev_poll[i], cudaEventDisableTiming));
        }
    }

    const int TPB = 256;
    const unsigned gB = (unsigned)((B + TPB - 1) / TPB);
    adam_init_kernel<<<gB, TPB, 0, st>>>((long long)B, f, (double)p->N, lr, m, v, vh, lrv, last, best_loss_dev, best_vars_dev, err, alive, ep);
    count_launch();
    if (loss_hist_dev && iterations > 0) {
        const long long n = (long long)iterations * B;
        fill_nan_kernel<<<(unsigned)((n + TPB - 1) / TPB), TPB, 0, st>>>(loss_hist_dev, n);
        count_launch();
    }
    if (var_hist_dev && iterations >= 10) {
        const long long n = (long long)(iterations / 10) * B * f;
        fill_nan_kernel<<<(unsigned)((n + TPB - 1) / TPB), TPB, 0, st>>>(var_hist_dev, n);
        count_launch();
    }
    AdamArgs a;
    a.f = f; a.xc = xc; a.loss = lossc; a.grad = gradc;
    a.x = chis_dev; a.m = m; a.v = v; a.vhat = vh; a.lr = lrv; a.last_loss = last; a.min_loss = best_loss_dev;
    a.best_vars = best_vars_dev; a.err_count = err; a.alive = alive; a.epochs = ep; a.train_mask = mask;
    a.beta_1 = beta_1; a.beta_2 = beta_2; a.epsilon = epsilon; a.tol = tol;
    a.amsgrad = amsgrad; a.B = B; a.loss_hist = loss_hist_dev; a.var_hist = var_hist_dev;

    // The loss+grad kernels of the direct path read the number of live guesses from device memory, so the epoch loop
    // never waits for the GPU: the host only keeps an UPPER BOUND of the live count (it can only fall) to size the grids,
    // refreshed from a pinned word that trails the device by kLag epochs, and stops once that word reads zero.
    const bool devcount = (p->path == QTET_PATH_DIRECT) && direct_eligible(p->view) && B <= (1LL << 22) && !std::getenv("QTET_ADAM_SYNC");
    int epoch = 0, executed = -1;
    if (devcount) {
        constexpr int kLag = 4;
        static_assert(kLag < kRing, "poll ring too small");
        QTET_CUDA(cudaMemsetAsync(cnt, 0, 2 * sizeof(int), st));
        // The live list is put in locality order (see bin_* above) at the start and again every `resort` epochs: the guesses move
        // by ~lr per epoch, so neighbours of the list drift apart.  QTET_ADAM_RESORT=0 keeps the caller's order, =n sets the period.
        static const int resort = [] { const char* e = std::getenv("QTET_ADAM_RESORT"); return e ? std::atoi(e) : 8; }();
        int* order = (int*)(base + o_order);
        int* keys = (int*)(base + o_keys);
        int* hist = (int*)(base + o_hist);
        unsigned long long* mm = (unsigned long long*)(base + o_mm);
        int* act_tmp = (int*)(base + o_acttmp);
        double* xc_tmp = (double*)(base + o_xctmp);
        const bool sort_on = resort > 0 && B >= 4096;
        int nb = 16;
        while (nb < 256 && (long long)nb * nb * 32 < B) nb *= 2;      // 16-32 guesses per bin (64 x 64 measured best for 10^5 guesses)
        if (nb < 32) nb = 32;                                        // (the scan wants a multiple of 1024 bins)
        static const unsigned long long mm0[4] = {~0ULL, 0ULL, ~0ULL, 0ULL};
        long long n_bound = B;
        for (; epoch < iterations; ++epoch) {
            if (epoch >= kLag) {                        // live count that ENTERED epoch - kLag
                const int slot = (epoch - kLag) % kRing;
                cudaError_t e = cudaEventSynchronize(p->ev_poll[slot]);
                if (e != cudaSuccess) return cuda_fail(e, "adam: poll");
                const long long nlive = p->h_poll[slot];
                if (nlive == 0) { executed = epoch - kLag; break; }
                if (nlive < n_bound) n_bound = nlive;
            }
            const int cur = epoch & 1;
            int* act_cur = active + (size_t)cur * B;
            const int* act_prev = epoch ? active + (size_t)(1 - cur) * B : nullptr;
            compact_gather_kernel<<<(unsigned)((n_bound + TPB - 1) / TPB), TPB, 0, st>>>(
                alive, act_prev, n_bound, epoch ? cnt + (1 - cur) : nullptr, f, chis_dev, act_cur, xc, cnt + cur);
            count_launch();
            // the count leaves on a side stream so that the D2H copy never sits between two kernels of the epoch
            const int slot = epoch % kRing;
            QTET_CUDA(cudaEventRecord(p->ev_cnt[slot], st));
            QTET_CUDA(cudaStreamWaitEvent(p->s_poll, p->ev_cnt[slot], 0));
            QTET_CUDA(cudaMemcpyAsync(p->h_poll + slot, cnt + cur, sizeof(int), cudaMemcpyDeviceToHost, p->s_poll));
            QTET_CUDA(cudaEventRecord(p->ev_poll[slot], p->s_poll));
            // (only while many guesses are alive: with a small list an epoch is a latency chain per kernel, and the six extra launches
            // cost more than the ordering saves -- 12 500 guesses per GPU: 0.205 s without, 0.223 s with)
            if (sort_on && epoch % resort == 0 && n_bound >= 32768) {
                const unsigned gN = (unsigned)((n_bound + TPB - 1) / TPB);
                QTET_CUDA(cudaMemcpyAsync(mm, mm0, sizeof(mm0), cudaMemcpyHostToDevice, st));
                QTET_CUDA(cudaMemsetAsync(hist, 0, (size_t)nb * nb * sizeof(int), st));
                bin_minmax_kernel<<<64, TPB, 0, st>>>(xc, n_bound, cnt + cur, f, 0, f - 1, mm);
                bin_hist_kernel<<<gN, TPB, 0, st>>>(xc, n_bound, cnt + cur, f, 0, f - 1, nb, mm, keys, hist);
                bin_scan_kernel<<<1, 1024, 0, st>>>(hist, nb * nb);
                bin_scatter_kernel<<<gN, TPB, 0, st>>>(keys, n_bound, cnt + cur, hist, order);
                bin_permute_kernel<<<gN, TPB, 0, st>>>(order, n_bound, cnt + cur, f, act_cur, xc, act_tmp, xc_tmp);
                bin_copyback_kernel<<<gN, TPB, 0, st>>>(n_bound, cnt + cur, f, act_tmp, xc_tmp, act_cur, xc);
                count_launch(6);
                QTET_CUDA(cudaGetLastError());
            }
            rc = loss_grad_dispatch(p, xc, n_bound, target_site, lossc, gradc, nullptr, nullptr, st, nullptr, cnt + cur);
            if (rc) return rc;
            a.n_active = n_bound; a.n_dev = cnt + cur; a.cnt_next = cnt + (1 - cur); a.active = act_cur;
            const int t = epoch + 1;
            a.s2 = std::sqrt(1.0 - std::pow(beta_2, (double)t));
            a.s1 = 1.0 - std::pow(beta_1, (double)t);
            a.epoch = epoch;
            // cnt[1 - cur] is zeroed by this kernel: the copy of the previous epoch's count must have left it
            if (epoch) QTET_CUDA(cudaStreamWaitEvent(st, p->ev_poll[(epoch - 1) % kRing], 0));
            adam_step_kernel<<<(unsigned)((n_bound + TPB - 1) / TPB), TPB, 0, st>>>(a);
            count_launch();
        }
        if (executed < 0) {                             // ran out of iterations: settle the polls still in flight
            cudaError_t e = cudaStreamSynchronize(p->s_poll);
            if (e != cudaSuccess) return cuda_fail(e, "adam: poll sync");
            executed = epoch;
            for (int e2 = (epoch > kLag ? epoch - kLag : 0); e2 < epoch; ++e2)
                if (p->h_poll[e2 % kRing] == 0) { executed = e2; break; }
        }
    } else {
        long long n_active = B;
        for (; epoch < iterations; ++epoch) {
            // live list (compaction) -- sync point: the host needs the count to size the launches
            cudaMemsetAsync(cnt, 0, sizeof(int), st);
            compact_kernel<<<gB, TPB, 0, st>>>(alive, (long long)B, active, cnt);
            count_launch();
            cudaMemcpyAsync(p->h_poll, cnt, sizeof(int), cudaMemcpyDeviceToHost, st);
            cudaError_t e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return cuda_fail(e, "adam: sync");
            n_active = p->h_poll[0];
            if (n_active == 0) break;
            const long long nf = n_active * f;
            gather_kernel<<<(unsigned)((nf + TPB - 1) / TPB), TPB, 0, st>>>(active, n_active, f, chis_dev, xc);
            count_launch();
            rc = loss_grad_dispatch(p, xc, n_active, target_site, lossc, gradc, nullptr, nullptr, st);
            if (rc) return rc;
            a.n_active = n_active; a.n_dev = nullptr; a.cnt_next = nullptr; a.active = active;
            const int t = epoch + 1;
            a.s2 = std::sqrt(1.0 - std::pow(beta_2, (double)t));
            a.s1 = 1.0 - std::pow(beta_1, (double)t);
            a.epoch = epoch;
            adam_step_kernel<<<(unsigned)((n_active + TPB - 1) / TPB), TPB, 0, st>>>(a);
            count_launch();
        }
        executed = epoch;
    }
    if (epochs_dev) cudaMemcpyAsync(epochs_dev, ep, (size_t)B * 4, cudaMemcpyDeviceToDevice, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "adam: final sync");
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "adam: kernel");
    return executed;
}

// kind 0: DFMA loop, kind 1: DMMA m8n8k4 loop; best of the repetitions that fit the time budget [TFLOP/s]
static int fp64_peak_kind(int kind, double seconds_budget, double* tflops_out) {
    int device = 0, sms = 0;
    QTET_CUDA(cudaGetDevice(&device));
    QTET_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    double* out = nullptr;
    const int blocks = sms * 8, threads = 256;
    QTET_CUDA(cudaMalloc(&out, (size_t)blocks * threads * 8));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 2000;
    double best = 0.0, spent = 0.0;
    for (int rep = 0; rep < 50 && spent < seconds_budget; ++rep) {
        cudaEventRecord(e0);
        if (kind == 0) dfma_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
        else dmma_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
        count_launch();
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { cudaFree(out); return cuda_fail(e, "fp64 peak"); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        // DFMA: 64 fma per thread and iteration; DMMA: 8 mma of 2*8*8*4 flop per warp and iteration
        const double flops = (kind == 0) ? 2.0 * 64.0 * (double)iters * blocks * threads
                                         : 8.0 * 512.0 * (double)iters * blocks * (threads / 32);
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
        spent += ms * 1e-3;
        if (ms < 20.f) iters *= 2;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    *tflops_out = best;
    return QTET_OK;
}

extern "C" int qtet_fp64_peak2(int device, double seconds_budget, double* dfma_tflops, double* dmma_tflops) {
    if (!dfma_tflops || !dmma_tflops) { set_error("null argument"); return QTET_EINVAL; }
    QTET_ON_DEVICE(device);
    int rc = fp64_peak_kind(0, seconds_budget * 0.5, dfma_tflops);
    if (rc) return rc;
    return fp64_peak_kind(1, seconds_budget * 0.5, dmma_tflops);
}

extern "C" int qtet_fp64_peak(int device, double seconds_budget, double* tflops_out) {
    if (!tflops_out) { set_error("null argument"); return QTET_EINVAL; }
    double a = 0.0, b = 0.0;
    const int rc = qtet_fp64_peak2(device, seconds_budget, &a, &b);
    if (rc) return rc;
    *tflops_out = a > b ? a : b;                 // the roofline denominator is the HIGHER of the two FP64 paths
    return QTET_OK;
}
