/*
 * qtet.h -- C ABI of libqtet_b200.so: the B200-native (sm_100a) replacement for the
 * loss + gradient hot path of JAndronis/Quantum-Targeted-Energy-Transfer.
 *
 * The reference has no FFI / plugin registry (it is pure Python on TensorFlow); the boundary it
 * exposes for this path is its Python API.  Each entry point below names the reference callable
 * (file:line under /root/reference) whose work it replaces; INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross the ABI;
 *   - the caller owns every buffer; the library owns only the opaque problem handle (and a
 *     scratch workspace hanging off it, grown on demand);
 *   - "_dev" arguments are device pointers on the handle's device; "_host" are host pointers;
 *   - calls taking a `stream` are asynchronous on that CUDA stream (a `cudaStream_t` passed as
 *     void*, NULL = the legacy default stream) and never synchronise behind the caller's back;
 *     the *_host convenience calls copy H2D/D2H themselves and return after completion;
 *   - return value 0 = success, negative = failure (QTET_E*); qtet_last_error() returns a
 *     thread-local human-readable message.  No exception crosses the ABI.
 *   - one process per GPU; a handle is bound to the device it was created on and may be used
 *     from one thread and on one stream at a time (its workspaces are ordered against the stream
 *     of the call that uses them); entry points switch to the handle's device and restore the caller's;
 *   - Fock dimensions up to 1536 (d <= 115: register / shared-memory kernels; above: the generic kernel with its
 *     matrices in global memory); larger systems are refused by qtet_problem_create.
 */
#ifndef QTET_H_
#define QTET_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QTET_OK            0
#define QTET_EINVAL       -1   /* bad argument (site out of range, null pointer, ...) */
#define QTET_ECUDA        -2   /* a CUDA runtime call or kernel launch failed */
#define QTET_ENOMEM       -3   /* workspace / handle allocation failed */
#define QTET_EUNSUPPORTED -4   /* problem too large for this build (Fock dimension limit) */

typedef struct qtet_problem qtet_problem;

/* Kernel family selection for qtet_set_path (diagnostics / parity tests). */
#define QTET_PATH_AUTO    0    /* direct path when eligible, else the split paths, else fused generic */
#define QTET_PATH_GENERIC 1    /* fused CTA-per-point kernel, any system */
#define QTET_PATH_SPLIT   2    /* thread-per-point scalar QL + rotation replay (register path for tridiagonal d<=32,
                                  shared-memory replay for d<=128 incl. dense H after Householder) */
#define QTET_PATH_DIRECT  3    /* tridiagonal d<=32 with non-zero hopping: eigenvalues by scalar QL, eigenvectors by
                                  two-sided three-term recurrences (twisted factorisation), no rotation stream */

/*
 * Build the chi-independent part of the problem once: Fock basis in the reference's order
 * (Loss.derive, HamiltonianLoss.py:62-110), the linear part of the diagonal and d H_mm / d chi_k
 * = n_k(m)^2 / 2 (:131-134), the constant hopping elements (:136-163), and the time grid
 * linspace(0, max_t, timesteps) (:221).  Replaces Loss.__init__ (HamiltonianLoss.py:20-47).
 *   omegas : [sites] host.   device : CUDA device ordinal.
 */
int qtet_problem_create(int max_N, int sites, const double* omegas, double coupling, double max_t,
                        int timesteps, int device, qtet_problem** out);
void qtet_problem_destroy(qtet_problem* p);

/* Fock dimension C(N+f-1, f-1) (HamiltonianLoss.py:39-40); <0 on a null handle. */
int qtet_problem_dim(const qtet_problem* p);
/* Basis occupations, row-major [dim, sites] host buffer  (Loss.states, HamiltonianLoss.py:43,62-110). */
int qtet_problem_states(const qtet_problem* p, double* states_host);
/* Time grid [timesteps] host buffer (HamiltonianLoss.py:221). */
int qtet_problem_time_grid(const qtet_problem* p, double* t_host);
/* Dense Hamiltonian of ONE point, built on the device and copied back: [dim, dim] row-major host
 * buffer (Loss.createHamiltonian, HamiltonianLoss.py:168-176).  Parity/debug aid. */
int qtet_hamiltonian_host(qtet_problem* p, const double* chis_host, double* H_host);

/*
 * Size the handle's workspaces for batches of up to B points NOW.  After it, qtet_loss_grad / qtet_trace / qtet_adam_run
 * with batches <= B on the direct path neither allocate nor synchronise (they are purely asynchronous on the caller's
 * stream, as the conventions above promise); without it a workspace grows on the first call that needs more room, and
 * growing means cudaDeviceSynchronize + cudaFree + cudaMalloc.  The split / large-d paths size theirs on first use.
 */
int qtet_problem_reserve(qtet_problem* p, int64_t B);

/*
 * Evaluate the batches of qtet_loss_grad (device-pointer entry point, B >= 4096) in LOCALITY order: the points are ordered by a
 * coarse 2-D binning of (first, last) parameter on the device, evaluated, and the results scattered back to the caller's order.
 * A warp of the kernels holds consecutive points and costs what its most expensive point costs, so unordered (random) batches
 * -- the reference's 'bins' initial guesses, solver_mp.py:17-84 -- run ~20 % faster this way; every result is bit-identical to the
 * unordered call.  Off by default (grid sweeps are ordered already); qtet_adam_run always orders its live list.
 */
int qtet_set_locality_sort(qtet_problem* p, int on);

/* Choose the kernel family (QTET_PATH_*); QTET_EUNSUPPORTED if the problem is not eligible. */
int qtet_set_path(qtet_problem* p, int path);
/* Which family the next call will use (QTET_PATH_GENERIC, QTET_PATH_SPLIT or QTET_PATH_DIRECT). */
int qtet_get_path(const qtet_problem* p);

/*
 * THE HOT PATH.  For each of B parameter points chi[b, 0..f-1]:
 *   H build -> eigendecomposition -> <n_site(t)> on the time grid -> loss -> d loss / d chi.
 * Replaces Optimizer.get_grads (Optimizer.py:85-91) = Loss.__call__ (HamiltonianLoss.py:49-60,
 * 218-233) under tf.GradientTape, for B points at once (the reference forks one process per
 * point: solver_mp.py:157-172).
 *   chis_dev   : [B, f] row-major float64
 *   target_site: 0..f-1; site == f-1 -> loss = N - max_t <n>, else loss = min_t <n>  (:227-231)
 *   loss_dev   : [B] float64
 *   grad_dev   : [B, f] float64, d loss / d chi_k with the arg-max/arg-min sample held fixed; may be NULL
 *   tstar_dev  : [B] int32, index of that sample on the time grid; may be NULL
 */
int qtet_loss_grad(qtet_problem* p, const double* chis_dev, int64_t B, int target_site,
                   double* loss_dev, double* grad_dev, int32_t* tstar_dev, void* stream);

/* Same through HOST buffers (H2D, kernels, D2H inside the call; returns when results are in host
 * memory).  This is the call a Python host makes when it holds numpy arrays. */
int qtet_loss_grad_host(qtet_problem* p, const double* chis_host, int64_t B, int target_site,
                        double* loss_host, double* grad_host, int32_t* tstar_host);

/* Loss.__call__(..., single_value=False) (HamiltonianLoss.py:232-233): trace_dev [B, timesteps]. */
int qtet_trace(qtet_problem* p, const double* chis_dev, int64_t B, int target_site,
               double* trace_dev, void* stream);

/*
 * Batched per-guess optimiser: Optimizer.train (Optimizer.py:99-242) for B initial guesses at
 * once, with the Keras Adam update and every stop/revert/lr rule of the reference applied per
 * guess on the device.  One entry of `train_mask` per site (non-zero = trained, Optimizer.py:120-130).
 *   chis_dev      : [B, f] initial guesses in, final variables out
 *   best_vars_dev : [B, f] variables at the lowest loss seen  (results['best_vars'], :172-175)
 *   best_loss_dev : [B]    min over recorded losses           (np.min(loss_data), :321)
 *   epochs_dev    : [B] int32, number of loss+grad evaluations spent on the guess; may be NULL
 *   loss_hist_dev : [iterations, B] float64 loss per epoch (NaN after a guess stopped); may be NULL
 *   var_hist_dev  : [iterations/10, B, f] variables sampled every 10 steps (:189-192); may be NULL
 * Returns the number of epochs executed (>= 0) or a negative error.
 */
int qtet_adam_run(qtet_problem* p, double* chis_dev, int64_t B, int target_site,
                  const int32_t* train_mask_host, int iterations, double lr, double beta_1,
                  double beta_2, double epsilon, int amsgrad, double tol,
                  double* best_vars_dev, double* best_loss_dev, int32_t* epochs_dev,
                  double* loss_hist_dev, double* var_hist_dev, void* stream);

/* arg-min over guesses (solver_mp.py:182-184): out_dev[0] = min value, out_dev[1] = index (as a
 * double-encoded int64 bit pattern is avoided: index is written to idx_dev).  First minimum wins. */
int qtet_argmin(const double* values_dev, int64_t B, double* min_dev, int64_t* idx_dev, void* stream);

/* Measured FP64 throughput of the device, TFLOP/s: the higher of a dependent-free DFMA loop and an mma.sync m8n8k4 (FP64
 * tensor path, DMMA) loop.  Used as the roofline denominator because MEASURED_PEAKS.json carries no FP64 figure.
 * qtet_fp64_peak2 returns both figures. */
int qtet_fp64_peak(int device, double seconds_budget, double* tflops_out);
int qtet_fp64_peak2(int device, double seconds_budget, double* dfma_tflops_out, double* dmma_tflops_out);

/* Per-kernel timing (CUDA events on the launching streams).  kind 0: scalar QL kernel, 1: rotation-replay /
 * evolution / gradient kernel (the dominant one), 2: fused generic kernel, or the
 * Householder kernel of the dense large-d path.  enable(on=1) clears the records;
 * read() synchronises the device and returns, per kind, the summed duration [ms], the number of launches and
 * the number of parameter points they processed (arrays of 3). */
int qtet_profile_enable(qtet_problem* p, int on);
int qtet_profile_read(qtet_problem* p, double* ms_sum, int64_t* launches, int64_t* points);

/* Cumulative health counters of the handle's kernels (synchronises the device): out4[0] = points the direct path handed to
 * the fused generic kernel (numerically degenerate spectrum, recurrence overflow, QL not converged), out4[1] = points whose
 * rotation stream overflowed on the split paths (also recomputed by the generic kernel), out4[2] = eigen-solves of the
 * generic kernel that hit its iteration limit -- the only case in which a result may be inaccurate; tests and bench.py
 * assert it is 0 -- out4[3] reserved.  reset != 0 clears the counters. */
int qtet_stats(qtet_problem* p, int64_t* out4, int reset);

/* Number of kernel launches this library has issued in the calling process so far. */
int64_t qtet_launch_count(void);

const char* qtet_last_error(void);
const char* qtet_version(void);

#ifdef __cplusplus
}
#endif
#endif /* QTET_H_ */
