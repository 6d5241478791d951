// Direct path for tridiagonal Hamiltonians with Fock dimension d <= 32 and non-zero hopping (every dimer up to
// N = 31; BASELINE's headline is the dimer N = 20, d = 21).
//
// The eigenvectors are NOT accumulated from plane rotations (qtet_split.cu does that: ~240 rotations x 21 rows per
// point plus a 6.7 KB/point rotation stream through HBM).  They are computed column by column from the converged
// eigenvalues, O(d) work per eigenvector:
//
//   eig_scalar_kernel (K_E)  one THREAD per parameter point: implicit-shift QL on (diag, off) only, eigenvalues to
//        full precision, sorted ascending, 8 d bytes per point written.
//   direct_kernel (K_D)      G lanes per point.  Lane q owns the eigen-indices i = q, q+G, ...: for each it runs the
//        three-term recurrence of (T - lam) z = 0 from the top (p_k) and from the bottom (q_k) and joins the two at
//        the row r that maximises |p_r q_r| -- the discrete Wronskian b_k (p_k q_{k+1} - p_{k+1} q_k) is constant in k,
//        so gamma_r = W / (p_r q_r) is the smallest residual any twist index can give (the twisted factorisation of
//        Parlett & Dhillon, written without divisions).  The columns go through shared memory once (transpose: the
//        later phases want ROWS of V in registers), neighbouring eigenvalues closer than 2^-12 ||T|| are
//        re-orthogonalised there (modified Gram-Schmidt inside the cluster), and points whose spectrum is numerically
//        degenerate (gap < 2^-34 ||T||, where two recurrences return the same vector) or whose recurrences left the
//        FP64 range are appended to a list and recomputed by the fused generic kernel -- no host synchronisation.
//        Then the shared phases of qtet_split.cuh: time evolution, loss, arg-max, analytic gradient.
//
// Accuracy (tools/direct_prototype.py, numpy restatement against LAPACK eigh and mpmath): on the whole 1024 x 1024
// config-2 grid the trace differs from the eigh-based one by at most 3e-12 (min gap 1.2e-2, no point needs the
// fallback); with the Gram-Schmidt step the error stays below 1e-12 down to relative gaps of 4e-12.
#include <cstdlib>

#include "qtet_split.cuh"

namespace qtet {

using namespace splitk;

namespace {

constexpr int kWarpsE = 4;                              // warps per CTA, eigenvalue kernel
constexpr double kGapGS = 2.44140625e-4;                // 2^-12: re-orthogonalise neighbours closer than this (x ||T||)
constexpr double kGapFB = 5.820766091346741e-11;        // 2^-34: numerically degenerate -> generic kernel
constexpr int kMaxF = 8;
#ifndef QTET_CSIG
#define QTET_CSIG 1e-14
#endif
// Eigenvectors with |c_i| = |V[0][i]| below this carry no weight in psi(t) and are left out of the evolution: dropping them
// moves psi by at most d * 1e-14 and <n(t)> by at most 2 N d 1e-14 (8e-12 at N = 20, 2e-11 at N = 31) -- two orders below the
// 1e-10 N parity tolerance.  On the config-2 grid 11.7 of the 21 columns survive on average (14.4 with 1e-19).
constexpr double kCsig = QTET_CSIG;

// Shifted diagonal of H.  The two kernels must see bit-identical values, hence explicit fma / no contraction freedom:
// a_m = lin_m + sum_k chi_k quad_mk  (HamiltonianLoss.py:131-134 with lin = sum_k omega_k n_k, quad = n_k^2 / 2).
__device__ __forceinline__ double diag_entry(const double* lin, const double* quad, const double* chi, int f, int m) {
    double acc = lin[m];
#pragma unroll
    for (int k = 0; k < kMaxF; ++k)
        if (k < f) acc = __fma_rn(chi[k], quad[m * f + k], acc);
    return acc;
}

#ifndef QTET_PREFETCH_NEXT
#define QTET_PREFETCH_NEXT 0  // 1: prefetch the next task's chi / eigenvalues into L1 at the top of a task
#endif
#ifndef QTET_KD_MINB
#define QTET_KD_MINB 4     // resident CTAs per SM the exact-size d = 21 instantiation of K_D is compiled for (register cap 65536 / (64 MINB)).
                           // 5 (200 registers, 116 B of spills, 10 warps per SM) measured SLOWER: 8.25 vs 7.28 ms per grid.
#endif
#ifndef QTET_REC_FUSED
#define QTET_REC_FUSED 0   // 1: bottom-up and top-down recurrences of K_D in one unrolled loop, p_k in registers (see there).
                           // Bit-identical results, 27 % fewer executed instructions -- and measured SLOWER (8.04 vs 7.27 ms per
                           // grid): like every unrolled variant of this kernel it loses on instruction fetch what it saves.
#endif
#ifndef QTET_REC_ROLLED
#define QTET_REC_ROLLED 1  // 1: the three recurrence passes of K_D as rolled loops (small code), 0: fully unrolled with register-resident constants
#endif
#ifndef QTET_KE_PWK
#define QTET_KE_PWK 0     // 1: square-root-free rational QL (dsterf recurrence) in K_E; 0: rotation-based implicit QL
#endif
#if QTET_KE_PWK
// ------------------------------------------------------------------------------------------------
// K_E: thread-per-point QL, eigenvalues only.  diag/off in shared memory as [2k + {0,1}][lane].
//
// Square-root-free rational QL (Pal-Walker-Kahan, the LAPACK dsterf recurrence) on the SQUARED off-diagonals: a plane costs
// two reciprocals and ~12 FP64 operations with a dependency chain of  r = p + e2 -> 1/r -> c -> gamma -> p'  (~65 cycles)
// where the rotation-based sweep of qtet_split.cu needs rsqrt and ~22 operations (~120 cycles).  The kernel is a pure latency
// chain (one point per thread, five warps per scheduler), so the chain length IS its run time.
// ------------------------------------------------------------------------------------------------
struct PwkLane {
    double* arr;       // diag k at arr[64 k], SQUARED off-diagonal k at arr[64 k + 32]
    int d;
    __device__ __forceinline__ double& dg(int k) const { return arr[64 * k]; }
    __device__ __forceinline__ double& e2(int k) const { return arr[64 * k + 32]; }
    // off-diagonal k negligible:  |e_k| <= eps (|d_k| + |d_k+1|), tested on the squares
    __device__ __forceinline__ static bool tiny(double e2v, double dk, double dk1) {
        const double tst = kEps * (fabs(dk) + fabs(dk1));
        return e2v <= tst * tst;
    }
    __device__ __forceinline__ unsigned scan_small() const {
        unsigned small = 1u << (d - 1);
        for (int k = 0; k < d - 1; ++k)
            if (tiny(e2(k), dg(k), dg(k + 1))) small |= 1u << k;
        return small;
    }
};

__global__ void __launch_bounds__(kWarpsE * 32) eig_scalar_kernel(SplitArgs a) {
    extern __shared__ double sm[];
    const ProblemView& pv = a.pv;
    const int d = pv.d, f = pv.f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PwkLane t;
    t.arr = sm + (size_t)warp * 2 * d * 32 + lane;
    t.d = d;
    const long long n = effective_n(a);
    const long long nbatch = (n + 31) / 32;
    const unsigned below_top = lowmask(d - 1);
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.fb_count = 0;    // K_D of this chunk (stream-ordered after us) appends to the list
    for (long long batch = (long long)blockIdx.x * kWarpsE + warp; batch < nbatch; batch += (long long)gridDim.x * kWarpsE) {
        const long long p = batch * 32 + lane;
        const bool valid = p < n;
        const long long pc = valid ? p : n - 1;
        double chi[kMaxF];
#pragma unroll
        for (int k = 0; k < kMaxF; ++k) chi[k] = (k < f) ? a.chis[pc * f + k] : 0.0;
        double sum = 0.0;
        for (int m = 0; m < d; ++m) {
            const double acc = diag_entry(pv.lin, pv.quad, chi, f, m);
            const double off = (m < d - 1) ? pv.offd[m] : 0.0;
            t.dg(m) = acc;
            t.e2(m) = off * off;
            sum = __dadd_rn(sum, acc);
        }
        const double mean = __ddiv_rn(sum, (double)d);       // global phase: subtract the mean of the diagonal
        for (int m = 0; m < d; ++m) t.dg(m) = __dsub_rn(t.dg(m), mean);

        unsigned small = t.scan_small();
        int l = 0, iter = 0;
        bool failed = false;
        while (true) {
            const unsigned alive = (~small) & below_top & ~lowmask(l);
            if (alive == 0u) break;
            const int lnew = __ffs(alive) - 1;
            if (lnew != l) iter = 0;
            l = lnew;
            const int m = l + __ffs(small >> l) - 1;           // first negligible off-diagonal above l
            if (++iter > 60) { failed = true; break; }
            // Wilkinson shift from the leading 2 x 2 block of [l, m]
            double sigma;
            {
                const double pl = t.dg(l), rte = sqrt(t.e2(l));
                const double sg = (t.dg(l + 1) - pl) / (2.0 * rte);
                const double rr = sqrt(fma(sg, sg, 1.0));
                sigma = pl - rte / (sg + copysign(rr, sg));
            }
            double c = 1.0, s = 0.0, gamma = t.dg(m) - sigma, pp = gamma * gamma, dnext = 0.0;
            double* q = t.arr + 64 * (m - 1);                 // &dg(i)
            double alpha = q[0], bb = q[32];                  // d_i, e2_i of the coming plane, loaded a plane ahead
            unsigned fbit = 1u << m;                          // flag of the off-diagonal the plane finalises (i + 1)
            for (int i = m - 1; i >= l; --i, q -= 64) {
                const double* qn = (i > l) ? q - 64 : q;
                const double alpha_n = qn[0], bb_n = qn[32];
                const double r = pp + bb;
                if (!(r > 1e-290)) { failed = true; break; }  // underflow of the whole block: hand the point to the fallback
                const double e2new = s * r;                   // squared off-diagonal i + 1 (s = 0 on the first plane: e2(m) stays)
                const double rinv = fast_rcp(r), pinv = fast_rcp(pp);
                const double oldc = c;
                c = pp * rinv; s = bb * rinv;
                const double oldgam = gamma;
                gamma = fma(c, alpha - sigma, -(s * oldgam));
                const double dnew = oldgam + (alpha - gamma);
                q[64] = dnew;                                 // dg(i + 1)
                pp = (c != 0.0) ? (gamma * gamma) * (r * pinv) : oldc * bb;
                if (i != m - 1) {
                    q[96] = e2new;                            // e2(i + 1)
                    small = PwkLane::tiny(e2new, dnew, dnext) ? (small | fbit) : (small & ~fbit);
                }
                fbit >>= 1;
                dnext = dnew; alpha = alpha_n; bb = bb_n;
            }
            if (failed) break;
            const double e2l = s * pp, dl = sigma + gamma;
            t.e2(l) = e2l; t.dg(l) = dl;
            const unsigned bit = 1u << l;
            small = PwkLane::tiny(e2l, dl, dnext) ? (small | bit) : (small & ~bit);
        }
        // ascending order (insertion sort; the QL output is close to sorted)
        for (int i = 1; i < d; ++i) {
            const double x = t.dg(i);
            int j = i - 1;
            while (j >= 0 && t.dg(j) > x) { t.dg(j + 1) = t.dg(j); --j; }
            t.dg(j + 1) = x;
        }
        if (valid) {
            // a QL that did not converge poisons the point: K_D sees the NaN and hands the point to the generic kernel
            if (failed) t.dg(0) = __longlong_as_double(0x7ff8000000000000LL);
            for (int m = 0; m < d; ++m) a.ev[p * d + m] = t.dg(m);
        }
        __syncwarp();
    }
}

#else
// ------------------------------------------------------------------------------------------------
// K_E: thread-per-point QL, eigenvalues only.  diag/off in shared memory as [2k + {0,1}][lane].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsE * 32) eig_scalar_kernel(SplitArgs a) {
    extern __shared__ double sm[];
    const ProblemView& pv = a.pv;
    const int d = pv.d, f = pv.f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    QlLane t;
    t.arr = sm + (size_t)warp * 2 * d * 32 + lane;
    t.d = d;
    const long long n = effective_n(a);
    const long long nbatch = (n + 31) / 32;
    const unsigned below_top = lowmask(d - 1);
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.fb_count = 0;    // K_D of this chunk (stream-ordered after us) appends to the list
    for (long long batch = (long long)blockIdx.x * kWarpsE + warp; batch < nbatch; batch += (long long)gridDim.x * kWarpsE) {
        const long long p = batch * 32 + lane;
        const bool valid = p < n;
        const long long pc = valid ? p : n - 1;
        double chi[kMaxF];
#pragma unroll
        for (int k = 0; k < kMaxF; ++k) chi[k] = (k < f) ? a.chis[pc * f + k] : 0.0;
        double sum = 0.0;
        for (int m = 0; m < d; ++m) {
            const double acc = diag_entry(pv.lin, pv.quad, chi, f, m);
            t.dg(m) = acc;
            t.of(m) = (m < d - 1) ? pv.offd[m] : 0.0;
            sum = __dadd_rn(sum, acc);
        }
        const double mean = __ddiv_rn(sum, (double)d);       // global phase: subtract the mean of the diagonal
        for (int m = 0; m < d; ++m) t.dg(m) = __dsub_rn(t.dg(m), mean);

        unsigned small = t.scan_small();
        int l = 0, iter = 0;
        bool failed = false;
        while (true) {
            const unsigned alive = (~small) & below_top & ~lowmask(l);
            if (alive == 0u) break;
            const int lnew = __ffs(alive) - 1;
            if (lnew != l) iter = 0;
            l = lnew;
            const int m = l + __ffs(small >> l) - 1;
            if (++iter > 60) { failed = true; break; }
            double g = t.wilkinson_g(l, m);
            double s = 1.0, c = 1.0, pp = 0.0, dnext = 0.0;
            bool broke = false;
            double* q = t.arr + 64 * (m - 1);             // &dg(i)
            double d_hi = q[64], d_i = q[0], e_i = q[32];
            unsigned fbit = 1u << m;
            for (int i = m - 1; i >= l; --i, q -= 64) {
                const double* qn = (i > l) ? q - 64 : q;
                const double d_n = qn[0], e_n = qn[32];
                const double fo = s * e_i, bo = c * e_i;
                const double r2 = fma(fo, fo, g * g);
                if (r2 < kTinyR2) { q[96] = 0.0; q[64] = d_hi - pp; t.of(m) = 0.0; broke = true; break; }
                const double rinv = rsqrt_fast(r2);
                const double r = r2 * rinv;
                s = fo * rinv; c = g * rinv;
                g = d_hi - pp;
                const double rr = fma(d_i - g, s, 2.0 * c * bo);
                pp = s * rr;
                const double dnew = g + pp;
                q[64] = dnew;                              // dg(i+1)
                g = fma(c, rr, -bo);
                q[96] = r;                                 // of(i+1) (of(m) is reset below)
                small = (r <= kEps * (fabs(dnew) + fabs(dnext))) ? (small | fbit) : (small & ~fbit);
                fbit >>= 1;
                dnext = dnew; d_hi = d_i; d_i = d_n; e_i = e_n;
            }
            if (broke) { small = t.scan_small(); continue; }
            const double dlnew = d_hi - pp;                // d_hi = dg(l) after the last plane
            t.dg(l) = dlnew; t.of(l) = g; t.of(m) = 0.0;
            small |= 1u << m;
            const unsigned bit = 1u << l;
            if (fabs(g) <= kEps * (fabs(dlnew) + fabs(dnext))) small |= bit; else small &= ~bit;
        }
        // ascending order (insertion sort; the QL output is close to sorted)
        for (int i = 1; i < d; ++i) {
            const double x = t.dg(i);
            int j = i - 1;
            while (j >= 0 && t.dg(j) > x) { t.dg(j + 1) = t.dg(j); --j; }
            t.dg(j + 1) = x;
        }
        if (valid) {
            // a QL that did not converge poisons the point: K_D sees the NaN and hands the point to the generic kernel
            if (failed) t.dg(0) = __longlong_as_double(0x7ff8000000000000LL);
            for (int m = 0; m < d; ++m) a.ev[p * d + m] = t.dg(m);
        }
        __syncwarp();
    }
}

#endif

template <int D, int G>
struct DLayout {
    using L = BLayout<D, G>;
    static constexpr int PPW = L::PPW;
    static constexpr int LDV = (D + 2) & ~1;                          // even row stride: 16 B aligned pairs
    static constexpr int VT = PPW * D * LDV;                           // double [PPW][D rows][LDV]
    static constexpr int PHASED = (VT > L::GRAD ? VT : L::GRAD);      // the transpose buffer aliases the gradient buffers
    static constexpr int PER_WARP = ((PHASED + L::ZB + PPW * D) + 1) & ~1;
    static constexpr int CONSTS = ((D * (3 + kMaxF)) + 1) & ~1;       // per CTA: lin, quad, b_k, 1 / b_k
};

// ------------------------------------------------------------------------------------------------
// K_D: G lanes per point; eigenvectors by two-sided recurrences, then the shared phases.
// ------------------------------------------------------------------------------------------------
// EXACT: the Fock dimension equals the template size (no padded rows / eigen-indices), d is a compile-time constant.
// Register cap for MINB CTAs of WARPS warps per SM, given explicitly: `__launch_bounds__(threads, MINB)` rounds the CTA up when it
// derives the cap (168 instead of 200 registers for five CTAs of two warps).
constexpr int kd_reg_cap(int warps, int minb) {
    int r = 65536 / (minb * warps * 32);
    r = r / 8 * 8;
    return r > 255 ? 255 : r;
}
template <int D, int G, int WARPS, int MINB, bool EXACT>
__global__ void __launch_bounds__(WARPS * 32) __maxnreg__(kd_reg_cap(WARPS, MINB)) direct_kernel(SplitArgs a) {
    using L = BLayout<D, G>;
    using DL = DLayout<D, G>;
    constexpr int R = L::R, RZ = L::RZ, PPW = L::PPW, LDV = DL::LDV;
    extern __shared__ double sm[];
    const ProblemView& pv = a.pv;
    const int d = EXACT ? D : pv.d, f = pv.f, T = pv.T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / G, q = lane % G;
    // CTA-wide constants
    double* slin = sm;                     // [D]
    double* squad = sm + D;                // [D][f]
    double* sb = sm + D * (1 + kMaxF);     // [D] b_k = H[k+1,k] (0 for k >= d-1)
    double* sib = sb + D;                  // [D] 1 / b_k        (0 for k >= d-1)
    for (int k = threadIdx.x; k < D; k += WARPS * 32) {
        sb[k] = (k < 32) ? a.boff[k] : 0.0;
        sib[k] = (k < 32) ? a.iboff[k] : 0.0;
        slin[k] = (k < d) ? pv.lin[k] : 0.0;
        for (int s = 0; s < f; ++s) squad[k * f + s] = (k < d) ? pv.quad[k * f + s] : 0.0;
    }
    __syncthreads();

    double* base = sm + DL::CONSTS + (size_t)warp * DL::PER_WARP;
    double* vt = base;                                                       // [PPW][D][LDV]   (phase 1)
    double* lv = base;                                                       // [8][PPW][D]     (gradient)
    double2* part = reinterpret_cast<double2*>(base + L::LV);                // [PPW][G][DH]
    double* sbuf = base + L::LV;                                             // [PPW][D + NP]
    double2* zbuf = reinterpret_cast<double2*>(base + DL::PHASED);           // [2][2][PPW][D]
    double* abuf = base + DL::PHASED;                                        // [PPW][D]        (phase 1, aliases zbuf)
    double* cbuf = base + DL::PHASED + L::ZB;                                // [PPW][D]
    double* rnbuf = cbuf;                                                    // [PPW][D] 1/|z_i| (phase 1; the tail rewrites cbuf later)

    const double dt = (T > 1) ? pv.tgrid[1] : 0.0;
    double wsite[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int row = q * R + k;
        wsite[k] = (row < d) ? pv.occ[row * f + a.site] : 0.0;
    }

    const long long n = effective_n(a);
    const long long ntask = (n + PPW - 1) / PPW;
    for (long long task = (long long)blockIdx.x * WARPS + warp; task < ntask; task += (long long)gridDim.x * WARPS) {
        long long _t0 = 0;
        PHASE_START();
        const long long p = task * PPW + g;
        const bool valid = p < n;
        const long long pc = valid ? p : n - 1;
#if QTET_PREFETCH_NEXT
        {   // the next task's parameters and eigenvalues into L1 now: its first instructions wait for nothing then
            const long long pn_ = p + (long long)gridDim.x * WARPS * PPW;
            if (pn_ < n) {
                if (q == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.chis + pn_ * f));
                if (q < 3) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.ev + pn_ * d + q * 8));      // d doubles = up to 3 lines
            }
        }
#endif

        // ---- shifted diagonal of my point in shared memory: every lane builds its entries, all lanes sum all of them
        // (the mean must be the bit-identical one K_E subtracted), every lane shifts its entries
        __syncwarp();
        double* ap = abuf + g * D;
        {
            double chi[kMaxF];
#pragma unroll
            for (int k = 0; k < kMaxF; ++k) chi[k] = (k < f) ? a.chis[pc * f + k] : 0.0;
            double mine[RZ];
#pragma unroll
            for (int t = 0; t < RZ; ++t) {
                const int m = q + t * G;
                mine[t] = (m < d) ? diag_entry(slin, squad, chi, f, m) : 0.0;
                if (m < D) ap[m] = mine[t];
            }
            __syncwarp();
            double sum = 0.0;
#pragma unroll
            for (int m = 0; m < D; ++m) sum = __dadd_rn(sum, ap[m]);          // padded entries are 0
            const double mean = __ddiv_rn(sum, (double)d);
            __syncwarp();
#pragma unroll
            for (int t = 0; t < RZ; ++t) {
                const int m = q + t * G;
                if (m < D) ap[m] = (m < d) ? __dsub_rn(mine[t], mean) : 0.0;
            }
            __syncwarp();
        }

        // ---- my eigenvalues, gaps to the lower neighbour
        const double* evp = a.ev + pc * d;
        const double nrm = fmax(fabs(evp[0]), fabs(evp[d - 1]));
        double lam[RZ];
        unsigned gsbits = 0u;
        bool fb = false;
#pragma unroll
        for (int t = 0; t < RZ; ++t) {
            const int i = q + t * G;
            lam[t] = (i < d) ? evp[i] : 0.0;
            if (i >= 1 && i < d) {
                const double gap = lam[t] - evp[i - 1];
                if (!(gap > kGapFB * nrm)) fb = true;               // also catches NaN (K_E did not converge)
                else if (gap < kGapGS * nrm) gsbits |= 1u << i;
            }
        }

        // ---- two-sided recurrences of (T - lam) z = 0 for my RZ eigen-indices.  The columns live in shared memory (their
        // destination anyway: the transpose), only the running pair of each recurrence sits in registers.
        double* vp = vt + (size_t)g * D * LDV;
        double* colp[RZ];
#pragma unroll
        for (int t = 0; t < RZ; ++t) { const int i = q + t * G; colp[t] = vp + ((i < D) ? i : (D - 1)); }   // i >= D: nothing is stored
#if QTET_REC_FUSED
        // FUSED: the bottom-up (q) and the top-down (p) recurrence are independent, so they run in ONE unrolled loop -- six
        // dependency chains per lane instead of three -- with every p_k kept in a register (V's registers are still free here);
        // the join search and the final scaling are then plain loops over k without a recurrence (and without the third pass
        // that recomputed p).  Same operations in the same order per eigen-index as the rolled passes: bit-identical results.
        double invp[RZ], invq[RZ];
        int rj[RZ];
        {
            double preg[RZ][D];
            double qn[RZ], qc[RZ], pm[RZ], pc[RZ];
#pragma unroll
            for (int t = 0; t < RZ; ++t) {
                qn[t] = 0.0; qc[t] = (D - 1 == d - 1) ? 1.0 : 0.0;
                if (q + t * G < D) colp[t][(D - 1) * LDV] = qc[t];
                pm[t] = 0.0; pc[t] = 1.0; preg[t][0] = 1.0;
            }
#pragma unroll
            for (int s = 0; s < D - 1; ++s) {
                const int k = D - 1 - s;                       // q_{k-1} from the bottom, p_{s+1} from the top
                const double ak = ap[k], bk = sb[k], ibm = sib[k - 1];
                const double ak2 = ap[s], bm2 = (s >= 1) ? sb[s - 1] : 0.0, ib2 = sib[s];
#pragma unroll
                for (int t = 0; t < RZ; ++t) {
                    double val = fma(lam[t] - ak, qc[t], -(bk * qn[t])) * ibm;
                    if (!EXACT && k - 1 == d - 1) val = 1.0;
                    qn[t] = qc[t]; qc[t] = val;
                    if (q + t * G < D) colp[t][(k - 1) * LDV] = val;
                    const double pn = fma(lam[t] - ak2, pc[t], -(bm2 * pm[t])) * ib2;
                    pm[t] = pc[t]; pc[t] = pn;
                    preg[t][s + 1] = pn;
                }
            }
            __syncwarp();
            PHASE_MARK(11);
            // join row r = argmax |p_k q_k|
            double best[RZ], pr[RZ], qr[RZ];
#pragma unroll
            for (int t = 0; t < RZ; ++t) {
                const double q0 = colp[t][0];
                best[t] = fabs(q0); rj[t] = 0; pr[t] = 1.0; qr[t] = q0;
            }
#pragma unroll
            for (int k = 1; k < D; ++k) {
#pragma unroll
                for (int t = 0; t < RZ; ++t) {
                    const double qk = colp[t][k * LDV];
                    const double prod = fabs(preg[t][k] * qk);
                    if (prod > best[t]) { best[t] = prod; rj[t] = k; pr[t] = preg[t][k]; qr[t] = qk; }
                }
            }
#pragma unroll
            for (int t = 0; t < RZ; ++t) { invp[t] = fast_rcp(pr[t]); invq[t] = fast_rcp(qr[t]); }
            PHASE_MARK(12);
            // z_k = p_k / p_r above the join, q_k / q_r from the join down; 1 / |z| goes to rnbuf
            double ss[RZ];
#pragma unroll
            for (int t = 0; t < RZ; ++t) ss[t] = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) {
#pragma unroll
                for (int t = 0; t < RZ; ++t) {
                    const int i = q + t * G;
                    const bool above = k < rj[t];
                    const double qk = colp[t][k * LDV];
                    double x = (above ? preg[t][k] : qk) * (above ? invp[t] : invq[t]);
                    if (!EXACT && i >= d) x = (k == i) ? 1.0 : 0.0;             // padded eigen-index: identity column
                    if (i < D) colp[t][k * LDV] = x;
                    ss[t] = fma(x, x, ss[t]);
                }
            }
#pragma unroll
            for (int t = 0; t < RZ; ++t) {
                const int i = q + t * G;
                if (i < d && !(ss[t] >= 0.5 && ss[t] < 1e300)) { fb = true; ss[t] = 1.0; }   // z_r = 1: |z|^2 >= 1 unless NaN / overflow
                if (i < D) rnbuf[g * D + i] = (i < d) ? rsqrt_fast(ss[t]) : 1.0;
            }
        }
#elif QTET_REC_ROLLED
        // ROLLED loops: the three passes are ~40 instructions each instead of ~1700 unrolled ones.  The kernel is a 7000-
        // instruction straight line per task and `no_instruction` was its third stall reason (profiles/r2_*): what need not be
        // unrolled (nothing here indexes a register array) is not.  Every operand of step k+1 is loaded during step k.
        // from the bottom: q_{k-1} = ((lam - a_k) q_k - b_k q_{k+1}) / b_{k-1},  q_{d-1} = 1
        {
            double qn[RZ], qc[RZ];
#pragma unroll
            for (int t = 0; t < RZ; ++t) {
                qn[t] = 0.0; qc[t] = (D - 1 == d - 1) ? 1.0 : 0.0;
                if (q + t * G < D) colp[t][(D - 1) * LDV] = qc[t];
            }
            double ak = ap[D - 1], bk = sb[D - 1], ibm = sib[D - 2];
#pragma unroll 2
            for (int k = D - 1; k >= 1; --k) {
                const int kn = (k >= 2) ? k - 1 : 1;
                const double ak_n = ap[kn], bk_n = sb[kn], ibm_n = sib[kn - 1];
#pragma unroll
                for (int t = 0; t < RZ; ++t) {
                    double val = fma(lam[t] - ak, qc[t], -(bk * qn[t])) * ibm;
                    if (!EXACT && k - 1 == d - 1) val = 1.0;
                    qn[t] = qc[t]; qc[t] = val;
                    if (q + t * G < D) colp[t][(k - 1) * LDV] = val;
                }
                ak = ak_n; bk = bk_n; ibm = ibm_n;
            }
        }
        __syncwarp();
        PHASE_MARK(11);
        // from the top: p_{k+1} = ((lam - a_k) p_k - b_{k-1} p_{k-1}) / b_k,  p_0 = 1; join row r = argmax |p_k q_k|
        double invp[RZ], invq[RZ];
        int rj[RZ];
        {
            double pm[RZ], pcur[RZ], best[RZ], pr[RZ], qr[RZ], qk1[RZ];
#pragma unroll
            for (int t = 0; t < RZ; ++t) {
                const double q0 = colp[t][0];
                pm[t] = 0.0; pcur[t] = 1.0; best[t] = fabs(q0); rj[t] = 0; pr[t] = 1.0; qr[t] = q0;
                qk1[t] = colp[t][LDV];
            }
            double ak = ap[0], bm = 0.0, ib = sib[0];
#pragma unroll 2
            for (int k = 0; k < D - 1; ++k) {
                const int kn = (k + 1 < D - 1) ? k + 1 : k;
                const double ak_n = ap[kn], bm_n = sb[kn - 1 >= 0 ? kn - 1 : 0], ib_n = sib[kn];
                double qn2[RZ];
#pragma unroll
                for (int t = 0; t < RZ; ++t) qn2[t] = colp[t][(kn + 1) * LDV];
#pragma unroll
                for (int t = 0; t < RZ; ++t) {
                    const double pn = fma(lam[t] - ak, pcur[t], -(bm * pm[t])) * ib;
                    const double prod = fabs(pn * qk1[t]);
                    if (prod > best[t]) { best[t] = prod; rj[t] = k + 1; pr[t] = pn; qr[t] = qk1[t]; }
                    pm[t] = pcur[t]; pcur[t] = pn;
                    qk1[t] = qn2[t];
                }
                ak = ak_n; bm = bm_n; ib = ib_n;
            }
#pragma unroll
            for (int t = 0; t < RZ; ++t) { invp[t] = fast_rcp(pr[t]); invq[t] = fast_rcp(qr[t]); }
        }
        PHASE_MARK(12);
        // second pass from the top: z_k = p_k / p_r above the join, q_k / q_r from the join down; 1 / |z| goes to rnbuf
        // (the scale is applied when the rows are read back)
        {
            double pm[RZ], pcur[RZ], ss[RZ], qk[RZ];
#pragma unroll
            for (int t = 0; t < RZ; ++t) { pm[t] = 0.0; pcur[t] = 1.0; ss[t] = 0.0; qk[t] = colp[t][0]; }
            double ak = ap[0], bm = 0.0, ib = sib[0];
#pragma unroll 2
            for (int k = 0; k < D; ++k) {
                const int kn = (k + 1 < D) ? k + 1 : k;
                const double ak_n = ap[kn], bm_n = sb[kn - 1 >= 0 ? kn - 1 : 0], ib_n = sib[kn];
                double qn2[RZ];
#pragma unroll
                for (int t = 0; t < RZ; ++t) qn2[t] = colp[t][kn * LDV];
#pragma unroll
                for (int t = 0; t < RZ; ++t) {
                    const int i = q + t * G;
                    const bool above = k < rj[t];
                    double x = (above ? pcur[t] : qk[t]) * (above ? invp[t] : invq[t]);
                    if (!EXACT && i >= d) x = (k == i) ? 1.0 : 0.0;             // padded eigen-index: identity column
                    if (i < D) colp[t][k * LDV] = x;
                    ss[t] = fma(x, x, ss[t]);
                    const double pn = fma(lam[t] - ak, pcur[t], -(bm * pm[t])) * ib;     // (unused after the last row)
                    pm[t] = pcur[t]; pcur[t] = pn;
                    qk[t] = qn2[t];
                }
                ak = ak_n; bm = bm_n; ib = ib_n;
            }
#pragma unroll
            for (int t = 0; t < RZ; ++t) {
                const int i = q + t * G;
                if (i < d && !(ss[t] >= 0.5 && ss[t] < 1e300)) { fb = true; ss[t] = 1.0; }   // z_r = 1: |z|^2 >= 1 unless NaN / overflow
                if (i < D) rnbuf[g * D + i] = (i < d) ? rsqrt_fast(ss[t]) : 1.0;
            }
        }
#else
        // Diagonal and hopping constants in registers for the three passes: with two warps per scheduler nothing hides a
        // load that sits right in front of its use, and phase 0 holds no other large state.  The loads are volatile asm so
        // that they stay up here instead of being re-materialised next to each use.
        double av[D], bv[D], ibv[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const unsigned addr = (unsigned)__cvta_generic_to_shared(ap + k);
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(av[k]) : "r"(addr));
            bv[k] = a.boff[k]; ibv[k] = a.iboff[k];
        }
        PHASE_MARK(10);
        // from the bottom: q_{k-1} = ((lam - a_k) q_k - b_k q_{k+1}) / b_{k-1},  q_{d-1} = 1
        {
            double qn[RZ], qc[RZ];
#pragma unroll
            for (int t = 0; t < RZ; ++t) {
                qn[t] = 0.0; qc[t] = (D - 1 == d - 1) ? 1.0 : 0.0;
                if (q + t * G < D) colp[t][(D - 1) * LDV] = qc[t];
            }
#pragma unroll
            for (int k = D - 1; k >= 1; --k) {
                const double bk = bv[k], ibm = ibv[k - 1], ak = av[k];
#pragma unroll
                for (int t = 0; t < RZ; ++t) {
                    double val = fma(lam[t] - ak, qc[t], -(bk * qn[t])) * ibm;
                    if (k - 1 == d - 1) val = 1.0;
                    qn[t] = qc[t]; qc[t] = val;
                    if (q + t * G < D) colp[t][(k - 1) * LDV] = val;
                }
            }
        }
        __syncwarp();
        PHASE_MARK(11);
        // from the top: p_{k+1} = ((lam - a_k) p_k - b_{k-1} p_{k-1}) / b_k,  p_0 = 1; join row r = argmax |p_k q_k|
        double invp[RZ], invq[RZ];
        int rj[RZ];
        {
            double pm[RZ], pcur[RZ], best[RZ], pr[RZ], qr[RZ];
#pragma unroll
            for (int t = 0; t < RZ; ++t) {
                const double q0 = colp[t][0];
                pm[t] = 0.0; pcur[t] = 1.0; best[t] = fabs(q0); rj[t] = 0; pr[t] = 1.0; qr[t] = q0;
            }
#pragma unroll
            for (int k = 0; k < D - 1; ++k) {
                const double bm = (k >= 1) ? bv[k - 1] : 0.0, ib = ibv[k], ak = av[k];
#pragma unroll
                for (int t = 0; t < RZ; ++t) {
                    const double qk1 = colp[t][(k + 1) * LDV];
                    const double pn = fma(lam[t] - ak, pcur[t], -(bm * pm[t])) * ib;
                    const double prod = fabs(pn * qk1);
                    if (prod > best[t]) { best[t] = prod; rj[t] = k + 1; pr[t] = pn; qr[t] = qk1; }
                    pm[t] = pcur[t]; pcur[t] = pn;
                }
            }
#pragma unroll
            for (int t = 0; t < RZ; ++t) { invp[t] = fast_rcp(pr[t]); invq[t] = fast_rcp(qr[t]); }
        }
        PHASE_MARK(12);
        // second pass from the top: z_k = p_k / p_r above the join, q_k / q_r from the join down; 1 / |z| goes to rnbuf
        // (the scale is applied when the rows are read back)
        {
            double pm[RZ], pcur[RZ], ss[RZ];
#pragma unroll
            for (int t = 0; t < RZ; ++t) { pm[t] = 0.0; pcur[t] = 1.0; ss[t] = 0.0; }
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const double bm = (k >= 1) ? bv[k - 1] : 0.0, ib = ibv[k], ak = av[k];
#pragma unroll
                for (int t = 0; t < RZ; ++t) {
                    const int i = q + t * G;
                    const bool above = k < rj[t];
                    double x = (above ? pcur[t] : colp[t][k * LDV]) * (above ? invp[t] : invq[t]);
                    if (i >= d) x = (k == i) ? 1.0 : 0.0;                       // padded eigen-index: identity column
                    if (i < D) colp[t][k * LDV] = x;
                    ss[t] = fma(x, x, ss[t]);
                    if (k < D - 1) {
                        const double pn = fma(lam[t] - ak, pcur[t], -(bm * pm[t])) * ib;
                        pm[t] = pcur[t]; pcur[t] = pn;
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < RZ; ++t) {
                const int i = q + t * G;
                if (i < d && !(ss[t] >= 0.5 && ss[t] < 1e300)) { fb = true; ss[t] = 1.0; }   // z_r = 1: |z|^2 >= 1 unless NaN / overflow
                if (i < D) rnbuf[g * D + i] = (i < d) ? rsqrt_fast(ss[t]) : 1.0;
            }
        }
#endif
        PHASE_MARK(0);

        unsigned gm = gsbits;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) gm |= __shfl_xor_sync(0xffffffffu, gm, o);
        __syncwarp();
        if (__any_sync(0xffffffffu, gm != 0u)) {
            // modified Gram-Schmidt inside clusters of close eigenvalues (rare): column i against the columns j < i of
            // its cluster, the G lanes of the point sharing the rows
            // (the columns of a cluster are normalised in place first, their read-back scale becomes 1)
#pragma unroll 1
            for (int i = 0; i < d; ++i) {
                const bool in_cluster = ((gm >> i) & 1u) || (i + 1 < 32 && ((gm >> (i + 1)) & 1u));
                if (in_cluster) {
                    const double rn = rnbuf[g * D + i];
                    for (int k = q; k < D; k += G) vp[k * LDV + i] *= rn;
                }
                __syncwarp();
                if (in_cluster && q == 0) rnbuf[g * D + i] = 1.0;
            }
            __syncwarp();
#pragma unroll 1
            for (int i = 1; i < d; ++i) {
                const bool need = (gm >> i) & 1u;
                if (!__any_sync(0xffffffffu, need)) continue;
#pragma unroll 1
                for (int j = i - 1; j >= 0; --j) {
                    const unsigned run = (i - j >= 32) ? 0xffffffffu : ((1u << (i - j)) - 1u);
                    const bool inc = need && (((gm >> (j + 1)) & run) == run);
                    if (!__any_sync(0xffffffffu, inc)) break;
                    double dot = 0.0;
                    for (int k = q; k < D; k += G) dot = fma(vp[k * LDV + i], vp[k * LDV + j], dot);
                    dot = group_sum<G>(dot);
                    if (inc)
                        for (int k = q; k < D; k += G) vp[k * LDV + i] = fma(-dot, vp[k * LDV + j], vp[k * LDV + i]);
                    __syncwarp();
                }
                double s2 = 0.0;
                for (int k = q; k < D; k += G) s2 = fma(vp[k * LDV + i], vp[k * LDV + i], s2);
                s2 = group_sum<G>(s2);
                if (need) {
                    if (!(s2 > 0.25)) { fb = true; s2 = 1.0; }      // the two recurrences returned (nearly) the same vector
                    const double rn = 1.0 / sqrt(s2);
                    for (int k = q; k < D; k += G) vp[k * LDV + i] *= rn;
                }
                __syncwarp();
            }
        }
        // points the direct method cannot do go to the generic kernel
        {
            int fbi = fb ? 1 : 0;
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) fbi |= __shfl_xor_sync(0xffffffffu, fbi, o);
            if (fbi && q == 0 && valid) {
                const int slot = atomicAdd(a.fb_count, 1);
                a.fb_list[slot] = (int)p;
                atomicAdd(pv.stats + 0, 1ULL);
            }
        }
        // ---- put the eigen-indices that carry weight first.  c_i = V[0][i] decays fast away from the eigenvalues near the
        // initial state's energy, so the indices with |c_i| > kCsig form (almost) one window [lo, hi] of the ascending order: the
        // new order is the CYCLIC SHIFT by lo -- new position i' holds eigen-index (i' + lo) mod D -- and the first
        // nsig = hi - lo + 1 positions contain every index that matters (14.3 on average on the config-2 grid, against 13.5 for
        // an exact sort by |c_i|, which costs a 21-step ranking loop per point).
        double* evs = abuf + g * D;                                   // [D] eigenvalues in the new order (ap is dead)
        int lo = D, hi = -1;
#pragma unroll
        for (int t = 0; t < RZ; ++t) {
            const int i = q + t * G;
            if (i < d) {
                const double cabs = fabs(vp[i] * rnbuf[g * D + i]);
                if (cabs > kCsig) { lo = min(lo, i); hi = max(hi, i); }      // (NaN compares false: handled by the fallback)
            }
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (hi < lo) { lo = 0; hi = 0; }
        const int nsig = hi - lo + 1;
        __syncwarp();                                                 // everyone is done with ap (aliases evs)
#pragma unroll
        for (int t = 0; t < RZ; ++t) {
            const int i = q + t * G;
            if (i < D) { int pos = i - lo; if (pos < 0) pos += D; evs[pos] = lam[t]; }
        }
        int ncol = __reduce_max_sync(0xffffffffu, nsig);
#ifdef QTET_PHASE_TIMING
        if (lane == 0) atomicAdd(pv.stats + 3, (unsigned long long)ncol);      // debug builds: sum of the per-warp column counts
#endif
        __syncwarp();
        double v[R][D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            int src = i + lo;
            if (src >= D) src -= D;
            const double rn = rnbuf[g * D + src];
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int row = q * R + k;
                v[k][i] = (row < D) ? vp[row * LDV + src] * rn : 0.0;
            }
        }
        PHASE_MARK(1);
        evolve_and_grad<D, G, true>(a, v, wsite, lv, part, sbuf, zbuf, cbuf, p, valid, g, q, lane, dt, _t0, evs, ncol, nsig);
    }
}

struct KernelCfg { int per_sm = 0; size_t smem = 0; };

template <int D, int G, int WARPS, int MINB, bool EXACT>
int config_direct(KernelCfg& cfg) {
    using DL = DLayout<D, G>;
    cfg.smem = ((size_t)DL::CONSTS + (size_t)DL::PER_WARP * WARPS) * sizeof(double);
    auto kern = direct_kernel<D, G, WARPS, MINB, EXACT>;
    QTET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
    QTET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    QTET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cfg.per_sm, kern, WARPS * 32, cfg.smem));
    if (cfg.per_sm < 1) cfg.per_sm = 1;
    if (const char* env = std::getenv("QTET_KD_CTAS_PER_SM")) { const int v = std::atoi(env); if (v > 0 && v < cfg.per_sm) cfg.per_sm = v; }
    return QTET_OK;
}

template <int D, int G, int WARPS, int MINB, bool EXACT>
int launch_direct_kernel(const SplitArgs& a, const KernelCfg& cfg, int sms, cudaStream_t st) {
    using L = BLayout<D, G>;
    long long grid = (long long)sms * cfg.per_sm;
    const long long ntask = (a.n + L::PPW - 1) / L::PPW;
    const long long need = (ntask + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    direct_kernel<D, G, WARPS, MINB, EXACT><<<(unsigned)grid, WARPS * 32, cfg.smem, st>>>(a);
    count_launch();
    QTET_CUDA(cudaGetLastError());
    return QTET_OK;
}

int padded_dim_direct(int d) {
    const int sizes[] = {4, 8, 16, 21, 24, 32};
    for (int s : sizes) if (d <= s) return s;
    return -1;
}

}  // namespace

struct DirectWorkspace {
    int device = 0, sms = 0, D = 0;
    KernelCfg cfg_d;                 // K_D
    double boff[32] = {}, iboff[32] = {};   // host copies of b_k and 1 / b_k (kernel parameters of K_D)
    int per_sm_e = 0; size_t smem_e = 0;
    char* blob = nullptr;            // two chunk buffers (K_E of chunk c+1 overlaps K_D of chunk c)
    size_t blob_bytes = 0;
    int* fb_count = nullptr;         // [2]
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_e[2] = {nullptr, nullptr}, ev_d[2] = {nullptr, nullptr};
};

#ifdef QTET_PHASE_TIMING
extern "C" void qtet_debug_phase_cycles_direct(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)); }
}
#endif

bool direct_eligible(const ProblemView& pv) {
    return pv.tridiag && pv.offd_nonzero && pv.d >= 2 && pv.d <= 32 && pv.f <= kMaxF;
}

void direct_workspace_free(DirectWorkspace* ws) {
    if (!ws) return;
    if (ws->blob) cudaFree(ws->blob);
    if (ws->fb_count) cudaFree(ws->fb_count);
    if (ws->aux) cudaStreamDestroy(ws->aux);
    if (ws->ev_fork) cudaEventDestroy(ws->ev_fork);
    for (int i = 0; i < 2; ++i) {
        if (ws->ev_e[i]) cudaEventDestroy(ws->ev_e[i]);
        if (ws->ev_d[i]) cudaEventDestroy(ws->ev_d[i]);
    }
    delete ws;
}

int launch_generic_list(const ProblemView& pv, const double* chis, const int* list, const int* count, long long max_count,
                        int site, double* loss, double* grad, int32_t* tstar, double* trace, cudaStream_t st);

// eig_scalar_kernel is ONE function shared by handles of every Fock dimension: its dynamic shared-memory limit only ever grows
// (a handle of a smaller system created later must not lower it under the feet of an older, larger one).
static int raise_eig_smem_limit(size_t bytes) {
    static size_t have[64] = {};
    int dev = 0;
    QTET_CUDA(cudaGetDevice(&dev));
    dev &= 63;
    if (bytes > have[dev]) {
        QTET_CUDA(cudaFuncSetAttribute(eig_scalar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        have[dev] = bytes;
    }
    return QTET_OK;
}

static int direct_workspace_init(const ProblemView& pv, DirectWorkspace** wsp) {
    if (*wsp) return QTET_OK;
    DirectWorkspace* ws = new DirectWorkspace();
    *wsp = ws;
    QTET_CUDA(cudaGetDevice(&ws->device));
    QTET_CUDA(cudaDeviceGetAttribute(&ws->sms, cudaDevAttrMultiProcessorCount, ws->device));
    ws->D = padded_dim_direct(pv.d);
    {
        double off[32] = {};
        QTET_CUDA(cudaMemcpy(off, pv.offd, (size_t)(pv.d - 1) * sizeof(double), cudaMemcpyDeviceToHost));
        for (int k = 0; k + 1 < pv.d; ++k) { ws->boff[k] = off[k]; ws->iboff[k] = 1.0 / off[k]; }
    }
    int rc;
    switch (ws->D) {
        case 4: rc = config_direct<4, 4, 4, 4, false>(ws->cfg_d); break;
        case 8: rc = config_direct<8, 8, 4, 4, false>(ws->cfg_d); break;
        case 16: rc = config_direct<16, 8, 2, 6, false>(ws->cfg_d); break;
        case 21: rc = (pv.d == 21) ? config_direct<21, 8, 2, QTET_KD_MINB, true>(ws->cfg_d) : config_direct<21, 8, 2, 4, false>(ws->cfg_d); break;
        case 24: rc = config_direct<24, 8, 2, 4, false>(ws->cfg_d); break;
        default: rc = config_direct<32, 16, 2, 4, false>(ws->cfg_d); break;
    }
    if (rc) return rc;
    ws->smem_e = (size_t)kWarpsE * 2 * pv.d * 32 * sizeof(double);
    { const int rc_e = raise_eig_smem_limit(ws->smem_e); if (rc_e) return rc_e; }
    QTET_CUDA(cudaFuncSetAttribute(eig_scalar_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    QTET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ws->per_sm_e, eig_scalar_kernel, kWarpsE * 32, ws->smem_e));
    if (ws->per_sm_e < 1) ws->per_sm_e = 1;
    QTET_CUDA(cudaMalloc(&ws->fb_count, 2 * sizeof(int)));
    QTET_CUDA(cudaStreamCreateWithFlags(&ws->aux, cudaStreamNonBlocking));
    QTET_CUDA(cudaEventCreateWithFlags(&ws->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        QTET_CUDA(cudaEventCreateWithFlags(&ws->ev_e[i], cudaEventDisableTiming));
        QTET_CUDA(cudaEventCreateWithFlags(&ws->ev_d[i], cudaEventDisableTiming));
    }
    return QTET_OK;
}

// chunk size (points) the path would use for a batch of B points
static long long direct_chunk(long long B, bool single) {
    const long long all = (B + 31) / 32 * 32;
    if (single) return all;                                  // device-side batch size: one chunk (B <= 2^22 checked by the caller)
    long long chunk = ((B + 3) / 4 + 31) / 32 * 32;          // four chunks per call: K_E(c+1) overlaps K_D(c)
    if (chunk < 65536) chunk = 65536;
    if (chunk > (1LL << 22)) chunk = 1LL << 22;
    if (chunk > all) chunk = all;
    return chunk;
}

static size_t direct_buf_bytes(long long chunk, int d) {
    const size_t ev = (((size_t)chunk * d * 8 + 255) / 256) * 256;
    const size_t list = (((size_t)chunk * 4 + 255) / 256) * 256;
    return ev + list;
}

// Make the workspace large enough for batches of up to B points, so that later calls neither allocate nor synchronise.
int direct_reserve(const ProblemView& pv, DirectWorkspace** wsp, int64_t B, bool single) {
    int rc = direct_workspace_init(pv, wsp);
    if (rc) return rc;
    DirectWorkspace* ws = *wsp;
    const size_t total = 2 * direct_buf_bytes(direct_chunk(B, single), pv.d);
    if (total > ws->blob_bytes) {
        if (ws->blob) { cudaDeviceSynchronize(); cudaFree(ws->blob); }
        ws->blob = nullptr; ws->blob_bytes = 0;
        cudaError_t e = cudaMalloc(&ws->blob, total);
        if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(direct workspace) failed"); return QTET_ENOMEM; }
        ws->blob_bytes = total;
    }
    return QTET_OK;
}

int launch_direct(const ProblemView& pv, DirectWorkspace** wsp, const double* chis, int64_t B, const int* n_dev, int site,
                  double* loss, double* grad, int32_t* tstar, double* trace, cudaStream_t st, Profiler* prof,
                  const ChunkHook* hook) {
    if (B <= 0) return QTET_OK;
    if (!direct_eligible(pv)) { set_error("direct path not eligible for this problem"); return QTET_EUNSUPPORTED; }
    if (n_dev && B > (1LL << 22)) { set_error("device-side batch size needs B <= 2^22"); return QTET_EINVAL; }
    int rc = direct_reserve(pv, wsp, B, n_dev != nullptr);
    if (rc) return rc;
    DirectWorkspace* ws = *wsp;
    const int d = pv.d;
    const long long chunk = direct_chunk(B, n_dev != nullptr);
    const size_t buf_bytes = direct_buf_bytes(chunk, d);
    const size_t o_list = (((size_t)chunk * d * 8 + 255) / 256) * 256;

    SplitArgs a{};
    a.pv = pv; a.site = site; a.acceptor = (site == pv.f - 1) ? 1 : 0;
    for (int k = 0; k < 32; ++k) { a.boff[k] = ws->boff[k]; a.iboff[k] = ws->iboff[k]; }
    a.n_dev = n_dev;                      // a device-side count only makes sense for a single chunk (the optimiser's case)

    // One chunk with a device-side count (the optimiser's epoch): K_E and K_D run back to back on the caller's stream -- nothing
    // could overlap, and the two cross-stream hops would only add their scheduling latency to every epoch.
    const bool inline_e = (n_dev != nullptr);
    cudaStream_t se = inline_e ? st : ws->aux;
    if (!inline_e) {
        QTET_CUDA(cudaEventRecord(ws->ev_fork, st));
        QTET_CUDA(cudaStreamWaitEvent(ws->aux, ws->ev_fork, 0));
    }
    int c = 0;
    for (long long lo = 0; lo < B; lo += chunk, ++c) {
        const int buf = c & 1;
        char* base = ws->blob + (size_t)buf * buf_bytes;
        const long long n = (B - lo < chunk) ? (B - lo) : chunk;
        a.ev = (double*)base; a.fb_list = (int*)(base + o_list); a.fb_count = ws->fb_count + buf;
        a.chis = chis + lo * pv.f; a.n = n;
        a.loss = loss ? loss + lo : nullptr;
        a.grad = grad ? grad + lo * pv.f : nullptr;
        a.tstar = tstar ? tstar + lo : nullptr;
        a.trace = trace ? trace + lo * pv.T : nullptr;
        // K_E(c) on aux, after K_D(c-2) and its fallback released this buffer
        if (c >= 2 && !inline_e) QTET_CUDA(cudaStreamWaitEvent(ws->aux, ws->ev_d[buf], 0));
        if (hook && hook->before) { const int hrc = hook->before(hook->ctx, lo, n, se); if (hrc) return hrc; }
        long long gridE = (long long)ws->sms * ws->per_sm_e;
        const long long needE = ((n + 31) / 32 + kWarpsE - 1) / kWarpsE;
        if (gridE > needE) gridE = needE;
        if (prof) prof->begin(0, n, se);
        eig_scalar_kernel<<<(unsigned)gridE, kWarpsE * 32, ws->smem_e, se>>>(a);
        count_launch();
        if (prof) prof->finish(0, se);
        QTET_CUDA(cudaGetLastError());
        if (!inline_e) {
            QTET_CUDA(cudaEventRecord(ws->ev_e[buf], ws->aux));
            // K_D(c) on the caller's stream
            QTET_CUDA(cudaStreamWaitEvent(st, ws->ev_e[buf], 0));
        }
        if (prof) prof->begin(1, n, st);
        switch (ws->D) {
            case 4: rc = launch_direct_kernel<4, 4, 4, 4, false>(a, ws->cfg_d, ws->sms, st); break;
            case 8: rc = launch_direct_kernel<8, 8, 4, 4, false>(a, ws->cfg_d, ws->sms, st); break;
            case 16: rc = launch_direct_kernel<16, 8, 2, 6, false>(a, ws->cfg_d, ws->sms, st); break;
            case 21: rc = (d == 21) ? launch_direct_kernel<21, 8, 2, QTET_KD_MINB, true>(a, ws->cfg_d, ws->sms, st)
                                   : launch_direct_kernel<21, 8, 2, 4, false>(a, ws->cfg_d, ws->sms, st); break;
            case 24: rc = launch_direct_kernel<24, 8, 2, 4, false>(a, ws->cfg_d, ws->sms, st); break;
            default: rc = launch_direct_kernel<32, 16, 2, 4, false>(a, ws->cfg_d, ws->sms, st); break;
        }
        if (prof) prof->finish(1, st);
        if (rc) return rc;
        // numerically degenerate points (if any) are recomputed by the fused generic kernel
        rc = launch_generic_list(pv, a.chis, a.fb_list, a.fb_count, n, site, a.loss, a.grad, a.tstar, a.trace, st);
        if (rc) return rc;
        if (!inline_e) QTET_CUDA(cudaEventRecord(ws->ev_d[buf], st));
        if (hook && hook->after) { const int hrc = hook->after(hook->ctx, lo, n, st); if (hrc) return hrc; }
    }
    return QTET_OK;
}

}  // namespace qtet
