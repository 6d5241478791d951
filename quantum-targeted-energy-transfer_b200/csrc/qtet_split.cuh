// Shared pieces of the two tridiagonal d <= 32 paths (qtet_split.cu: QL rotation stream + replay; qtet_direct.cu:
// eigenvalues by QL, eigenvectors directly by two-sided recurrences): launch arguments, small device helpers, the per-warp
// shared-memory layout and the phases that follow once V sits in registers -- time evolution, loss, arg-max, and the
// analytic gradient (HamiltonianLoss.py:204-233, Optimizer.py:85-91).
#pragma once
#include <type_traits>

#include "qtet_common.cuh"

#ifndef QTET_EVO_GROUP
#define QTET_EVO_GROUP 3   // columns per exit test of the trimmed evolution loop
#endif
#ifndef QTET_EVO_TRIM
#define QTET_EVO_TRIM 1    // evolution over the significant columns only: 0 off, 1 exit test per group of four, 2 unrolled lengths
#endif

namespace qtet {
namespace splitk {

constexpr double kEps = 1.1102230246251565e-16;
constexpr int kMaxRounds = 128;            // header rows per batch
constexpr int kRing = 4;                   // rounds of rotation records in flight per warp (K_B)
constexpr int kTail = 3;                   // identity rows appended to every round (K_B sweeps planes in groups of 4)

struct SplitArgs {
    ProblemView pv;
    const double* chis;       // [n, f] (already offset to the chunk)
    long long n;              // points in this chunk
    int site, acceptor;
    int capit;                // stream rows (warp iterations) per batch
    double2* stream;          // [nb][capit][32]
    unsigned* hdr;            // [nb][kMaxRounds][32]   l (63 = idle) | mtop << 6 | lbot << 12 | it0 << 18
    int* nrounds;             // [nb]
    double* ev;               // [n, d]  eigenvalues (shifted by the mean of the diagonal)
    int* overflow_list;       // [n]
    int* overflow_count;      // [1]
    double* loss; double* grad; int* tstar; double* trace;   // chunk-offset outputs
    const int* n_dev = nullptr;   // optional: the number of points to process is min(n, *n_dev), read on the device
    int* fb_list = nullptr;       // direct path: points handed to the fused generic kernel (clustered eigenvalues) ...
    int* fb_count = nullptr;      // ... and their count
    double boff[32] = {};         // direct path: b_k = H[k+1,k] (0 for k >= d-1) and 1 / b_k, as kernel parameters: with the
    double iboff[32] = {};        // recurrences unrolled they are constant-bank operands of the FMAs, no load at all
};

// number of points this launch processes (host value, or the device-side count of a compacted list)
__device__ __forceinline__ long long effective_n(const SplitArgs& a) {
    if (a.n_dev == nullptr) return a.n;
    const long long nd = *a.n_dev;
    return nd < a.n ? nd : a.n;
}

__device__ __forceinline__ unsigned lowmask(int l) { return (l >= 32) ? 0xffffffffu : ((1u << l) - 1u); }
// 1/sqrt(x) for normal positive x: the hardware seed and one third-order step, i.e. what rsqrt() does on its fast path,
// without the range check / slow-path branch (the callers guard x themselves).
__device__ __forceinline__ double rsqrt_fast(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y * y, 1.0);
    return fma(y * e, fma(e, 0.375, 0.5), y);
}
constexpr double kTinyR2 = 1e-280;        // below this a rotation is treated as underflowed (as r2 == 0)

struct QlLane {
    double* arr;       // diag k at arr[64 k], off k at arr[64 k + 32]
    int d;
    __device__ __forceinline__ double& dg(int k) const { return arr[64 * k]; }
    __device__ __forceinline__ double& of(int k) const { return arr[64 * k + 32]; }
    __device__ __forceinline__ unsigned scan_small(double eps = kEps) const {
        unsigned small = 1u << (d - 1);
        for (int k = 0; k < d - 1; ++k) {
            const double dd = fabs(dg(k)) + fabs(dg(k + 1));
            if (fabs(of(k)) <= eps * dd) small |= 1u << k;
        }
        return small;
    }
    __device__ __forceinline__ double wilkinson_g(int l, int m) const {
        const double dl = dg(l), el = of(l);
        const double gg = (dg(l + 1) - dl) / (2.0 * el);
        const double rr = sqrt(fma(gg, gg, 1.0));
        return dg(m) - dl + el / (gg + copysign(rr, gg));
    }
};
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

#ifdef QTET_PHASE_TIMING
static __device__ unsigned long long g_phase_cycles[16];   // one copy per translation unit
#define PHASE_MARK(idx) do { const long long _t = clock64(); if (lane == 0) atomicAdd(&g_phase_cycles[idx], (unsigned long long)(_t - _t0)); _t0 = _t; } while (0)
#define PHASE_START() _t0 = clock64()
#else
#define PHASE_MARK(idx) do {} while (0)
#define PHASE_START() do {} while (0)
#endif

// Sum over the G lanes of a point (G a power of two, groups are aligned).
template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int D, int G>
struct BLayout {
    static constexpr int R = (D + G - 1) / G;          // rows of V per lane
    static constexpr int RZ = (D + G - 1) / G;         // eigen-indices owned per lane
    static constexpr int PPW = 32 / G;                 // points per warp
    static constexpr int NPL = D - 1;                  // planes
    static constexpr int ROWS_PER_COPY = 32 / PPW;     // stream rows moved by one warp-wide cp.async
    static constexpr int NCOPY = (NPL + kTail + ROWS_PER_COPY - 1) / ROWS_PER_COPY;
    static constexpr int SLOT_ROWS = 2 * NPL + kTail;  // [NPL identity guard rows][NPL + kTail data rows]
    static constexpr int DH = (D + 1) / 2;             // half of the eigen-indices (a-reduction in two halves)
    static constexpr int HH = D / 2;                   // cyclic offsets per eigen-index: the pair {i, (i + o) mod D}, o = 1 .. HH
    static constexpr int NP = D * HH;                  // ... covers every index pair once (even D: o = D/2 only for i < D/2)
    // shared memory per warp, in doubles
    static constexpr int RING = kRing * SLOT_ROWS * PPW * 2;           // double2 [kRing][SLOT_ROWS][PPW]
    static constexpr int HDR = kMaxRounds * PPW / 2;                   // unsigned [kMaxRounds][PPW]
    static constexpr int ZB = 4 * PPW * D * 2;                         // double2 [2 passes][2 samples][PPW][D]
    static constexpr int LV = 8 * PPW * D;                             // double [8][PPW][D]
    static constexpr int PART = PPW * G * DH * 2;                      // double2 [PPW][G][DH]
    static constexpr int SM_S = PPW * (D + NP);                        // double [PPW][D + NP]
    static constexpr int GRAD = ((LV + (PART > SM_S ? PART : SM_S)) + 1) & ~1;
    static constexpr int ROT = RING + HDR;
    static constexpr int PHASED = (ROT > GRAD ? ROT : GRAD);
    static constexpr int PER_WARP = ((PHASED + ZB + PPW * D) + 1) & ~1;    // + c vector [PPW][D]
};

// ------------------------------------------------------------------------------------------------
// Phases after the eigendecomposition, shared by apply_kernel (rotation replay) and direct_kernel
// (two-sided recurrences).  G lanes per point, lane q holds rows q*R .. q*R+R-1 of V in v[][].
//   lv [8][PPW][D], part [PPW][G][DH] / sbuf [PPW][D + NP] (aliased), zbuf [2][2][PPW][D], cbuf [PPW][D]
// ------------------------------------------------------------------------------------------------
//
// PERM (direct path): the eigen-indices arrive sorted by |c_i| descending, `evs` [D] holds the eigenvalues of this point
// in that order (shared memory) and only the first `ncol` (warp-uniform) columns carry weight in psi(t) = V (c . z(t)):
// c_i = V[0][i] is the overlap of eigenvector i with the initial state, and with |c_i| <= kCsig (1e-14, qtet_direct.cu) a column
// cannot move <n(t)> by more than 2e-14 N.  On the config-2 grid 11.7 of the 21 columns survive on average.  The gradient still needs
// every column (a = V^T w is not small for those i).  `nsig` is the point's own count: its c_i beyond that are set to
// exactly 0, so that a point's result is bit-identical whatever the other points of the warp need (batch invariance).
template <int D, int G, bool PERM = false>
__device__ __forceinline__ void evolve_and_grad(const SplitArgs& a, double (&v)[BLayout<D, G>::R][D],
                                                const double (&wsite)[BLayout<D, G>::R], double* lv, double2* part,
                                                double* sbuf, double2* zbuf, double* cbuf, long long p, bool valid,
                                                int g, int q, int lane, double dt, long long& _t0,
                                                const double* evs = nullptr, int ncol = D, int nsig = D) {
    using L = BLayout<D, G>;
    constexpr int R = L::R, RZ = L::RZ, PPW = L::PPW, NP = L::NP;
    const ProblemView& pv = a.pv;
    const int d = pv.d, f = pv.f, T = pv.T;
    (void)lane; (void)_t0;
    // ---- c = V[0,:] (row 0 lives in lane q = 0, k = 0), eigenvalues and unit phase steps of my indices
    __syncwarp();
    if (q == 0) {
#pragma unroll
        for (int i = 0; i < D; ++i) cbuf[g * D + i] = v[0][i];
    }
    __syncwarp();
    double ei[RZ], ci[RZ], wre[RZ], wim[RZ], zr[RZ], zi[RZ];
#pragma unroll
    for (int t = 0; t < RZ; ++t) {
        const int i = q + t * G;
        ci[t] = (i < D) ? cbuf[g * D + i] : 0.0;
        if (PERM && i >= nsig) ci[t] = 0.0;             // exact zeros: the result does not depend on the neighbours' ncol
        if (PERM) ei[t] = (i < D) ? evs[i] : 0.0;
        else ei[t] = (i < d && valid) ? a.ev[p * d + i] : 0.0;
        const double th = ei[t] * dt;
        const double lo = fma(ei[t], dt, -th);          // exact product tail
        double sn, cs;
        sincos(th, &sn, &cs);
        wre[t] = fma(-sn, lo, cs);                      // exp(-i (th + lo)) = (cs - i sn)(1 - i lo)
        wim[t] = -fma(cs, lo, sn);
        zr[t] = ci[t]; zi[t] = 0.0;
    }
    if (PERM) __syncwarp();                             // evs may alias zbuf
    PHASE_MARK(1);

    // ---- time evolution: n(t_k) = sum_rows n_site |psi_row(t_k)|^2, running arg-max / arg-min
    int kbest = -1;
    double vbest = 0.0;
    // two samples per pass: every V element is read once for four fmas (operand reuse), the pass overhead halves
    for (int k = 0; k < T; k += 2) {
        double2* zb0 = zbuf + ((((k >> 1) & 1) * 2 + 0) * PPW + g) * D;
        double2* zb1 = zbuf + ((((k >> 1) & 1) * 2 + 1) * PPW + g) * D;
#pragma unroll
        for (int t = 0; t < RZ; ++t) {
            const int i = q + t * G;
            const double nr_ = fma(zr[t], wre[t], -(zi[t] * wim[t]));
            const double ni_ = fma(zr[t], wim[t], zi[t] * wre[t]);
            if (i < D) { zb0[i] = make_double2(zr[t], zi[t]); zb1[i] = make_double2(nr_, ni_); }
            zr[t] = fma(nr_, wre[t], -(ni_ * wim[t]));
            zi[t] = fma(nr_, wim[t], ni_ * wre[t]);
        }
        __syncwarp();
        double pr0[R], pi0[R], pr1[R], pi1[R];
#pragma unroll
        for (int kk = 0; kk < R; ++kk) { pr0[kk] = pi0[kk] = pr1[kk] = pi1[kk] = 0.0; }
#if QTET_EVO_TRIM == 2
        // PERM: only the first ncol columns carry weight (the rest have z_i = 0).  The column loop comes in a few fully unrolled
        // lengths picked by ONE warp-uniform branch: an exit test inside the loop would keep the loads from running ahead
        // of the FMAs (measured: no gain from the shorter loop at all with a test per column, 13 % with one per four).
        auto columns = [&](auto NCc) {
            constexpr int NC = decltype(NCc)::value;
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const double2 za = zb0[i], zc = zb1[i];
#pragma unroll
                for (int kk = 0; kk < R; ++kk) {
                    pr0[kk] = fma(v[kk][i], za.x, pr0[kk]);
                    pi0[kk] = fma(v[kk][i], za.y, pi0[kk]);
                    pr1[kk] = fma(v[kk][i], zc.x, pr1[kk]);
                    pi1[kk] = fma(v[kk][i], zc.y, pi1[kk]);
                }
            }
        };
        if constexpr (PERM && D >= 16) {
            constexpr int C1 = (D * 2 + 4) / 5, C2 = (D * 3 + 4) / 5, C3 = (D * 4 + 4) / 5;      // ~40 %, 60 %, 80 % of D
            if (ncol <= C1) columns(std::integral_constant<int, C1>{});
            else if (ncol <= C2) columns(std::integral_constant<int, C2>{});
            else if (ncol <= C3) columns(std::integral_constant<int, C3>{});
            else columns(std::integral_constant<int, D>{});
        } else {
            columns(std::integral_constant<int, D>{});
        }
#else
        // columns in groups of four with ONE warp-uniform exit test per group (PERM: the columns from ncol on have |c_i| <= kCsig
        // and z_i = 0): inside a group the loads run ahead of the FMAs, a test per column would serialise them
#pragma unroll
        for (int i0 = 0; i0 < D; i0 += QTET_EVO_GROUP) {
            if (PERM && QTET_EVO_TRIM && i0 >= ncol) break;
            double2 za[QTET_EVO_GROUP], zc[QTET_EVO_GROUP];
#pragma unroll
            for (int ii = 0; ii < QTET_EVO_GROUP; ++ii) {
                if (i0 + ii < D) { za[ii] = zb0[i0 + ii]; zc[ii] = zb1[i0 + ii]; }
            }
#pragma unroll
            for (int ii = 0; ii < QTET_EVO_GROUP; ++ii) {
                const int i = i0 + ii;
                if (i < D) {
#pragma unroll
                    for (int kk = 0; kk < R; ++kk) {
                        pr0[kk] = fma(v[kk][i], za[ii].x, pr0[kk]);
                        pi0[kk] = fma(v[kk][i], za[ii].y, pi0[kk]);
                        pr1[kk] = fma(v[kk][i], zc[ii].x, pr1[kk]);
                        pi1[kk] = fma(v[kk][i], zc[ii].y, pi1[kk]);
                    }
                }
            }
        }
#endif
        double part0 = 0.0, part1 = 0.0;
#pragma unroll
        for (int kk = 0; kk < R; ++kk) {
            part0 = fma(wsite[kk], fma(pr0[kk], pr0[kk], pi0[kk] * pi0[kk]), part0);
            part1 = fma(wsite[kk], fma(pr1[kk], pr1[kk], pi1[kk] * pi1[kk]), part1);
        }
        {
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                part0 += __shfl_xor_sync(0xffffffffu, part0, o);
                part1 += __shfl_xor_sync(0xffffffffu, part1, o);
            }
            if (a.trace && valid) {
                if (q == (k % G)) a.trace[p * T + k] = part0;
                if (k + 1 < T && q == ((k + 1) % G)) a.trace[p * T + k + 1] = part1;
            }
            const double key0 = a.acceptor ? part0 : -part0, key1 = a.acceptor ? part1 : -part1;
            if (kbest < 0 || key0 > vbest) { vbest = key0; kbest = k; }        // first extremum wins ties
            if (k + 1 < T && key1 > vbest) { vbest = key1; kbest = k + 1; }
        }
    }
    if (q == 0 && valid) {
        if (a.loss) a.loss[p] = a.acceptor ? ((double)pv.N - vbest) : -vbest;
        if (a.tstar) a.tstar[p] = kbest;
    }
    PHASE_MARK(2);
    if (a.grad == nullptr) { __syncwarp(); return; }

    // ---- gradient at t* = t_kbest: f_i = w_i^kbest by binary powering (my eigen-indices)
    const double ts = pv.tgrid[kbest];
    double fr[RZ], fi[RZ];
#pragma unroll
    for (int t = 0; t < RZ; ++t) {
        double xr = 1.0, xi = 0.0, br = wre[t], bi = wim[t];
        for (int e = kbest; e > 0; e >>= 1) {
            if (e & 1) { const double tt = fma(xr, br, -(xi * bi)); xi = fma(xr, bi, xi * br); xr = tt; }
            const double t2 = fma(br, br, -(bi * bi)); bi = 2.0 * br * bi; br = t2;
        }
        fr[t] = xr; fi[t] = xi;
    }
    PHASE_MARK(6);
    __syncwarp();
    {
        double2* zb = zbuf + g * D;
#pragma unroll
        for (int t = 0; t < RZ; ++t) {
            const int i = q + t * G;
            if (i < D) zb[i] = make_double2(ci[t] * fr[t], ci[t] * fi[t]);
        }
    }
    __syncwarp();
    double wcr[R], wci[R];                               // n_site(row) conj(psi_row(t*))
    {
        const double2* zb = zbuf + g * D;
        double pr[R], pi[R];
#pragma unroll
        for (int kk = 0; kk < R; ++kk) { pr[kk] = 0.0; pi[kk] = 0.0; }
#pragma unroll
        for (int i0 = 0; i0 < D; i0 += 4) {
            if (PERM && i0 >= ncol) break;
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                const int i = i0 + ii;
                if (i < D) {
                    const double2 z = zb[i];
#pragma unroll
                    for (int kk = 0; kk < R; ++kk) { pr[kk] = fma(v[kk][i], z.x, pr[kk]); pi[kk] = fma(v[kk][i], z.y, pi[kk]); }
                }
            }
        }
#pragma unroll
        for (int kk = 0; kk < R; ++kk) { wcr[kk] = wsite[kk] * pr[kk]; wci[kk] = -wsite[kk] * pi[kk]; }
    }
    PHASE_MARK(7);
    // a_i = sum_rows V[row][i] wc_row : per-lane partial sums over my rows, combined through shared
    // memory in two halves of i; the owner of eigen-index i adds the G partials.
    double ar[RZ], ai[RZ];
#pragma unroll
    for (int t = 0; t < RZ; ++t) { ar[t] = 0.0; ai[t] = 0.0; }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        __syncwarp();
#pragma unroll
        for (int ii = 0; ii < L::DH; ++ii) {
            const int i = half * L::DH + ii;
            if (i < D) {
                double sr = 0.0, si = 0.0;
#pragma unroll
                for (int kk = 0; kk < R; ++kk) { sr = fma(v[kk][i], wcr[kk], sr); si = fma(v[kk][i], wci[kk], si); }
                part[(g * G + q) * L::DH + ii] = make_double2(sr, si);
            }
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < RZ; ++t) {
            const int i = q + t * G;
            const int ii = i - half * L::DH;
            if (i < D && ii >= 0 && ii < L::DH) {
#pragma unroll
                for (int qq = 0; qq < G; ++qq) { const double2 x = part[(g * G + qq) * L::DH + ii]; ar[t] += x.x; ai[t] += x.y; }
            }
        }
    }
    __syncwarp();
    PHASE_MARK(8);
    // per-eigen-index vectors of my point: e, c, f, a, b = a f ; diagonal of the kernel K_ii = c_i t* Im(b_i)
    double* lvp = lv + g * D;
    double* sp = sbuf + g * (D + NP);
#pragma unroll
    for (int t = 0; t < RZ; ++t) {
        const int i = q + t * G;
        if (i < D) {
            const double bre = fma(ar[t], fr[t], -(ai[t] * fi[t])), bim = fma(ar[t], fi[t], ai[t] * fr[t]);
            lvp[0 * PPW * D + i] = ei[t]; lvp[1 * PPW * D + i] = ci[t];
            lvp[2 * PPW * D + i] = fr[t]; lvp[3 * PPW * D + i] = fi[t];
            lvp[4 * PPW * D + i] = ar[t]; lvp[5 * PPW * D + i] = ai[t];
            lvp[6 * PPW * D + i] = bre; lvp[7 * PPW * D + i] = bim;
            sp[i] = ci[t] * ts * bim;
        }
    }
    __syncwarp();
    PHASE_MARK(9);
    // S_ij = K_ij + K_ji.  Lane q pairs each of its own eigen-indices i (whose e, c, f, a, b it holds in registers)
    // with j = (i + o) mod D, o = 1 .. D/2: every index pair exactly once, the partner's values are read from
    // consecutive addresses by the lanes of a point (no table look-up, no gather), and S lands in the order
    // [i][o] the row sums below walk through with static indices.
    {
        constexpr int HH = L::HH, U = 5;
        bool near_any = false;
#pragma unroll
        for (int t = 0; t < RZ; ++t) {
            const int i = q + t * G;
            if (i < D) {
                const double e_i = ei[t], c_i = ci[t], fr_i = fr[t], fi_i = fi[t], ar_i = ar[t], ai_i = ai[t];
                const double br_i = fma(ar_i, fr_i, -(ai_i * fi_i));
                const int omax = ((D & 1) || i < D / 2) ? HH : HH - 1;
#pragma unroll 1
                for (int o0 = 1; o0 <= omax; o0 += U) {
                    double dl[U], kd[U];
                    int jj[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        int j = i + ((o0 + u <= omax) ? (o0 + u) : omax);
                        if (j >= D) j -= D;
                        jj[u] = j;
                        dl[u] = e_i - lvp[j];
                        // K_ij = c_j (br_i - ar_i fr_j + ai_i fi_j) ; K_ji = c_i (br_j - ar_j fr_i + ai_j fi_i)
                        const double kij = lvp[PPW * D + j] * (br_i - ar_i * lvp[2 * PPW * D + j] + ai_i * lvp[3 * PPW * D + j]);
                        const double kji = c_i * (lvp[6 * PPW * D + j] - lvp[4 * PPW * D + j] * fr_i + lvp[5 * PPW * D + j] * fi_i);
                        kd[u] = kij - kji;
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) kd[u] *= fast_rcp(dl[u]);
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (o0 + u <= omax) {
                            near_any |= fabs(dl[u] * ts) < 1e-2;
                            sp[D + i * HH + (o0 + u - 1)] = kd[u];
                        }
                    }
                }
            }
        }
        // near-degenerate pairs (|dl t*| < 1e-2, rare): Phi_ij = ts f_j (exp(-i x) - 1)/x ; S = Re[(a_i c_j + a_j c_i) Phi_ij].
        // Kept out of the hot loop above (a rolled fix-up pass over the lane's pairs, entered only if some lane of the warp saw one):
        // five inlined copies of this block inside the unrolled pair loop were 170 instructions of mostly-skipped code.
        if (__any_sync(0xffffffffu, near_any)) {
#pragma unroll 1
            for (int t = 0; t < RZ; ++t) {
                const int i = q + t * G;
                if (i >= D) continue;
                const double e_i = lvp[i], c_i = lvp[PPW * D + i], ar_i = lvp[4 * PPW * D + i], ai_i = lvp[5 * PPW * D + i];
                const int omax = ((D & 1) || i < D / 2) ? HH : HH - 1;
#pragma unroll 1
                for (int o = 1; o <= omax; ++o) {
                    int j = i + o;
                    if (j >= D) j -= D;
                    const double x = (e_i - lvp[j]) * ts;
                    if (fabs(x) < 1e-2) {
                        double gr, gi;
                        expm1_over_x(x, gr, gi);
                        const double fjr = lvp[2 * PPW * D + j], fji = lvp[3 * PPW * D + j];
                        const double phr = ts * (fjr * gr - fji * gi), phi = ts * (fjr * gi + fji * gr);
                        const double c_j = lvp[PPW * D + j];
                        const double mr = ar_i * c_j + lvp[4 * PPW * D + j] * c_i;
                        const double mi = ai_i * c_j + lvp[5 * PPW * D + j] * c_i;
                        sp[D + i * HH + (o - 1)] = mr * phr - mi * phi;
                    }
                }
            }
        }
    }
    __syncwarp();
    PHASE_MARK(3);
    double Gm[R];
#pragma unroll
    for (int kk = 0; kk < R; ++kk) Gm[kk] = 0.0;
    {
        constexpr int HH = L::HH;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const double kd = sp[i];
            double acc[R];
#pragma unroll
            for (int kk = 0; kk < R; ++kk) acc[kk] = kd * v[kk][i];
#pragma unroll
            for (int o = 1; o <= HH; ++o) {
                if ((D & 1) || o < HH || i < D / 2) {
                    const int j = (i + o) % D;
                    const double sv = sp[D + i * HH + (o - 1)];
#pragma unroll
                    for (int kk = 0; kk < R; ++kk) acc[kk] = fma(sv, v[kk][j], acc[kk]);
                }
            }
#pragma unroll
            for (int kk = 0; kk < R; ++kk) Gm[kk] = fma(v[kk][i], acc[kk], Gm[kk]);
        }
    }
    PHASE_MARK(4);
    for (int k = 0; k < f; ++k) {
        double s = 0.0;
#pragma unroll
        for (int kk = 0; kk < R; ++kk) {
            const int row = q * R + kk;
            const double qk = (row < d) ? pv.quad[row * f + k] : 0.0;
            s = fma(qk, 2.0 * Gm[kk], s);
        }
        s = group_sum<G>(s);
        if (q == 0 && valid) a.grad[p * f + k] = a.acceptor ? -s : s;
    }
    __syncwarp();
    PHASE_MARK(5);
}

}  // namespace splitk
}  // namespace qtet
