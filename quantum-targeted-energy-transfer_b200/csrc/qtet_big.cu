// Split path for LARGE Fock dimensions (32 < d <= 115: dimers up to N = 114, trimer N = 10 with d = 66, ...).
//
//   [dense H only]  generic_kernel mode 1   CTA per point: H build + Householder -> tridiagonal + Q
//   ql_scalar_big_kernel                    THREAD per point: two-pass QL with perfect shifts on the tridiagonal,
//                                           rounds recorded as coalesced rows (same scheme as qtet_split.cu, with
//                                           128-bit convergence masks and the tridiagonal in shared memory)
//   generic_kernel mode 2                   CTA per point, thread = row of V in shared memory: replays the rounds
//                                           on V = I (or Q), then evolution + gradient
//
// The fused generic kernel (mode 0) spends >90 % of its time in the serial QL recurrence, recomputed by every
// warp of every CTA; here that recurrence runs once per point in one thread, 32 points per warp.
#include <cstdlib>

#include "qtet_big.cuh"

namespace qtet {

namespace {

constexpr double kEps = 1.1102230246251565e-16;
// Pass 1 only supplies shifts: an eigenvalue taken when its off-diagonal is below 2^-33 relative is still accurate to
// ~e^2/gap (second order), and the looser test saves a quarter of pass 1's rotations (tools/ql_count.py).
constexpr double kEpsPass1 = kEps * 1048576.0;
constexpr int kWarpsBig = 2;

struct Mask128 {
    unsigned long long lo = 0ull, hi = 0ull;
    __device__ __forceinline__ void set(int k) { if (k < 64) lo |= 1ull << k; else hi |= 1ull << (k - 64); }
    __device__ __forceinline__ void clr(int k) { if (k < 64) lo &= ~(1ull << k); else hi &= ~(1ull << (k - 64)); }
    __device__ __forceinline__ void put(int k, bool v) { if (v) set(k); else clr(k); }
    // first set bit at index >= l, or -1
    __device__ __forceinline__ int first_set_ge(int l) const {
        if (l < 64) {
            const unsigned long long x = lo >> l;
            if (x) return l + __ffsll((long long)x) - 1;
            if (hi) return 64 + __ffsll((long long)hi) - 1;
            return -1;
        }
        if (l >= 128) return -1;
        const unsigned long long x = hi >> (l - 64);
        return x ? (l + __ffsll((long long)x) - 1) : -1;
    }
    // first CLEAR bit at index in [l, limit), or -1
    __device__ __forceinline__ int first_clear_ge(int l, int limit) const {
        Mask128 inv;
        inv.lo = ~lo; inv.hi = ~hi;
        const int k = inv.first_set_ge(l);
        return (k >= 0 && k < limit) ? k : -1;
    }
};

// 1/sqrt(x) for normal positive x: hardware seed + one third-order step (rsqrt()'s fast path without its range check)
__device__ __forceinline__ double rsqrt_fast(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y * y, 1.0);
    return fma(y * e, fma(e, 0.375, 0.5), y);
}
constexpr double kTinyR2 = 1e-280;        // below this a rotation is treated as underflowed (as r2 == 0)

// A single bit walking down a Mask128 (the flag of the off-diagonal a sweep has just finalised)
struct WalkBit {
    unsigned long long lo = 0ull, hi = 0ull;
    __device__ __forceinline__ void at(int k) { lo = (k < 64) ? (1ull << k) : 0ull; hi = (k >= 64) ? (1ull << (k - 64)) : 0ull; }
    __device__ __forceinline__ void down() { lo = (lo >> 1) | (hi << 63); hi >>= 1; }
};
__device__ __forceinline__ void put_bit(Mask128& m, const WalkBit& b, bool v) {
    m.lo = v ? (m.lo | b.lo) : (m.lo & ~b.lo);
    m.hi = v ? (m.hi | b.hi) : (m.hi & ~b.hi);
}

struct BigQlArgs {
    ProblemView pv;
    const double* chis;       // [n, f]
    long long n;
    BigStream bs;
    int use_tri;              // 1: tridiagonal comes from bs.tri_d / bs.tri_e (dense H), 0: built from chis
    int* overflow_list;
    int* overflow_count;
};

// Storage of the lane's tridiagonal: shared memory [2k + {0,1}][lane] (bank = lane), or -- LOCAL -- per-thread arrays
// in local memory (L1/L2-cached, lane-interleaved by the hardware).  The shared layout caps the kernel at
// 228 KB / (2 d 256 B) warps per SM (4 at d = 101), which leaves the serial rsqrt/FMA chain of a sweep exposed; the
// local layout trades some L2 traffic for 4x the resident warps.
template <bool LOCAL>
struct QlLaneBig {
    double* arr;       // shared: diag k at arr[64 k], off k at arr[64 k + 32]
    double ld[LOCAL ? 128 : 1];
    double lo[LOCAL ? 128 : 1];
    int d;
    __device__ __forceinline__ double& dg(int k) { if constexpr (LOCAL) return ld[k]; else return arr[64 * k]; }
    __device__ __forceinline__ double& of(int k) { if constexpr (LOCAL) return lo[k]; else return arr[64 * k + 32]; }
    __device__ __forceinline__ Mask128 scan_small(double eps = kEps) {
        Mask128 m;
        m.set(d - 1);
        for (int k = 0; k < d - 1; ++k) {
            const double dd = fabs(dg(k)) + fabs(dg(k + 1));
            if (fabs(of(k)) <= eps * dd) m.set(k);
        }
        return m;
    }
    __device__ __forceinline__ double wilkinson_g(int l, int m) {
        const double dl = dg(l), el = of(l);
        const double gg = (dg(l + 1) - dl) / (2.0 * el);
        const double rr = sqrt(fma(gg, gg, 1.0));
        return dg(m) - dl + el / (gg + copysign(rr, gg));
    }
};

template <bool LOCAL>
__global__ void __launch_bounds__(kWarpsBig * 32) ql_scalar_big_kernel(BigQlArgs a) {
    extern __shared__ double sm[];
    const ProblemView& pv = a.pv;
    const int d = pv.d, f = pv.f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    QlLaneBig<LOCAL> t;
    t.arr = LOCAL ? nullptr : sm + (size_t)warp * 2 * d * 32 + lane;
    t.d = d;
    const long long nbatch = (a.n + 31) / 32;
    for (long long batch = (long long)blockIdx.x * kWarpsBig + warp; batch < nbatch; batch += (long long)gridDim.x * kWarpsBig) {
        const long long p = batch * 32 + lane;
        const bool valid = p < a.n;
        const long long pc = valid ? p : a.n - 1;
        double hchi[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) hchi[k] = (k < f && !a.use_tri) ? 0.5 * a.chis[pc * f + k] : 0.0;
        double lam[128];                                   // eigenvalue pass 1 left at each index (local memory)
        double mean = 0.0;
        for (int pass = 0; pass < 2; ++pass) {
            double sum = 0.0;
            for (int m = 0; m < d; ++m) {
                double acc, off;
                if (a.use_tri) {
                    acc = a.bs.tri_d[pc * d + m];
                    off = (m < d - 1) ? a.bs.tri_e[pc * d + m] : 0.0;
                } else {
                    acc = 0.0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (k < f) {
                            const double n = pv.occ[m * f + k];
                            acc = acc + (pv.omegas[k] * n + hchi[k] * (n * n));
                        }
                    }
                    off = (m < d - 1) ? pv.offd[m] : 0.0;
                }
                t.dg(m) = acc;
                t.of(m) = off;
                sum += acc;
            }
            if (pass == 0) mean = sum / (double)d;
            for (int m = 0; m < d; ++m) t.dg(m) -= mean;
            if (pass == 1) break;
            // ---- pass 1: eigenvalues only (plain implicit QL, Wilkinson shifts)
            Mask128 small = t.scan_small(kEpsPass1);
            int l = 0, iter = 0;
            while (true) {
                const int lnew = small.first_clear_ge(l, d - 1);
                if (lnew < 0) break;
                if (lnew != l) iter = 0;
                l = lnew;
                const int m = small.first_set_ge(l);
                if (++iter > 60) break;                    // give up quietly: pass 2 does not rely on pass 1
                double g = t.wilkinson_g(l, m);
                double s = 1.0, c = 1.0, pp = 0.0, dnext = 0.0;
                bool broke = false;
                // dg(i+1), dg(i), of(i) of the coming plane are carried in registers: every diagonal entry is loaded once
                // (two consecutive planes read it) and one plane ahead of its use, off the dependency chain
                double d_hi = t.dg(m), d_i = t.dg(m - 1), e_i = t.of(m - 1);
                WalkBit fb; fb.at(m);
                for (int i = m - 1; i >= l; --i) {
                    const int in = (i > l) ? i - 1 : i;
                    const double d_n = t.dg(in), e_n = t.of(in);
                    const double fo = s * e_i, bo = c * e_i;
                    const double r2 = fma(fo, fo, g * g);
                    if (r2 < kTinyR2) { t.of(i + 1) = 0.0; t.dg(i + 1) = d_hi - pp; t.of(m) = 0.0; broke = true; break; }
                    const double rinv = rsqrt_fast(r2);
                    const double r = r2 * rinv;
                    s = fo * rinv; c = g * rinv;
                    g = d_hi - pp;
                    const double rr = fma(d_i - g, s, 2.0 * c * bo);
                    pp = s * rr;
                    const double dnew = g + pp;
                    t.dg(i + 1) = dnew;
                    g = fma(c, rr, -bo);
                    t.of(i + 1) = r;                           // of(m) is reset below
                    put_bit(small, fb, r <= kEpsPass1 * (fabs(dnew) + fabs(dnext)));
                    fb.down();
                    dnext = dnew; d_hi = d_i; d_i = d_n; e_i = e_n;
                }
                if (broke) { small = t.scan_small(kEpsPass1); continue; }
                const double dlnew = d_hi - pp;                // d_hi = dg(l) after the last plane
                t.dg(l) = dlnew; t.of(l) = g; t.of(m) = 0.0;
                small.set(m);
                small.put(l, fabs(g) <= kEpsPass1 * (fabs(dlnew) + fabs(dnext)));
            }
            for (int m = 0; m < d; ++m) lam[m] = t.dg(m);
        }

        // ---- pass 2: recorded rounds, perfect shift on the first sweep at every index
        double2* rec = a.bs.stream + (size_t)batch * a.bs.capit * 32 + lane;
        unsigned* hdr = a.bs.hdr + (size_t)batch * a.bs.maxr * 32 + lane;
        int* it0s = a.bs.it0 + (size_t)batch * a.bs.maxr;
        Mask128 small = t.scan_small();
        int it = 0, round = 0, l = 0, iter = 0, tried = -1;
        bool overflow = false;
        while (true) {
            const int lnew0 = small.first_clear_ge(l, d - 1);
            const bool done = (lnew0 < 0);
            const int lnew = done ? d : lnew0;
            if (lnew != l) iter = 0;
            l = lnew;
            if (__all_sync(0xffffffffu, done)) break;
            const int m = done ? 0 : small.first_set_ge(l);
            ++iter;
            const int mtop = __reduce_max_sync(0xffffffffu, m);
            const int lbot = __reduce_min_sync(0xffffffffu, done ? 255 : l);
            const int nrows = mtop - lbot;
            if (__any_sync(0xffffffffu, !done && iter > 60) || round >= a.bs.maxr - 1 || it + nrows > a.bs.capit) {
                overflow = true;
                break;
            }
            // l (0xff = idle) | mtop << 8 (batch-uniform top of the round's rows) | lbot << 16 | my own block top m << 24
            hdr[(size_t)round * 32] = (unsigned)(done ? 0xff : l) | ((unsigned)mtop << 8) | ((unsigned)lbot << 16) | ((unsigned)m << 24);
            if (lane == 0) it0s[round] = it;
            double g = 0.0, s = 1.0, c = 1.0, pp = 0.0, dnext = 0.0;
            if (!done) {
                if (tried != l) { tried = l; g = t.dg(m) - lam[l]; }
                else g = t.wilkinson_g(l, m);
            }
            bool broke = false;
            // stream row it + (mtop-1-i) holds, for every lane, the rotation of plane i (identity if untouched)
            double2* row = rec + (size_t)it * 32;
            double d_hi = 0.0, d_i = 0.0, e_i = 0.0;
            WalkBit fb;
            if (!done) { d_hi = t.dg(m); d_i = t.dg(m - 1); e_i = t.of(m - 1); fb.at(m); }
            for (int i = mtop - 1; i >= lbot; --i, row += 32) {
                double2 out = make_double2(1.0, 0.0);
                if (!done && !broke && i < m && i >= l) {
                    const int in = (i > l) ? i - 1 : i;
                    const double d_n = t.dg(in), e_n = t.of(in);
                    const double fo = s * e_i, bo = c * e_i;
                    const double r2 = fma(fo, fo, g * g);
                    if (r2 < kTinyR2) {
                        t.of(i + 1) = 0.0; t.dg(i + 1) = d_hi - pp; t.of(m) = 0.0;
                        broke = true;
                    } else {
                        const double rinv = rsqrt_fast(r2);
                        const double r = r2 * rinv;
                        s = fo * rinv; c = g * rinv;
                        g = d_hi - pp;
                        const double rr = fma(d_i - g, s, 2.0 * c * bo);
                        pp = s * rr;
                        const double dnew = g + pp;
                        t.dg(i + 1) = dnew;
                        g = fma(c, rr, -bo);
                        t.of(i + 1) = r;                       // of(m) is reset below
                        put_bit(small, fb, r <= kEps * (fabs(dnew) + fabs(dnext)));
                        fb.down();
                        dnext = dnew; d_hi = d_i; d_i = d_n; e_i = e_n;
                        out = make_double2(c, s);
                    }
                }
                *row = out;                                // all 32 lanes: one coalesced 512 B row
            }
            it += nrows;
            if (!done) {
                if (broke) small = t.scan_small();
                else {
                    const double dlnew = d_hi - pp;            // d_hi = dg(l) after the last plane
                    t.dg(l) = dlnew; t.of(l) = g; t.of(m) = 0.0;
                    small.set(m);
                    small.put(l, fabs(g) <= kEps * (fabs(dlnew) + fabs(dnext)));
                }
            }
            ++round;
        }
        if (lane == 0) a.bs.nrounds[batch] = overflow ? 0 : round;
        if (valid) {
            if (overflow) {
                const int slot = atomicAdd(a.overflow_count, 1);
                a.overflow_list[slot] = (int)p;
                atomicAdd(a.pv.stats + 1, 1ULL);
            }
            for (int m = 0; m < d; ++m) a.bs.ev[p * d + m] = t.dg(m);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Direct path: eigenvalues only (one sweep family, full precision), sorted ascending, plus the shifted diagonal the
// eigenvector kernel starts its recurrences from.  No rotation stream: 16 d bytes per point leave this kernel instead of
// ~(d^2 + 8 d) * 16.
// ------------------------------------------------------------------------------------------------
#ifndef QTET_KE_PWK
#define QTET_KE_PWK 0     // 1: square-root-free rational QL (dsterf recurrence); 0: rotation-based implicit QL
#endif
#if QTET_KE_PWK
__device__ __forceinline__ double fast_rcp_big(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}
// off-diagonal negligible, tested on its square:  |e_k| <= eps (|d_k| + |d_k+1|)
__device__ __forceinline__ bool tiny_sq(double e2v, double dk, double dk1) {
    const double tst = kEps * (fabs(dk) + fabs(dk1));
    return e2v <= tst * tst;
}

// Square-root-free rational QL (Pal-Walker-Kahan / LAPACK dsterf) on the squared off-diagonals, see qtet_direct.cu: the kernel
// is a latency chain per point, and this recurrence's chain is about half as long as the rotation-based one.
template <bool LOCAL>
__global__ void __launch_bounds__(kWarpsBig * 32) eig_scalar_big_kernel(BigQlArgs a) {
    extern __shared__ double sm[];
    const ProblemView& pv = a.pv;
    const int d = pv.d, f = pv.f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    QlLaneBig<LOCAL> t;                                   // of(k) holds the SQUARED off-diagonal here
    t.arr = LOCAL ? nullptr : sm + (size_t)warp * 2 * d * 32 + lane;
    t.d = d;
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.overflow_count = 0;      // the eigenvector kernel (stream-ordered after us) appends
    const long long nbatch = (a.n + 31) / 32;
    for (long long batch = (long long)blockIdx.x * kWarpsBig + warp; batch < nbatch; batch += (long long)gridDim.x * kWarpsBig) {
        const long long p = batch * 32 + lane;
        const bool valid = p < a.n;
        const long long pc = valid ? p : a.n - 1;
        double hchi[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) hchi[k] = (k < f && !a.use_tri) ? 0.5 * a.chis[pc * f + k] : 0.0;
        double sum = 0.0;
        for (int m = 0; m < d; ++m) {
            double acc, off;
            if (a.use_tri) {
                acc = a.bs.tri_d[pc * d + m];
                off = (m < d - 1) ? a.bs.tri_e[pc * d + m] : 0.0;
            } else {
                acc = 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (k < f) {
                        const double n = pv.occ[m * f + k];
                        acc = acc + (pv.omegas[k] * n + hchi[k] * (n * n));
                    }
                }
                off = (m < d - 1) ? pv.offd[m] : 0.0;
            }
            t.dg(m) = acc;
            t.of(m) = off * off;
            sum += acc;
        }
        const double mean = sum / (double)d;
        for (int m = 0; m < d; ++m) {
            const double x = t.dg(m) - mean;
            t.dg(m) = x;
            if (valid) a.bs.adiag[p * d + m] = x;         // what the recurrences of the eigenvector kernel run on
        }
        Mask128 small;
        small.set(d - 1);
        for (int k = 0; k < d - 1; ++k)
            if (tiny_sq(t.of(k), t.dg(k), t.dg(k + 1))) small.set(k);
        int l = 0, iter = 0;
        bool failed = false;
        while (true) {
            const int lnew = small.first_clear_ge(l, d - 1);
            if (lnew < 0) break;
            if (lnew != l) iter = 0;
            l = lnew;
            const int m = small.first_set_ge(l);
            if (++iter > 60) { failed = true; break; }
            double sigma;
            {
                const double pl = t.dg(l), rte = sqrt(t.of(l));
                const double sg = (t.dg(l + 1) - pl) / (2.0 * rte);
                const double rr = sqrt(fma(sg, sg, 1.0));
                sigma = pl - rte / (sg + copysign(rr, sg));
            }
            double c = 1.0, s = 0.0, gamma = t.dg(m) - sigma, pp = gamma * gamma, dnext = 0.0;
            double alpha = t.dg(m - 1), bb = t.of(m - 1);   // loaded a plane ahead of their use
            WalkBit fb; fb.at(m);
            for (int i = m - 1; i >= l; --i) {
                const int in = (i > l) ? i - 1 : i;
                const double alpha_n = t.dg(in), bb_n = t.of(in);
                const double r = pp + bb;
                if (!(r > 1e-290)) { failed = true; break; }
                const double e2new = s * r;
                const double rinv = fast_rcp_big(r), pinv = fast_rcp_big(pp);
                const double oldc = c;
                c = pp * rinv; s = bb * rinv;
                const double oldgam = gamma;
                gamma = fma(c, alpha - sigma, -(s * oldgam));
                const double dnew = oldgam + (alpha - gamma);
                t.dg(i + 1) = dnew;
                pp = (c != 0.0) ? (gamma * gamma) * (r * pinv) : oldc * bb;
                if (i != m - 1) {
                    t.of(i + 1) = e2new;
                    put_bit(small, fb, tiny_sq(e2new, dnew, dnext));
                }
                fb.down();
                dnext = dnew; alpha = alpha_n; bb = bb_n;
            }
            if (failed) break;
            const double e2l = s * pp, dl = sigma + gamma;
            t.of(l) = e2l; t.dg(l) = dl;
            small.put(l, tiny_sq(e2l, dl, dnext));
        }
        for (int i = 1; i < d; ++i) {                      // ascending order (insertion sort; the QL output is close to sorted)
            const double x = t.dg(i);
            int j = i - 1;
            while (j >= 0 && t.dg(j) > x) { t.dg(j + 1) = t.dg(j); --j; }
            t.dg(j + 1) = x;
        }
        if (valid) {
            if (failed) t.dg(0) = __longlong_as_double(0x7ff8000000000000LL);      // poisons the point: it takes the fallback
            for (int m = 0; m < d; ++m) a.bs.ev[p * d + m] = t.dg(m);
        }
        __syncwarp();
    }
}

#else
template <bool LOCAL>
__global__ void __launch_bounds__(kWarpsBig * 32) eig_scalar_big_kernel(BigQlArgs a) {
    extern __shared__ double sm[];
    const ProblemView& pv = a.pv;
    const int d = pv.d, f = pv.f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    QlLaneBig<LOCAL> t;
    t.arr = LOCAL ? nullptr : sm + (size_t)warp * 2 * d * 32 + lane;
    t.d = d;
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.overflow_count = 0;      // the eigenvector kernel (stream-ordered after us) appends
    const long long nbatch = (a.n + 31) / 32;
    for (long long batch = (long long)blockIdx.x * kWarpsBig + warp; batch < nbatch; batch += (long long)gridDim.x * kWarpsBig) {
        const long long p = batch * 32 + lane;
        const bool valid = p < a.n;
        const long long pc = valid ? p : a.n - 1;
        double hchi[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) hchi[k] = (k < f && !a.use_tri) ? 0.5 * a.chis[pc * f + k] : 0.0;
        double sum = 0.0;
        for (int m = 0; m < d; ++m) {
            double acc, off;
            if (a.use_tri) {
                acc = a.bs.tri_d[pc * d + m];
                off = (m < d - 1) ? a.bs.tri_e[pc * d + m] : 0.0;
            } else {
                acc = 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (k < f) {
                        const double n = pv.occ[m * f + k];
                        acc = acc + (pv.omegas[k] * n + hchi[k] * (n * n));
                    }
                }
                off = (m < d - 1) ? pv.offd[m] : 0.0;
            }
            t.dg(m) = acc;
            t.of(m) = off;
            sum += acc;
        }
        const double mean = sum / (double)d;
        for (int m = 0; m < d; ++m) {
            const double x = t.dg(m) - mean;
            t.dg(m) = x;
            if (valid) a.bs.adiag[p * d + m] = x;         // what the recurrences of the eigenvector kernel run on
        }
        Mask128 small = t.scan_small();
        int l = 0, iter = 0;
        bool failed = false;
        while (true) {
            const int lnew = small.first_clear_ge(l, d - 1);
            if (lnew < 0) break;
            if (lnew != l) iter = 0;
            l = lnew;
            const int m = small.first_set_ge(l);
            if (++iter > 60) { failed = true; break; }
            double g = t.wilkinson_g(l, m);
            double s = 1.0, c = 1.0, pp = 0.0, dnext = 0.0;
            bool broke = false;
            double d_hi = t.dg(m), d_i = t.dg(m - 1), e_i = t.of(m - 1);
            WalkBit fb; fb.at(m);
            for (int i = m - 1; i >= l; --i) {
                const int in = (i > l) ? i - 1 : i;
                const double d_n = t.dg(in), e_n = t.of(in);
                const double fo = s * e_i, bo = c * e_i;
                const double r2 = fma(fo, fo, g * g);
                if (r2 < kTinyR2) { t.of(i + 1) = 0.0; t.dg(i + 1) = d_hi - pp; t.of(m) = 0.0; broke = true; break; }
                const double rinv = rsqrt_fast(r2);
                const double r = r2 * rinv;
                s = fo * rinv; c = g * rinv;
                g = d_hi - pp;
                const double rr = fma(d_i - g, s, 2.0 * c * bo);
                pp = s * rr;
                const double dnew = g + pp;
                t.dg(i + 1) = dnew;
                g = fma(c, rr, -bo);
                t.of(i + 1) = r;                           // of(m) is reset below
                put_bit(small, fb, r <= kEps * (fabs(dnew) + fabs(dnext)));
                fb.down();
                dnext = dnew; d_hi = d_i; d_i = d_n; e_i = e_n;
            }
            if (broke) { small = t.scan_small(); continue; }
            const double dlnew = d_hi - pp;
            t.dg(l) = dlnew; t.of(l) = g; t.of(m) = 0.0;
            small.set(m);
            small.put(l, fabs(g) <= kEps * (fabs(dlnew) + fabs(dnext)));
        }
        for (int i = 1; i < d; ++i) {                      // ascending order (insertion sort; the QL output is close to sorted)
            const double x = t.dg(i);
            int j = i - 1;
            while (j >= 0 && t.dg(j) > x) { t.dg(j + 1) = t.dg(j); --j; }
            t.dg(j + 1) = x;
        }
        if (valid) {
            if (failed) t.dg(0) = __longlong_as_double(0x7ff8000000000000LL);      // poisons the point: it takes the fallback
            for (int m = 0; m < d; ++m) a.bs.ev[p * d + m] = t.dg(m);
        }
        __syncwarp();
    }
}

#endif

}  // namespace

struct BigWorkspace {
    int device = 0, sms = 0, capit = 0, maxr = 0;
    char* blob = nullptr;
    size_t blob_bytes = 0;
    int* overflow_count = nullptr;
    unsigned short* pair_tab = nullptr;   // register-resident replay: index pairs of the padded dimension
    int regD = 0;                         // padded dimension of the register-resident kernels, 0 = shared-memory kernels
    bool ql_local = false;                // scalar QL keeps the tridiagonal in local memory (more resident warps)
    int ql_ctas = 8;                      // ... CTAs of kWarpsBig warps per SM in that mode
    size_t budget = 0;                    // bytes the workspace may take (from the free HBM, queried when it has to grow)
    // direct path: two chunk buffers, the Householder + eigenvalue kernels of chunk c+1 run on `aux` beside the eigenvector /
    // evolution / gradient kernel of chunk c
    bool direct = false;
    int* fb_count = nullptr;              // [2]
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_a[2] = {nullptr, nullptr}, ev_b[2] = {nullptr, nullptr};
};

bool big_eligible(const ProblemView& pv) {
    if (std::getenv("QTET_NO_BIG_SPLIT")) return false;
    if (pv.d <= 32 && pv.tridiag) return false;          // the register-resident split path is faster there
    return pv.d >= 4 && pv.d <= 115 && pv.f <= 8;         // 115: V and K must both fit 227 KB of shared memory
}

void big_workspace_free(BigWorkspace* ws) {
    if (!ws) return;
    if (ws->blob) cudaFree(ws->blob);
    if (ws->overflow_count) cudaFree(ws->overflow_count);
    if (ws->pair_tab) cudaFree(ws->pair_tab);
    if (ws->fb_count) cudaFree(ws->fb_count);
    if (ws->aux) cudaStreamDestroy(ws->aux);
    if (ws->ev_fork) cudaEventDestroy(ws->ev_fork);
    for (int i = 0; i < 2; ++i) {
        if (ws->ev_a[i]) cudaEventDestroy(ws->ev_a[i]);
        if (ws->ev_b[i]) cudaEventDestroy(ws->ev_b[i]);
    }
    delete ws;
}

int launch_generic_list(const ProblemView& pv, const double* chis, const int* list, const int* count, long long max_count,
                        int site, double* loss, double* grad, int32_t* tstar, double* trace, cudaStream_t st);

int launch_big(const ProblemView& pv, BigWorkspace** wsp, const double* chis, int64_t B, int site, double* loss,
               double* grad, int32_t* tstar, double* trace, cudaStream_t st, Profiler* prof) {
    if (B <= 0) return QTET_OK;
    if (!big_eligible(pv)) { set_error("large-d split path not eligible for this problem"); return QTET_EUNSUPPORTED; }
    BigWorkspace* ws = *wsp;
    const int d = pv.d;
    if (ws && ws->direct) { cudaDeviceSynchronize(); big_workspace_free(ws); ws = nullptr; *wsp = nullptr; }     // path switched: rebuild
    if (!ws) {
        ws = new BigWorkspace();
        *wsp = ws;
        QTET_CUDA(cudaGetDevice(&ws->device));
        QTET_CUDA(cudaDeviceGetAttribute(&ws->sms, cudaDevAttrMultiProcessorCount, ws->device));
        ws->capit = d * d + 8 * d;                        // stream rows per batch
        ws->maxr = (3 * d > 128) ? 3 * d : 128;
        if (const char* env = std::getenv("QTET_SPLIT_CAPIT")) {   // test hook: force the overflow -> fused fallback
            const int v = std::atoi(env);
            if (v > 0 && v < ws->capit) ws->capit = v;
        }
        QTET_CUDA(cudaMalloc(&ws->overflow_count, sizeof(int)));
        // rows in registers when the dimension fits (d <= 101); QTET_BIG_SMEM=1 keeps the shared-memory kernels (cross-check)
        ws->regD = (std::getenv("QTET_BIG_SMEM") || pv.T > bigreg_max_timesteps()) ? 0 : bigreg_padded_dim(d);
        ws->ql_local = true; ws->ql_ctas = 16;
        if (const char* env = std::getenv("QTET_QL_LOCAL")) { ws->ql_local = std::atoi(env) > 0; if (std::atoi(env) > 1) ws->ql_ctas = std::atoi(env); }
        if (ws->regD) {
            std::vector<unsigned short> tab;
            for (int i = 0; i < ws->regD; ++i)
                for (int j = i + 1; j < ws->regD; ++j) tab.push_back((unsigned short)(i | (j << 8)));
            if (tab.empty()) tab.push_back(0);
            QTET_CUDA(cudaMalloc(&ws->pair_tab, tab.size() * sizeof(unsigned short)));
            QTET_CUDA(cudaMemcpy(ws->pair_tab, tab.data(), tab.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
        }
    }
    const bool dense = !pv.tridiag;
    const size_t per_batch = (size_t)ws->capit * 32 * sizeof(double2) + (size_t)ws->maxr * 32 * 4 + (size_t)ws->maxr * 4 + 4 +
                             32 * ((size_t)d * 8 + 4 + (dense ? (2 * (size_t)d * 8 + (size_t)d * d * 8) : 0));
    // The scalar QL kernel is a serial chain per point (milliseconds per batch of 32 points, whatever the machine
    // load): its cost is amortised only if MANY batches are in flight, so the chunk takes what the card can spare --
    // up to 40 % of the free HBM, at most 8 GiB (QTET_BIG_WS_GIB overrides) -- rather than a fixed few hundred MiB.
    const long long all_b0 = (B + 31) / 32;
    if (ws->budget == 0 || (size_t)all_b0 * per_batch > ws->blob_bytes) {      // (re)query only when the buffer would have to grow
        size_t budget = (size_t)4 << 30;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            const size_t avail = free_b + ws->blob_bytes;           // our own blob is reusable
            size_t want = (size_t)(0.4 * (double)avail);
            if (want > ((size_t)8 << 30)) want = (size_t)8 << 30;      // 8 GiB holds > 45 000 points per launch at d = 101: every warp slot busy
            if (want > budget) budget = want;
        } else cudaGetLastError();
        if (const char* env = std::getenv("QTET_BIG_WS_GIB")) { const double g = std::atof(env); if (g > 0) budget = (size_t)(g * 1073741824.0); }
        ws->budget = budget;
    }
    const size_t budget = ws->budget;
    long long chunk_b = (long long)budget / (long long)per_batch;
    if (chunk_b < 64) chunk_b = 64;
    const long long all_b = (B + 31) / 32;
    if (chunk_b > all_b) chunk_b = all_b;
    auto align = [](size_t x) { return (x + 255) / 256 * 256; };
    long long chunk = 0;
    size_t o_hdr = 0, o_it0 = 0, o_nr = 0, o_ev = 0, o_list = 0, o_trid = 0, o_trie = 0, o_q = 0, total = 0;
    while (true) {
        chunk = chunk_b * 32;
        o_hdr = align((size_t)chunk_b * ws->capit * 32 * sizeof(double2));
        o_it0 = o_hdr + align((size_t)chunk_b * ws->maxr * 32 * 4);
        o_nr = o_it0 + align((size_t)chunk_b * ws->maxr * 4);
        o_ev = o_nr + align((size_t)chunk_b * 4);
        o_list = o_ev + align((size_t)chunk * d * 8);
        o_trid = o_list + align((size_t)chunk * 4);
        o_trie = o_trid + (dense ? align((size_t)chunk * d * 8) : 0);
        o_q = o_trie + (dense ? align((size_t)chunk * d * 8) : 0);
        total = o_q + (dense ? align((size_t)chunk * d * d * 8) : 0);
        if (total <= ws->blob_bytes) break;
        if (ws->blob) { cudaDeviceSynchronize(); cudaFree(ws->blob); }
        ws->blob = nullptr; ws->blob_bytes = 0;
        if (cudaMalloc(&ws->blob, total) == cudaSuccess) { ws->blob_bytes = total; break; }
        cudaGetLastError();
        ws->blob = nullptr;
        // the card cannot spare that much (fragmentation, other tenants): halve the chunk and remember the smaller budget
        if (chunk_b <= 64) { set_error("cudaMalloc(large-d workspace) failed"); return QTET_ENOMEM; }
        chunk_b = (chunk_b / 2 < 64) ? 64 : chunk_b / 2;
        ws->budget = (size_t)chunk_b * per_batch;
    }
    BigStream bs;
    bs.stream = (double2*)ws->blob; bs.hdr = (unsigned*)(ws->blob + o_hdr); bs.it0 = (int*)(ws->blob + o_it0);
    bs.nrounds = (int*)(ws->blob + o_nr); bs.ev = (double*)(ws->blob + o_ev);
    bs.capit = ws->capit; bs.maxr = ws->maxr;
    if (dense) { bs.tri_d = (double*)(ws->blob + o_trid); bs.tri_e = (double*)(ws->blob + o_trie); bs.qmat = (double*)(ws->blob + o_q); }
    int* overflow_list = (int*)(ws->blob + o_list);

    // Scalar QL storage: the shared-memory variant has the shorter chain per batch but only 228 KB / (2 d 256 B) warps
    // per SM; when a chunk holds more batches than that many slots the local-memory variant (4x the warps) wins.
    const size_t smem_sh = (size_t)kWarpsBig * 2 * d * 32 * sizeof(double);
    QTET_CUDA(cudaFuncSetAttribute(ql_scalar_big_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sh));
    QTET_CUDA(cudaFuncSetAttribute(ql_scalar_big_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int per_sm_sh = 0, per_sm_lo = 0;
    QTET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_sh, ql_scalar_big_kernel<false>, kWarpsBig * 32, smem_sh));
    QTET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_lo, ql_scalar_big_kernel<true>, kWarpsBig * 32, 0));
    if (per_sm_sh < 1) per_sm_sh = 1;
    if (per_sm_lo < 1) per_sm_lo = 1;
    if (per_sm_lo > ws->ql_ctas) per_sm_lo = ws->ql_ctas;
    const long long slots_sh = (long long)ws->sms * per_sm_sh * kWarpsBig;
    const bool ql_local = ws->ql_local && chunk_b > slots_sh;
    const size_t smemA = ql_local ? 0 : smem_sh;
    auto qlk = ql_local ? ql_scalar_big_kernel<true> : ql_scalar_big_kernel<false>;
    const int per_sm_a = ql_local ? per_sm_lo : per_sm_sh;

    for (long long lo = 0; lo < B; lo += chunk) {
        const long long n = (B - lo < chunk) ? (B - lo) : chunk;
        const double* cchis = chis + lo * pv.f;
        double* closs = loss ? loss + lo : nullptr;
        double* cgrad = grad ? grad + lo * pv.f : nullptr;
        int32_t* cts = tstar ? tstar + lo : nullptr;
        double* ctrace = trace ? trace + lo * pv.T : nullptr;
        int rc;
        if (dense) {
            if (prof) prof->begin(2, n, st);
            rc = ws->regD ? launch_hh_reg(pv, cchis, n, bs, ws->sms, st)
                          : launch_generic_mode(pv, cchis, n, site, nullptr, nullptr, nullptr, nullptr, st, 1, bs);
            if (prof) prof->finish(2, st);
            if (rc) return rc;
        }
        QTET_CUDA(cudaMemsetAsync(ws->overflow_count, 0, sizeof(int), st));
        BigQlArgs qa;
        qa.pv = pv; qa.chis = cchis; qa.n = n; qa.bs = bs; qa.use_tri = dense ? 1 : 0;
        qa.overflow_list = overflow_list; qa.overflow_count = ws->overflow_count;
        long long gridA = (long long)ws->sms * per_sm_a;
        const long long needA = ((n + 31) / 32 + kWarpsBig - 1) / kWarpsBig;
        if (gridA > needA) gridA = needA;
        if (prof) prof->begin(0, n, st);
        qlk<<<(unsigned)gridA, kWarpsBig * 32, smemA, st>>>(qa);
        count_launch();
        if (prof) prof->finish(0, st);
        QTET_CUDA(cudaGetLastError());
        if (prof) prof->begin(1, n, st);
        rc = ws->regD ? launch_replay_reg(pv, cchis, n, site, bs, ws->pair_tab, closs, cgrad, cts, ctrace, ws->sms, st)
                      : launch_generic_mode(pv, cchis, n, site, closs, cgrad, cts, ctrace, st, 2, bs);
        if (prof) prof->finish(1, st);
        if (rc) return rc;
        rc = launch_generic_list(pv, cchis, overflow_list, ws->overflow_count, n, site, closs, cgrad, cts, ctrace, st);
        if (rc) return rc;
    }
    return QTET_OK;
}

// ------------------------------------------------------------------------------------------------
// Direct large-d path: [Householder] -> eigenvalues (scalar QL, no stream) -> eigenvectors by recurrences [+ V = Q Z] ->
// evolution -> gradient.  Chunks are double-buffered: the Householder and eigenvalue kernels of chunk c+1 run on an auxiliary
// stream while the eigenvector / evolution / gradient kernel of chunk c runs on the caller's stream.
// ------------------------------------------------------------------------------------------------
bool big_direct_eligible(const ProblemView& pv) {
    if (!big_eligible(pv)) return false;
    if (std::getenv("QTET_BIG_SMEM") || pv.T > bigreg_max_timesteps()) return false;
    const int D = bigreg_padded_dim(pv.d);
    if (D == 0 || D > bigreg_direct_max_dim(!pv.tridiag)) return false;
    if (pv.tridiag && !pv.offd_nonzero) return false;
    return true;
}

static long long big_direct_chunk(const ProblemView& pv, long long B) {
    const bool dense = !pv.tridiag;
    const size_t per_point = (size_t)pv.d * 8 * (dense ? 4 : 2) + 4 + (dense ? (size_t)pv.d * pv.d * 8 + hh_frag_scratch_bytes(pv.d) : 0);
    // Four chunks per call, but never fewer than 65536 points: the eigenvalue kernel is one thread per point and a pure latency
    // chain -- with 16384 points it runs 3.5 warps per SM and takes as long as with 65536 (measured: trimer N=10, 4 x 0.84 ms
    // against 1 x 2.0 ms per 65536 points), and the kernels of consecutive chunks do not overlap anyway (each fills the machine).
    long long chunk = ((B + 3) / 4 + 31) / 32 * 32;
    const long long lo = 65536;
    (void)dense;
    if (chunk < lo) chunk = lo;
    long long cap = (long long)(((size_t)4 << 30) / per_point) / 32 * 32;     // <= 4 GiB per buffer
    if (const char* env = std::getenv("QTET_BIG_WS_GIB")) { const double g = std::atof(env); if (g > 0) cap = (long long)(g * 0.5 * 1073741824.0 / (double)per_point) / 32 * 32; }
    if (cap < 32) cap = 32;
    if (const char* env = std::getenv("QTET_BIG_CHUNK")) { const long long v = std::atoll(env) / 32 * 32; if (v >= 32) chunk = v; }
    if (chunk > cap) chunk = cap;
    const long long all = (B + 31) / 32 * 32;
    if (chunk > all) chunk = all;
    return chunk;
}

int launch_big_direct(const ProblemView& pv, BigWorkspace** wsp, const double* chis, int64_t B, int site, double* loss,
                      double* grad, int32_t* tstar, double* trace, cudaStream_t st, Profiler* prof, const ChunkHook* hook) {
    if (B <= 0) return QTET_OK;
    if (!big_direct_eligible(pv)) { set_error("large-d direct path not eligible for this problem"); return QTET_EUNSUPPORTED; }
    BigWorkspace* ws = *wsp;
    const int d = pv.d;
    const bool dense = !pv.tridiag;
    if (ws && !ws->direct) { cudaDeviceSynchronize(); big_workspace_free(ws); ws = nullptr; *wsp = nullptr; }   // path switched: rebuild
    if (!ws) {
        ws = new BigWorkspace();
        *wsp = ws;
        ws->direct = true;
        QTET_CUDA(cudaGetDevice(&ws->device));
        QTET_CUDA(cudaDeviceGetAttribute(&ws->sms, cudaDevAttrMultiProcessorCount, ws->device));
        ws->regD = bigreg_padded_dim(d);
        ws->maxr = 4;                                     // no round headers: the replay region of the kernel stays empty
        ws->ql_local = true; ws->ql_ctas = 16;
        if (const char* env = std::getenv("QTET_QL_LOCAL")) { ws->ql_local = std::atoi(env) > 0; if (std::atoi(env) > 1) ws->ql_ctas = std::atoi(env); }
        std::vector<unsigned short> tab;
        for (int i = 0; i < ws->regD; ++i)
            for (int j = i + 1; j < ws->regD; ++j) tab.push_back((unsigned short)(i | (j << 8)));
        if (tab.empty()) tab.push_back(0);
        QTET_CUDA(cudaMalloc(&ws->pair_tab, tab.size() * sizeof(unsigned short)));
        QTET_CUDA(cudaMemcpy(ws->pair_tab, tab.data(), tab.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
        QTET_CUDA(cudaMalloc(&ws->fb_count, 2 * sizeof(int)));
        QTET_CUDA(cudaStreamCreateWithFlags(&ws->aux, cudaStreamNonBlocking));
        QTET_CUDA(cudaEventCreateWithFlags(&ws->ev_fork, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) {
            QTET_CUDA(cudaEventCreateWithFlags(&ws->ev_a[i], cudaEventDisableTiming));
            QTET_CUDA(cudaEventCreateWithFlags(&ws->ev_b[i], cudaEventDisableTiming));
        }
    }
    const long long chunk = big_direct_chunk(pv, B);
    auto align = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t o_ad = align((size_t)chunk * d * 8);
    const size_t o_list = o_ad + align((size_t)chunk * d * 8);
    const size_t o_trid = o_list + align((size_t)chunk * 4);
    const size_t o_trie = o_trid + (dense ? align((size_t)chunk * d * 8) : 0);
    const size_t o_q = o_trie + (dense ? align((size_t)chunk * d * 8) : 0);
    const size_t o_rfl = o_q + (dense ? align((size_t)chunk * d * d * 8) : 0);
    const size_t rfl_pp = dense ? hh_frag_scratch_bytes(d) : 0;
    const size_t buf_bytes = o_rfl + align((size_t)chunk * rfl_pp);
    if (2 * buf_bytes > ws->blob_bytes) {
        if (ws->blob) { cudaDeviceSynchronize(); cudaFree(ws->blob); }
        ws->blob = nullptr; ws->blob_bytes = 0;
        if (cudaMalloc(&ws->blob, 2 * buf_bytes) != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(large-d direct workspace) failed"); return QTET_ENOMEM; }
        ws->blob_bytes = 2 * buf_bytes;
    }
    // eigenvalue kernel: shared-memory tridiagonal when the chunk has fewer batches than that variant has warp slots
    const size_t smem_sh = (size_t)kWarpsBig * 2 * d * 32 * sizeof(double);
    QTET_CUDA(cudaFuncSetAttribute(eig_scalar_big_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sh));
    QTET_CUDA(cudaFuncSetAttribute(eig_scalar_big_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int per_sm_sh = 0, per_sm_lo = 0;
    QTET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_sh, eig_scalar_big_kernel<false>, kWarpsBig * 32, smem_sh));
    QTET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_lo, eig_scalar_big_kernel<true>, kWarpsBig * 32, 0));
    if (per_sm_sh < 1) per_sm_sh = 1;
    if (per_sm_lo < 1) per_sm_lo = 1;
    if (per_sm_lo > ws->ql_ctas) per_sm_lo = ws->ql_ctas;
    const long long slots_sh = (long long)ws->sms * per_sm_sh * kWarpsBig;
    const bool ql_local = ws->ql_local && chunk / 32 > slots_sh;
    const size_t smemA = ql_local ? 0 : smem_sh;
    auto eigk = ql_local ? eig_scalar_big_kernel<true> : eig_scalar_big_kernel<false>;
    const int per_sm_a = ql_local ? per_sm_lo : per_sm_sh;

    QTET_CUDA(cudaEventRecord(ws->ev_fork, st));
    QTET_CUDA(cudaStreamWaitEvent(ws->aux, ws->ev_fork, 0));
    int c = 0;
    for (long long lo = 0; lo < B; lo += chunk, ++c) {
        const int buf = c & 1;
        char* base = ws->blob + (size_t)buf * buf_bytes;
        const long long n = (B - lo < chunk) ? (B - lo) : chunk;
        BigStream bs;
        bs.ev = (double*)base; bs.adiag = (double*)(base + o_ad);
        bs.capit = 0; bs.maxr = ws->maxr;
        if (dense) { bs.tri_d = (double*)(base + o_trid); bs.tri_e = (double*)(base + o_trie); bs.qmat = (double*)(base + o_q); }
        if (rfl_pp) bs.rfl = (double*)(base + o_rfl);
        int* fb_list = (int*)(base + o_list);
        int* fb_count = ws->fb_count + buf;
        const double* cchis = chis + lo * pv.f;
        double* closs = loss ? loss + lo : nullptr;
        double* cgrad = grad ? grad + lo * pv.f : nullptr;
        int32_t* cts = tstar ? tstar + lo : nullptr;
        double* ctrace = trace ? trace + lo * pv.T : nullptr;
        // aux: Householder + eigenvalues of chunk c, once the consumer of chunk c-2 has released this buffer
        if (c >= 2) QTET_CUDA(cudaStreamWaitEvent(ws->aux, ws->ev_b[buf], 0));
        if (hook && hook->before) { const int hrc = hook->before(hook->ctx, lo, n, ws->aux); if (hrc) return hrc; }
        int rc;
        if (dense) {
            if (prof) prof->begin(2, n, ws->aux);
            rc = launch_hh_reg(pv, cchis, n, bs, ws->sms, ws->aux);
            if (prof) prof->finish(2, ws->aux);
            if (rc) return rc;
        }
        BigQlArgs qa;
        qa.pv = pv; qa.chis = cchis; qa.n = n; qa.bs = bs; qa.use_tri = dense ? 1 : 0;
        qa.overflow_list = fb_list; qa.overflow_count = fb_count;
        long long gridA = (long long)ws->sms * per_sm_a;
        const long long needA = ((n + 31) / 32 + kWarpsBig - 1) / kWarpsBig;
        if (gridA > needA) gridA = needA;
        if (prof) prof->begin(0, n, ws->aux);
        eigk<<<(unsigned)gridA, kWarpsBig * 32, smemA, ws->aux>>>(qa);
        count_launch();
        if (prof) prof->finish(0, ws->aux);
        QTET_CUDA(cudaGetLastError());
        QTET_CUDA(cudaEventRecord(ws->ev_a[buf], ws->aux));
        // caller's stream: eigenvectors, evolution, gradient of chunk c
        QTET_CUDA(cudaStreamWaitEvent(st, ws->ev_a[buf], 0));
        if (prof) prof->begin(1, n, st);
        rc = launch_replay_reg(pv, cchis, n, site, bs, ws->pair_tab, closs, cgrad, cts, ctrace, ws->sms, st, true, fb_list, fb_count);
        if (prof) prof->finish(1, st);
        if (rc) return rc;
        rc = launch_generic_list(pv, cchis, fb_list, fb_count, n, site, closs, cgrad, cts, ctrace, st);
        if (rc) return rc;
        QTET_CUDA(cudaEventRecord(ws->ev_b[buf], st));
        if (hook && hook->after) { const int hrc = hook->after(hook->ctx, lo, n, st); if (hrc) return hrc; }
    }
    return QTET_OK;
}

}  // namespace qtet
