// Register-resident kernels of the large-d split path (qtet_big.cu): one CTA per parameter point, one THREAD per
// row, and that row (of the dense matrix A, of the Householder factor Q, of the eigenvector matrix V) lives in
// REGISTERS for the whole phase.  Everything a row needs from the others is a broadcast operand read from shared
// memory with LDS.128 (two operands per load, each feeding one or two DFMAs), so the FP64 pipe rather than the
// shared-memory crossbar is the busy unit -- the shared-memory formulation these kernels replace needed two LDS
// per DFMA and ran at 3-15 % of the FP64 peak.
//
//   hh_reg_kernel<D, NW>      dense H only (trimer, chains): H build (HamiltonianLoss.py:124-176), Householder
//                             tridiagonalisation with the row of A in registers, then the forward accumulation of
//                             Q = H_0 H_1 ... in the same registers; writes (diag, off) and Q (column-major).
//   replay_reg_kernel<D, NW>  V <- I (or Q); replays the plane rotations recorded by ql_scalar_big_kernel on the
//                             register rows (statically indexed, plane groups skipped with warp-uniform branches),
//                             then psi(t) = V (c . exp(-i e t)) (HamiltonianLoss.py:204-215) as FP64 tensor-core products
//                             (mma.sync m8n8k4: V parked in shared memory is the A operand, the phase table of 8 time
//                             samples the B operand), loss (:227-231) and the analytic gradient through the symmetrised
//                             Daleckii-Krein kernel (replaces Optimizer.py:85-91).
//
// D is the compile-time padded dimension (register arrays must be statically indexed), NW = ceil(D / 32) warps.
#include <cstdlib>
#include <type_traits>
#include <utility>

#include "qtet_big.cuh"

namespace qtet {

namespace {

constexpr int kRingR = 4;      // trips of (c, s) in flight per CTA (cp.async ring)
#ifndef QTET_FUSE
#define QTET_FUSE 2
#endif
constexpr int kFuse = QTET_FUSE;   // QL rounds replayed together per trip
[[maybe_unused]] constexpr int kTC = 16;        // time samples per evolution chunk (DFMA evolution)
#ifndef QTET_EVO_MMA
#define QTET_EVO_MMA 1         // 1: time evolution psi = V z on the FP64 tensor path (mma.sync m8n8k4), 0: DFMA loop with broadcast z
#endif
#ifndef QTET_S_CYCLIC
#define QTET_S_CYCLIC 1        // storage order of the pair kernel S: 1 = [i][o] (cyclic offsets), 0 = packed upper triangle
#endif
constexpr int kNC = 16;        // real columns (= 8 time samples x re|im) per DMMA evolution chunk

// D(8x8) += A(8x4, row) B(4x8, col) in FP64: lane l holds A[l/4][l%4], B[l%4][l/4], C[l/4][2(l%4) .. +1]
__device__ __forceinline__ void dmma884(double2& c, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c.x), "+d"(c.y) : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Shared-memory address as an opaque 32-bit register: the compiler otherwise rematerialises the window base
// (S2UR SR_CgaCtaId + ULEA, ~25 cycles of scoreboard stall) in every basic block of an unrolled, branchy sweep.
__device__ __forceinline__ unsigned smem_addr(const void* p) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));
    return a;
}
__device__ __forceinline__ double lds64(unsigned addr) {
    double r;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ double2 lds128(unsigned addr) {
    double2 r;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(addr));
    return r;
}

__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}

#ifdef QTET_PHASE_TIMING
__device__ unsigned long long g_big_phase_cycles[16];
#define BPHASE_DECL() long long _t0 = 0
#define BPHASE_START() _t0 = clock64()
#define BPHASE_MARK(idx) do { const long long _t = clock64(); if (tid == 0) atomicAdd(&g_big_phase_cycles[idx], (unsigned long long)(_t - _t0)); _t0 = _t; } while (0)
#else
#define BPHASE_DECL() do {} while (0)
#define BPHASE_START() do {} while (0)
#define BPHASE_MARK(idx) do {} while (0)
#endif

struct RegArgs {
    ProblemView pv;
    const double* chis;
    long long n;
    int site, acceptor;
    BigStream bs;
    const unsigned short* pair_tab;   // [D(D-1)/2]  i | j << 8, lexicographic (i, j > i); used by the D <= 66 kernels
    double* loss; double* grad; int* tstar; double* trace;
    int* fb_list = nullptr;           // direct variant: points handed to the fused generic kernel ...
    int* fb_count = nullptr;          // ... and their count
};

// a[k] and a[k+1] of a register array for a CTA-uniform dynamic k: a jump table, not D predicated moves.
template <int K, int D>
__device__ __forceinline__ double elem(const double (&a)[D]) {
    if constexpr (K < D) return a[K];
    else return 0.0;
}
#define QTET_PICK_CASE(K) case K: x = elem<K>(a); y = elem<K + 1>(a); break;
#define QTET_PICK_8(K) QTET_PICK_CASE(K) QTET_PICK_CASE(K + 1) QTET_PICK_CASE(K + 2) QTET_PICK_CASE(K + 3) \
                       QTET_PICK_CASE(K + 4) QTET_PICK_CASE(K + 5) QTET_PICK_CASE(K + 6) QTET_PICK_CASE(K + 7)
template <int D>
__device__ __forceinline__ void pick2(const double (&a)[D], int k, double& x, double& y) {
    x = 0.0; y = 0.0;
    switch (k) {
        QTET_PICK_8(0) QTET_PICK_8(8) QTET_PICK_8(16) QTET_PICK_8(24) QTET_PICK_8(32) QTET_PICK_8(40) QTET_PICK_8(48) QTET_PICK_8(56)
        QTET_PICK_8(64) QTET_PICK_8(72) QTET_PICK_8(80) QTET_PICK_8(88) QTET_PICK_8(96) QTET_PICK_8(104) QTET_PICK_8(112) QTET_PICK_8(120)
        default: break;
    }
}

// compile-time loop: f(integral_constant<int, 0>) ... f(integral_constant<int, N-1>) (forces full expansion where
// `#pragma unroll` gives up on thousands of iterations)
template <class F, int... Is>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, Is...>) { (f(std::integral_constant<int, Is>{}), ...); }
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) { static_for_impl(f, std::make_integer_sequence<int, N>{}); }

template <int NW>
__device__ __forceinline__ double cta_sum(double v, double* red, int lane, int warp) {
    v = warp_sum(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += red[w];
    return s;
}

// ------------------------------------------------------------------------------------------------
// Householder tridiagonalisation + Q, rows in registers
// ------------------------------------------------------------------------------------------------
template <int D>
struct HhLayout {
    static constexpr int DP = (D + 7) & ~7;
    // doubles
    static constexpr int XS = 0;                   // [2][DP]  raw column below the diagonal (zeros above)
    static constexpr int VS = XS + 2 * DP;         // [2][DP]  reflector
    static constexpr int PS = VS + 2 * DP;         // [2][DP]  p = beta A v
    static constexpr int DIAG = PS + 2 * DP;       // [DP]
    static constexpr int OFF = DIAG + DP;          // [DP]
    static constexpr int BETA = OFF + DP;          // [DP]
    static constexpr int RED = BETA + DP;          // [2][2][8] partial sums + [2] x0
    static constexpr int REFL = RED + 40;          // [D][DP]
    static constexpr int TOTAL = REFL + D * DP;
};

template <int D, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) hh_reg_kernel(RegArgs a_) {
    using L = HhLayout<D>;
    constexpr int DP = L::DP;
    extern __shared__ __align__(16) double sm[];
    const ProblemView& pv = a_.pv;
    const int d = pv.d, f = pv.f;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = tid;
    double* xs = sm + L::XS;
    double* vs = sm + L::VS;
    double* ps = sm + L::PS;
    double* s_diag = sm + L::DIAG;
    double* s_off = sm + L::OFF;
    double* s_beta = sm + L::BETA;
    double* red = sm + L::RED;
    double* refl = sm + L::REFL;
    constexpr int NG8 = DP / 8;                    // column groups of 8
    const unsigned xs_a = smem_addr(xs), vs_a = smem_addr(vs), ps_a = smem_addr(ps), refl_a = smem_addr(refl);

    for (long long b = blockIdx.x; b < a_.n; b += gridDim.x) {
        const double* chi = a_.chis + b * f;
        __syncthreads();
        BPHASE_DECL();
        BPHASE_START();
        // ---- my row of H: hopping part (symmetric, so column `row` read coalesced) + diagonal
        double a[D];
        {
            double dg = 0.0;
            if (row < d)
                for (int k = 0; k < f; ++k)
                    dg = dg + (pv.omegas[k] * pv.occ[row * f + k] + (0.5 * chi[k]) * (pv.occ[row * f + k] * pv.occ[row * f + k]));
#pragma unroll
            for (int j = 0; j < D; ++j) {
                double h = 0.0;
                if (row < d && j < d) h = (j == row) ? dg : pv.offd[(size_t)j * d + row];
                a[j] = h;
            }
        }
        if (row < DP) { xs[row] = 0.0; xs[DP + row] = 0.0; vs[row] = 0.0; vs[DP + row] = 0.0; ps[row] = 0.0; ps[DP + row] = 0.0; }

        BPHASE_MARK(0);
        // ---- reduction: two CTA barriers per column
        for (int k = 0; k < d - 2; ++k) {
            const int par = k & 1;
            double* xsp = xs + par * DP;
            double* vsp = vs + par * DP;
            double* psp = ps + par * DP;
            double* redp = red + par * 18;
            double x, y;
            pick2<D>(a, k, x, y);                        // A[row][k], A[row][k+1]
            if (row == k) s_diag[k] = x;
            const double xi = (row > k && row < d) ? x : 0.0;
            if (row < DP) { xsp[row] = xi; vsp[row] = xi; }
            {
                double part = (row > k + 1) ? xi * xi : 0.0;
                part = warp_sum(part);
                if (lane == 0) redp[warp] = part;
                if (row == k + 1) redp[16] = xi;
            }
            __syncthreads();                             // #1
            double sigma = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) sigma += redp[w];
            const double x0 = redp[16];
            if (sigma == 0.0) {                          // nothing to annihilate in this column (CTA-uniform)
                if (tid == 0) { s_off[k] = x0; s_beta[k] = 0.0; }
                continue;
            }
            const double nrm2 = fma(x0, x0, sigma);
            const double mu = nrm2 * rsqrt(nrm2);          // a few ulp off sqrt: only moves the tridiagonal entry by that much;
            const double alpha = (x0 >= 0.0) ? -mu : mu;   // H stays orthogonal because beta is formed from the v actually used
            const double v0 = x0 - alpha;
            const double beta = 2.0 * fast_rcp(fma(v0, v0, sigma));
            const double vi = (row == k + 1) ? v0 : ((row > k + 1) ? xi : 0.0);
            if (row == k + 1) vsp[row] = v0;
            if (row < DP) refl[(size_t)k * DP + row] = vi;
            if (tid == 0) { s_off[k] = alpha; s_beta[k] = beta; }
            const bool warp_active = (warp * 32 + 31 > k);
            const unsigned xa_ = xs_a + par * (DP * 8), va_ = vs_a + par * (DP * 8), pa_ = ps_a + par * (DP * 8);
            double p = 0.0;
            if (warp_active) {
                double acc[4] = {0.0, 0.0, 0.0, 0.0};
                bool live = true;
                static_for<NG8>([&](auto Gc) {
                    constexpr int j0 = (NG8 - 1 - decltype(Gc)::value) * 8;
                    if (live && j0 + 7 <= k) live = false;
                    if (live) {
#pragma unroll
                        for (int u = 0; u < 8; u += 2) {
                            if (j0 + u < D) {
                                const double2 x2 = lds128(xa_ + (j0 + u) * 8);
                                acc[u & 3] = fma(a[j0 + u < D ? j0 + u : 0], x2.x, acc[u & 3]);
                                if (j0 + u + 1 < D) acc[(u + 1) & 3] = fma(a[j0 + u + 1 < D ? j0 + u + 1 : 0], x2.y, acc[(u + 1) & 3]);
                            }
                        }
                    }
                });
                // the reflector differs from the raw column only in entry k+1 (v0 = x0 - alpha)
                p = beta * (((acc[0] + acc[1]) + (acc[2] + acc[3])) - alpha * y);
                if (!(row > k && row < d)) p = 0.0;
            }
            if (row < DP) psp[row] = p;
            {
                double vp = warp_sum(vi * p);
                if (lane == 0) redp[8 + warp] = vp;
            }
            __syncthreads();                             // #2
            double vp = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) vp += redp[8 + w];
            const double kk = 0.5 * beta * vp;
            // A -= v w^T + w v^T with w = p - kk v  ==>  A_ij -= v_i p_j + (p_i - 2 kk v_i) v_j
            const double a2 = fma(-2.0 * kk, vi, p);
            if (warp_active) {
                bool live = true;
                static_for<NG8>([&](auto Gc) {
                    constexpr int j0 = (NG8 - 1 - decltype(Gc)::value) * 8;
                    if (live && j0 + 7 <= k) live = false;
                    if (live) {
#pragma unroll
                        for (int u = 0; u < 8; u += 2) {
                            if (j0 + u < D) {
                                const double2 p2 = lds128(pa_ + (j0 + u) * 8);
                                const double2 v2 = lds128(va_ + (j0 + u) * 8);
                                constexpr int dm = D - 1;
                                const int c0 = (j0 + u <= dm) ? j0 + u : 0, c1 = (j0 + u + 1 <= dm) ? j0 + u + 1 : 0;
                                a[c0] = fma(-vi, p2.x, fma(-a2, v2.x, a[c0]));
                                if (j0 + u + 1 < D) a[c1] = fma(-vi, p2.y, fma(-a2, v2.y, a[c1]));
                            }
                        }
                    }
                });
            }
        }
        BPHASE_MARK(1);
        // last 2 x 2 block
        if (d >= 2) {
            double x, y;
            pick2<D>(a, d - 2, x, y);
            if (row == d - 2) s_diag[d - 2] = x;
            if (row == d - 1) { s_off[d - 2] = x; s_diag[d - 1] = y; s_off[d - 1] = 0.0; }
        } else if (row == 0) {
            s_diag[0] = a[0]; s_off[0] = 0.0;
        }
        __syncthreads();
        if (row < d) {
            a_.bs.tri_d[(size_t)b * d + row] = s_diag[row];
            a_.bs.tri_e[(size_t)b * d + row] = s_off[row];
        }

        // ---- Q = H_0 H_1 ... H_{d-3}: my row, accumulated forward (no barrier: reflectors are all in shared memory)
#pragma unroll
        for (int j = 0; j < D; ++j) a[j] = (j == row) ? 1.0 : 0.0;
        for (int k = 0; k < d - 2; ++k) {
            const double beta = s_beta[k];
            if (beta == 0.0) continue;
            const unsigned ra_ = refl_a + (unsigned)k * (DP * 8);
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            {
                bool live = true;
                static_for<NG8>([&](auto Gc) {
                    constexpr int j0 = (NG8 - 1 - decltype(Gc)::value) * 8;
                    if (live && j0 + 7 <= k) live = false;
                    if (live) {
#pragma unroll
                        for (int u = 0; u < 8; u += 2) {
                            if (j0 + u < D) {
                                const double2 r2 = lds128(ra_ + (j0 + u) * 8);
                                acc[u & 3] = fma(a[j0 + u < D ? j0 + u : 0], r2.x, acc[u & 3]);
                                if (j0 + u + 1 < D) acc[(u + 1) & 3] = fma(a[j0 + u + 1 < D ? j0 + u + 1 : 0], r2.y, acc[(u + 1) & 3]);
                            }
                        }
                    }
                });
            }
            const double u_ = beta * ((acc[0] + acc[1]) + (acc[2] + acc[3]));
            {
                bool live = true;
                static_for<NG8>([&](auto Gc) {
                    constexpr int j0 = (NG8 - 1 - decltype(Gc)::value) * 8;
                    if (live && j0 + 7 <= k) live = false;
                    if (live) {
#pragma unroll
                        for (int u = 0; u < 8; u += 2) {
                            if (j0 + u < D) {
                                const double2 r2 = lds128(ra_ + (j0 + u) * 8);
                                constexpr int dm = D - 1;
                                const int c0 = (j0 + u <= dm) ? j0 + u : 0, c1 = (j0 + u + 1 <= dm) ? j0 + u + 1 : 0;
                                a[c0] = fma(-u_, r2.x, a[c0]);
                                if (j0 + u + 1 < D) a[c1] = fma(-u_, r2.y, a[c1]);
                            }
                        }
                    }
                });
            }
        }
        BPHASE_MARK(2);
        // column-major store: for a fixed column the lanes write consecutive rows
        if (row < d) {
            double* q = a_.bs.qmat + (size_t)b * d * d + row;
#pragma unroll
            for (int j = 0; j < D; ++j)
                if (j < d) q[(size_t)j * d] = a[j];
        }
        BPHASE_MARK(3);
    }
}

// ------------------------------------------------------------------------------------------------
// Replay + evolution + gradient, rows of V in registers
// ------------------------------------------------------------------------------------------------
template <int D, int NW>
struct RpLayout {
    static constexpr int NT = NW * 32;
    static constexpr int DP = (D + 1) & ~1;
    static constexpr int KP = (D + 3) & ~3;                  // eigen-index range padded to the DMMA k step
    static constexpr int LDV = KP | 1;                       // odd: row stores and column reads are conflict-free
    static constexpr int LDZ = KP + ((4 - KP % 16) + 16) % 16;   // = 4 (mod 16): a B fragment load takes its minimum of 2 wavefronts
    static constexpr int MT = (D + 7) / 8;                   // DMMA row tiles
    static constexpr int MTW = (MT + NW - 1) / NW;           // ... per warp (tile mt belongs to warp mt % NW)
    static constexpr int HH = D / 2;                         // cyclic offsets per eigen-index: pair {i, (i + o) mod D}, o = 1 .. HH
    // Storage / evaluation order of the pair kernel S.  Cyclic ([i][o], partner (i + o) mod D): no gathers in the pair
    // phase, but every register of V stays live through the row sums; packed upper triangle: V registers die row by
    // row, which the register-starved D <= 66 kernels (168 registers) need more than the conflict-free loads.
    static constexpr bool CYC = (QTET_S_CYCLIC != 0) && (D > 66);
    static constexpr int HHP = (HH + 1) & ~1;                // cyclic: row stride of S (even: every row starts on a 16 B boundary)
    static constexpr int NP = CYC ? D * HHP : D * (D - 1) / 2;
    // always live (doubles)
    static constexpr int VEC = 0;                            // [8][DP]  e, c, fr, fi, ar, ai, br, bi
    static constexpr int WV = VEC + 8 * DP;                  // double2 [NT]   n_site conj(psi(t*))
    static constexpr int RED = WV + 2 * NT;                  // [64]
    static constexpr int PHASED = RED + 64;
    // replay phase
    static constexpr int RING = 0;                           // double2 [kRingR][kFuse][DP]   (c, s) by plane index
    static constexpr int R1 = kRingR * kFuse * DP * 2;       // followed by uint2 hdr[maxr] (runtime size)
    // evolution phase
#if QTET_EVO_MMA
    static constexpr int ZM = (D * LDV + 1) & ~1;                     // behind V (VSM below): double [kNC][LDZ], column n = (sample, re|im)
    static constexpr int PARTM = ZM + kNC * LDZ;             // [NW][kNC / 2]
    static constexpr int R2 = PARTM + NW * (kNC / 2);
    static constexpr int ZT = ZM;                            // psi(t*) broadcast buffer (double2 [DP]) once the evolution is over
    static_assert(kNC * LDZ >= 2 * DP, "psi* buffer must fit the Z chunk");
    static_assert(NT >= KP, "every padded eigen-index needs a thread");
#else
    static constexpr int ZT = 0;                             // double2 [kTC][DP]
    static constexpr int PART = ZT + kTC * DP * 2;           // [kTC][NT + 1]
    static constexpr int R2 = PART + kTC * (NT + 1);
#endif
    // gradient phase
    static constexpr int VSM = 0;                            // [D][LDV], then S: [DP diag][NP pairs]
    static constexpr int R3a = D * LDV;
    static constexpr int R3b = DP + ((NP + 1) & ~1);
    static constexpr int R3 = R3a > R3b ? R3a : R3b;
    static constexpr int R23 = R2 > R3 ? R2 : R3;
};

// Register budget for MINB CTAs of NW warps per SM.  (`__launch_bounds__(threads, MINB)` rounds a 3-warp CTA up to four
// warps when it derives the cap -- 168 instead of 224 registers for three CTAs of 96 threads -- so the cap is given explicitly.)
constexpr int reg_cap(int nw, int minb) {
    int r = 65536 / (minb * nw * 32);
    r = r / 8 * 8;
    return r > 255 ? 255 : r;
}

// ------------------------------------------------------------------------------------------------
// DIRECT variant: eigenvectors of the tridiagonal straight from its eigenvalues (no rotation stream, no replay).
// Thread i runs the three-term recurrence of (T - lam_i) z = 0 from the bottom and from the top and joins the two at
// the row r that maximises |p_r q_r| (see qtet_direct.cu).  At d ~ 100 the recurrences leave the FP64 range (growth up to
// ||T|| / b per row), so the running pair is rescaled by 2^-256 whenever it exceeds 2^200; the rescale events of the
// bottom-up pass are kept in a 128-bit mask, the top-down pass is deterministic and simply replays its own.
// ------------------------------------------------------------------------------------------------
constexpr double kGapGS = 2.44140625e-4;                // 2^-12: re-orthogonalise neighbours closer than this (x ||T||)
constexpr double kGapFB = 5.820766091346741e-11;        // 2^-34: numerically degenerate -> generic kernel
constexpr double kRecBig = 1.6069380442589903e60;       // 2^200
constexpr double kRecF = 8.636168555094445e-78;         // 2^-256
constexpr double kCsigBig = 1e-14;                      // eigenvectors with |c_i| below this are left out of the time evolution
constexpr double kCsigGrad = 1e-18;                     // ... and index pairs whose c_i, c_j are BOTH below this out of the gradient kernel
constexpr int kWmax = 16;                               // widest window the strip form of the gradient handles

__device__ __forceinline__ double rsqrt_fast_d(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y * y, 1.0);
    return fma(y * e, fma(e, 0.375, 0.5), y);
}
__device__ __forceinline__ double rec_scale(int m) {      // (2^-256)^m, 0 from m = 4 on (2^-1024 is not a double)
    return m <= 0 ? 1.0 : (m == 1 ? kRecF : (m == 2 ? kRecF * kRecF : (m == 3 ? kRecF * kRecF * kRecF : 0.0)));
}
__device__ __forceinline__ int rec_key(double sp, double sq, int e) {      // ~ log2 |p q| of the TRUE (unscaled) values
    const double pr = fabs(sp * sq);
    return ((__double2hiint(pr) >> 20) & 0x7ff) + 256 * e;
}

// Column `i` of Z (eigenvector of the tridiagonal (sa, sb) for eigenvalue lam), normalised, into zc[k * ld], k < d.
// Returns false when the recurrences produced something unusable (NaN / overflow): the point then takes the fallback.
// Scale epochs: a rescale of the bottom-up pass at entry k applies to the entries <= k; the (few) positions are kept as a stack of
// 7-bit fields in one 64-bit word, smallest position in the low bits, so that the top-down passes meet them in order with one
// compare per step; the factors that bring an entry to the scale of the join row change only at those events.
__device__ __forceinline__ bool direct_column(int d, double lam, const double* __restrict__ sa, const double* __restrict__ sb,
                                              const double* __restrict__ sib, double* zc, int ld) {
    unsigned long long qev = ~0ull;                      // event stack (127 = none)
    int eq0 = 0;                                         // scale epoch of q_0 = number of events
    // from the bottom: q_{k-1} = ((lam - a_k) q_k - b_k q_{k+1}) / b_{k-1},  q_{d-1} = 1
    {
        double qn = 0.0, qc = 1.0;
        zc[(size_t)(d - 1) * ld] = 1.0;
#pragma unroll 4
        for (int k = d - 1; k >= 1; --k) {
            double val = fma(lam - sa[k], qc, -(sb[k] * qn)) * sib[k - 1];
            qn = qc; qc = val;
            if (fabs(qc) > kRecBig) {
                qc *= kRecF; qn *= kRecF;
                qev = (qev << 7) | (unsigned long long)(k - 1);
                ++eq0;
            }
            zc[(size_t)(k - 1) * ld] = qc;
        }
    }
    if (eq0 > 8) return false;                           // (cannot happen below 2^2048 of growth; the stack holds nine fields)
    constexpr unsigned long long kRefill = 127ull << 57;
    // from the top: p_{k+1} = ((lam - a_k) p_k - b_{k-1} p_{k-1}) / b_k,  p_0 = 1; join row r = argmax |p_k q_k|
    int r = 0, epr = 0, eqr = eq0;
    double pr = 1.0, qr = zc[0];
    {
        double pm = 0.0, pc = 1.0;
        int ep = 0, eq = eq0;
        unsigned long long qs = qev;
        int nextq = (int)(qs & 127ull);
        int best = rec_key(1.0, qr, eq0);
#pragma unroll 4
        for (int k = 0; k < d - 1; ++k) {
            if (k == nextq) { --eq; qs = (qs >> 7) | kRefill; nextq = (int)(qs & 127ull); }      // epoch of q_{k+1}
            const double bm = (k >= 1) ? sb[k - 1] : 0.0;
            const double pn = fma(lam - sa[k], pc, -(bm * pm)) * sib[k];
            pm = pc; pc = pn;
            if (fabs(pc) > kRecBig) { pc *= kRecF; pm *= kRecF; ++ep; }
            const double sq = zc[(size_t)(k + 1) * ld];
            const int ky = rec_key(pc, sq, ep + eq);
            if (ky > best) { best = ky; r = k + 1; pr = pc; qr = sq; epr = ep; eqr = eq; }
        }
    }
    // z_k = p_k / p_r above the join (the recurrence once more, up to the join only), q_k / q_r from the join down
    double ss = 0.0;
    {
        const double invp = fast_rcp(pr);
        double pm = 0.0, pc = 1.0, fp = rec_scale(epr) * invp;
        int ep = 0;
#pragma unroll 4
        for (int k = 0; k < r; ++k) {
            const double x = pc * fp;
            zc[(size_t)k * ld] = x;
            ss = fma(x, x, ss);
            const double bm = (k >= 1) ? sb[k - 1] : 0.0;
            const double pn = fma(lam - sa[k], pc, -(bm * pm)) * sib[k];
            pm = pc; pc = pn;
            if (fabs(pc) > kRecBig) { pc *= kRecF; pm *= kRecF; ++ep; fp = rec_scale(epr - ep) * invp; }
        }
        const double invq = fast_rcp(qr);
        unsigned long long qs = qev;
        int nextq = (int)(qs & 127ull);
        while (nextq < r) { qs = (qs >> 7) | kRefill; nextq = (int)(qs & 127ull); }    // events above the join row are behind us
        int eq = eqr;
        double fq = invq;
#pragma unroll 4
        for (int k = r; k < d; ++k) {
            const double x = zc[(size_t)k * ld] * fq;
            zc[(size_t)k * ld] = x;
            ss = fma(x, x, ss);
            if (k == nextq) { --eq; qs = (qs >> 7) | kRefill; nextq = (int)(qs & 127ull); fq = rec_scale(eqr - eq) * invq; }
        }
    }
    const bool ok = (ss >= 0.5 && ss < 1e300);           // z_r = 1: |z|^2 >= 1 unless NaN / overflow
    const double rn = ok ? rsqrt_fast_d(ss) : 0.0;
#pragma unroll 4
    for (int k = 0; k < d; ++k) zc[(size_t)k * ld] *= rn;
    return ok;
}

template <int D, int NW, int MINB, int GS, bool DIRECT = false>
__global__ void __launch_bounds__(NW * 32) __maxnreg__(reg_cap(NW, MINB)) replay_reg_kernel(RegArgs a) {
    using L = RpLayout<D, NW>;
    constexpr int NT = L::NT, DP = L::DP, LDV = L::LDV, NP = L::NP;
    constexpr int NG = (D - 1 + 2 * (kFuse - 1) + GS - 1) / GS;     // plane groups of GS steps per trip
    extern __shared__ __align__(16) double sm[];
    const ProblemView& pv = a.pv;
    const int d = pv.d, f = pv.f, T = pv.T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = tid;
    const bool dense = !pv.tridiag;

    double* vec = sm + L::VEC;
    double2* wv = reinterpret_cast<double2*>(sm + L::WV);
    double* red = sm + L::RED;
    int* ired = reinterpret_cast<int*>(red + 48);
    double* ph = sm + L::PHASED;
    double2* ring = reinterpret_cast<double2*>(ph + L::RING);
    uint2* hs = reinterpret_cast<uint2*>(ph + L::R1);
    int2* tw = reinterpret_cast<int2*>(ph + L::R1 + a.bs.maxr);     // per-trip windows, behind the headers
    const unsigned ring_a = smem_addr(ring);
    double2* zt = reinterpret_cast<double2*>(ph + L::ZT);
#if QTET_EVO_MMA
    double* zm = ph + L::ZM;
    double* partm = ph + L::PARTM;
#else
    double* part = ph + L::PART;
#endif
    double* vsm = ph + L::VSM;
    double* sp = ph + L::VSM;
    // tr[T] sits behind the largest phase region (runtime offset handed in through the launch: bs.maxr bounds hs)
    const int r1 = L::R1 + 2 * a.bs.maxr;
    const int phased = (r1 > L::R23 ? r1 : L::R23);
    double* tr = ph + ((phased + 1) & ~1);

    const double wsite = (row < d) ? pv.occ[row * f + a.site] : 0.0;
    const double dt = (T > 1) ? pv.tgrid[1] : 0.0;

    for (long long b = blockIdx.x; b < a.n; b += gridDim.x) {
        BPHASE_DECL();
        double v[D];
        if constexpr (DIRECT) {
            __syncthreads();                                 // previous point is done with every shared region
            BPHASE_START();
            constexpr int KP = L::KP;
            double* zs = vsm;                                // [KP rows k][LDV]: column i = eigenvector i of the tridiagonal
            double* sa = vec;                                // [DP] shifted diagonal   (vec is rewritten later)
            double* sb = vec + DP;                           // [DP] b_k = T[k+1,k]
            double* sib = vec + 2 * DP;                      // [DP] 1 / b_k
            unsigned* gsm = reinterpret_cast<unsigned*>(ired + 4);      // [4] bit i: eigenvalue i is close to eigenvalue i-1
            const double* adg = a.bs.adiag + (size_t)b * d;
            const double* evp = a.bs.ev + (size_t)b * d;
            if (tid < 8) ired[tid] = 0;
            for (int k = tid; k < DP; k += NT) {
                const double off = (k < d - 1) ? (dense ? a.bs.tri_e[(size_t)b * d + k] : pv.offd[k]) : 0.0;
                sa[k] = (k < d) ? adg[k] : 0.0;
                sb[k] = off;
                sib[k] = (k < d - 1) ? 1.0 / off : 0.0;
            }
            // padding: rows d .. KP-1 (the V = Q Z product reads whole k tiles) and columns d .. LDV-1 hold zeros
            for (int idx = tid; idx < (KP - d) * LDV; idx += NT) zs[(size_t)d * LDV + idx] = 0.0;
            if (row >= d && row < LDV) {
                for (int k = 0; k < d; ++k) zs[(size_t)k * LDV + row] = 0.0;
            }
            __syncthreads();
            if (row < d) {
                const double lam = evp[row];
                const double nrm = fmax(fabs(evp[0]), fabs(evp[d - 1]));
                bool fb = !direct_column(d, lam, sa, sb, sib, zs + row, LDV);
                if (row >= 1) {
                    const double gap = lam - evp[row - 1];
                    if (!(gap > kGapFB * nrm)) fb = true;            // also catches NaN (the QL did not converge)
                    else if (gap < kGapGS * nrm) atomicOr(&gsm[row >> 5], 1u << (row & 31));
                }
                if (fb) ired[2] = 1;
            }
            __syncthreads();
            BPHASE_MARK(4);
            if (gsm[0] | gsm[1] | gsm[2] | gsm[3]) {
                // modified Gram-Schmidt inside clusters of close eigenvalues (rare): one warp, lanes share the rows
                if (warp == 0) {
                    for (int i = 1; i < d; ++i) {
                        if (!((gsm[i >> 5] >> (i & 31)) & 1u)) continue;
                        for (int j = i - 1; j >= 0; --j) {
                            double dot = 0.0;
                            for (int k = lane; k < d; k += 32) dot = fma(zs[(size_t)k * LDV + i], zs[(size_t)k * LDV + j], dot);
                            dot = warp_sum(dot);
                            for (int k = lane; k < d; k += 32) zs[(size_t)k * LDV + i] = fma(-dot, zs[(size_t)k * LDV + j], zs[(size_t)k * LDV + i]);
                            __syncwarp();
                            if (!((gsm[j >> 5] >> (j & 31)) & 1u)) break;      // j was the first member of the cluster
                        }
                        double s2 = 0.0;
                        for (int k = lane; k < d; k += 32) s2 = fma(zs[(size_t)k * LDV + i], zs[(size_t)k * LDV + i], s2);
                        s2 = warp_sum(s2);
                        if (!(s2 > 0.25)) { if (lane == 0) ired[2] = 1; s2 = 1.0; }   // the recurrences returned (nearly) the same vector
                        const double rn = 1.0 / sqrt(s2);
                        for (int k = lane; k < d; k += 32) zs[(size_t)k * LDV + i] *= rn;
                        __syncwarp();
                    }
                }
                __syncthreads();
            }
            if (tid == 0 && ired[2]) {                       // numerically degenerate / overflowed: the generic kernel redoes the point
                const int slot = atomicAdd(a.fb_count, 1);
                a.fb_list[slot] = (int)b;
                atomicAdd(pv.stats + 0, 1ULL);
            }
            if (dense) {
                // ---- V = Q Z on the FP64 tensor path: warp w owns the row tiles w, w + NW, ...; A = Q fragments straight from
                // global memory (column-major: 8 consecutive rows of one column are 64 contiguous bytes), B = Z from shared
                // memory.  The C tiles stay in registers until every warp is done reading Z, then replace it.
                if constexpr (D <= 84) {
                    constexpr int MTW = L::MTW, KT = KP / 4, NT8 = (D + 7) / 8;
                    const int lg = lane >> 2, lt = lane & 3;
                    const double* qb = a.bs.qmat + (size_t)b * d * d;
                    const unsigned zaddr = smem_addr(zs) + (unsigned)((lt * LDV + lg) * 8);
                    double2 acc[MTW][NT8];
#pragma unroll
                    for (int m = 0; m < MTW; ++m)
#pragma unroll
                        for (int n = 0; n < NT8; ++n) acc[m][n] = make_double2(0.0, 0.0);
#pragma unroll
                    for (int m = 0; m < MTW; ++m) {
                        const int r0 = 8 * (warp + m * NW);
                        if (r0 < d) {
                            const int r = r0 + lg;
                            double af[KT];
#pragma unroll
                            for (int kt = 0; kt < KT; ++kt) {
                                const int k = 4 * kt + lt;
                                af[kt] = (r < d && k < d) ? qb[(size_t)k * d + r] : 0.0;
                            }
#pragma unroll
                            for (int kt = 0; kt < KT; ++kt)
#pragma unroll
                                for (int n = 0; n < NT8; ++n)
                                    dmma884(acc[m][n], af[kt], lds64(zaddr + (unsigned)((4 * kt * LDV + 8 * n) * 8)));
                        }
                    }
                    __syncthreads();                         // every warp has read Z
#pragma unroll
                    for (int m = 0; m < MTW; ++m) {
                        const int r = 8 * (warp + m * NW) + lg;
                        if (r < d) {
#pragma unroll
                            for (int n = 0; n < NT8; ++n) {
                                const int c0 = 8 * n + 2 * lt;
                                if (c0 < d) zs[(size_t)r * LDV + c0] = acc[m][n].x;
                                if (c0 + 1 < d) zs[(size_t)r * LDV + c0 + 1] = acc[m][n].y;
                            }
                        }
                    }
                    __syncthreads();
                }
            }
#pragma unroll
            for (int j = 0; j < D; ++j) v[j] = (row < d) ? zs[(size_t)row * LDV + j] : 0.0;
            __syncthreads();                                 // sa / sb / sib (in vec) and Z are dead from here on
            BPHASE_MARK(5);
        } else {
        const long long batch = b >> 5;
        const int blane = (int)(b & 31);
        const int nr_all = a.bs.nrounds[batch];
        const double2* rec = a.bs.stream + (size_t)batch * a.bs.capit * 32 + blane;
        const unsigned* hdr = a.bs.hdr + (size_t)batch * a.bs.maxr * 32 + blane;
        const int* it0s = a.bs.it0 + (size_t)batch * a.bs.maxr;
        __syncthreads();                                 // previous point is done with every shared region
        BPHASE_START();
        if (tid == 0) ired[0] = 0;
        // ---- V <- I or Q (column-major: coalesced over the rows)
        if (dense && row < d) {
            const double* q = a.bs.qmat + (size_t)b * d * d + row;
#pragma unroll
            for (int j = 0; j < D; ++j) v[j] = (j < d) ? q[(size_t)j * d] : 0.0;
        } else {
#pragma unroll
            for (int j = 0; j < D; ++j) v[j] = (j == row) ? 1.0 : 0.0;
        }
        __syncthreads();
        // ---- headers of my point; rounds after my point's last active one are skipped
        {
            int last = 0;
            for (int r = tid; r < nr_all; r += NT) {
                const unsigned h = hdr[(size_t)r * 32];
                hs[r] = make_uint2(h, (unsigned)it0s[r]);
                if ((h & 0xff) != 0xff) last = r + 1;
            }
            last = __reduce_max_sync(0xffffffffu, last);
            if (lane == 0 && last > 0) atomicMax(&ired[0], last);
        }
        __syncthreads();
        const int nr = ired[0];
        BPHASE_MARK(4);
        // ---- replay.  kFuse consecutive rounds are swept TOGETHER: round f of a trip trails round f-1 by two
        // planes, so the trip carries kFuse independent dependency chains per row and costs one barrier.  A ring slot
        // holds the (c, s) of those rounds indexed by PLANE; every plane entry is rewritten for every trip (identity
        // outside a round's window), so plane groups may overhang the windows on both sides.
        const int ntrip = (nr + kFuse - 1) / kFuse;
        // window of every trip in units of the LEADER's plane index (round f works on plane i + 2 f at step i)
        for (int s_ = tid; s_ < ntrip; s_ += NT) {
            int ilo = 1 << 20, ihi = -(1 << 20);
#pragma unroll
            for (int f_ = 0; f_ < kFuse; ++f_) {
                const int r = s_ * kFuse + f_;
                if (r < nr) {
                    const uint2 h = hs[r];
                    const int lp = h.x & 0xff, mine = h.x >> 24;     // my point sweeps planes [lp, mine)
                    if (lp != 0xff) { ilo = min(ilo, lp - 2 * f_); ihi = max(ihi, mine - 1 - 2 * f_); }
                }
            }
            tw[s_] = make_int2(ilo, ihi);
        }
        __syncthreads();
        // (c, s) of trip s_: plane entry `tid` of every fused round, copied global -> shared (identity outside the window)
        auto issue = [&](int s_, const uint2 (&hh)[kFuse]) {
            if (s_ < ntrip && tid < D - 1) {
#pragma unroll
                for (int f_ = 0; f_ < kFuse; ++f_) {
                    double2* dst = ring + ((s_ % kRingR) * kFuse + f_) * DP + tid;
                    const int lp = hh[f_].x & 0xff, mtop = (hh[f_].x >> 8) & 0xff, mine = hh[f_].x >> 24;
                    if (lp != 0xff && tid >= lp && tid < mine) cp_async16(dst, rec + (size_t)((int)hh[f_].y + (mtop - 1 - tid)) * 32);
                    else *dst = make_double2(1.0, 0.0);
                }
            }
            cp_async_commit();
        };
        auto headers = [&](int s_, uint2 (&hh)[kFuse]) {
#pragma unroll
            for (int f_ = 0; f_ < kFuse; ++f_) {
                const int r = s_ * kFuse + f_;
                hh[f_] = (r < nr) ? hs[r] : make_uint2(0xffu, 0u);
            }
        };
#pragma unroll
        for (int s_ = 0; s_ < kRingR - 1; ++s_) { uint2 hh[kFuse]; headers(s_, hh); issue(s_, hh); }
        int2 wnext = (ntrip > 0) ? tw[0] : make_int2(0, -1);
        for (int s_ = 0; s_ < ntrip; ++s_) {
            cp_async_wait<kRingR - 2>();                 // my copies of trip s_ have landed
            __syncthreads();                             // ... and everyone else's; trip s_-1 is fully consumed
            uint2 hn[kFuse];
            headers(s_ + kRingR - 1, hn);                // loaded now, used after the sweep
            const int2 w = wnext;
            wnext = tw[(s_ + 1 < ntrip) ? s_ + 1 : s_];
            if (w.y >= w.x) {
                const int ilo = w.x;
                const unsigned ca = ring_a + (unsigned)((s_ % kRingR) * kFuse * DP * 16);
                // plane groups top-down from plane D-2 (the windows always reach the top; identity entries cover the
                // exceptions), the next group's (c, s) fetched while this one is applied, one uniform exit test per group
                double2 cur[GS][kFuse], nxt[GS][kFuse];
                auto fetch = [&](double2 (&dst)[GS][kFuse], auto Gc) {
                    constexpr int i0 = D - 2 - GS * decltype(Gc)::value;
#pragma unroll
                    for (int k4 = 0; k4 < GS; ++k4)
#pragma unroll
                        for (int f_ = 0; f_ < kFuse; ++f_) {
                            const int pl = i0 - k4 + 2 * f_;
                            if (pl >= 0 && pl <= D - 2) dst[k4][f_] = lds128(ca + (unsigned)((f_ * DP + (pl >= 0 ? pl : 0)) * 16));
                        }
                };
                fetch(cur, std::integral_constant<int, 0>{});
                bool live = true;
                static_for<NG>([&](auto Gc) {
                    constexpr int g_ = decltype(Gc)::value;
                    constexpr int i0 = D - 2 - GS * g_;
                    if (live && i0 < ilo) live = false;
                    if (live) {
                        if constexpr (g_ + 1 < NG) fetch(nxt, std::integral_constant<int, g_ + 1>{});
#pragma unroll
                        for (int k4 = 0; k4 < GS; ++k4)
#pragma unroll
                            for (int f_ = 0; f_ < kFuse; ++f_) {
                                const int pl = i0 - k4 + 2 * f_;
                                if (pl >= 0 && pl <= D - 2) {
                                    const int pc = (pl >= 0 && pl <= D - 2) ? pl : 0;
                                    const double x = v[pc], y = v[pc + 1];
                                    const double t1 = cur[k4][f_].x * x, t2 = cur[k4][f_].y * x;
                                    v[pc] = fma(-cur[k4][f_].y, y, t1);
                                    v[pc + 1] = fma(cur[k4][f_].x, y, t2);
                                }
                            }
#pragma unroll
                        for (int k4 = 0; k4 < GS; ++k4)
#pragma unroll
                            for (int f_ = 0; f_ < kFuse; ++f_) cur[k4][f_] = nxt[k4][f_];
                    }
                });
            }
            issue(s_ + kRingR - 1, hn);
        }
        cp_async_wait<0>();
        __syncthreads();                                 // replay region is dead from here on
        BPHASE_MARK(5);

        }

        // ---- c = V[0,:], eigenvalues, unit phase steps
        if (row == 0) {
#pragma unroll
            for (int i = 0; i < D; ++i) vec[DP + i] = v[i];
            ired[8] = 1 << 20; ired[9] = -1;             // window of the eigen-indices that carry weight (below)
            ired[10] = 1 << 20; ired[11] = -1;           // ... and the wider one the gradient uses
        }
        __syncthreads();
        double ei = 0.0, ci = 0.0, wre = 1.0, wim = 0.0, zr = 0.0, zi = 0.0;
        if (row < D) {
            ci = (row < d) ? vec[DP + row] : 0.0;
            // c_i = V[0][i] decays fast away from the initial state's energy: eigen-indices with |c_i| <= 1e-14 cannot move <n(t)>
            // by more than 2 N d 1e-14, so the evolution only runs over the window [lo, hi] that holds the others (dimer
            // N=100: 10 of 101 columns on average, trimer N=10: 55 of 66)
            if (fabs(ci) > kCsigBig) { atomicMin(&ired[8], row); atomicMax(&ired[9], row); }
            if (fabs(ci) > kCsigGrad) { atomicMin(&ired[10], row); atomicMax(&ired[11], row); }
            ei = (row < d) ? a.bs.ev[(size_t)b * d + row] : 0.0;
            const double th = ei * dt;
            const double lo = fma(ei, dt, -th);          // exact product tail
            double sn, cs_;
            sincos(th, &sn, &cs_);
            wre = fma(-sn, lo, cs_);                     // exp(-i (th + lo))
            wim = -fma(cs_, lo, sn);
            zr = ci;
        }
#if QTET_EVO_MMA
        // ---- time evolution on the FP64 tensor path: P[row][n] = sum_i V[row][i] Z[i][n], n = (sample, re|im), as
        // m8n8k4 DMMAs with V (parked in shared memory once -- the gradient needs it there anyway) as the A operand and
        // the phase table Z of a chunk of 8 samples as B.  A C fragment holds (re, im) of psi_row(t_s) in one lane,
        // so |psi|^2 is lane-local; the weighted sum over rows is three shuffles and one partial per warp.
        {
            constexpr int KP = L::KP, LDZ = L::LDZ, MTW = L::MTW;
            if (row < d) {
#pragma unroll
                for (int i = 0; i < D; ++i) vsm[(size_t)row * LDV + i] = v[i];
#pragma unroll
                for (int i = D; i < KP; ++i) vsm[(size_t)row * LDV + i] = 0.0;
            }
            // (v is reloaded from shared memory after the evolution, so its registers are free for the DMMA loop)
            const int lg = lane >> 2, lt = lane & 3;
            unsigned aaddr[MTW];                             // A fragment base of my row tiles (rows >= d: clamped, weight 0)
            double wrow[MTW];
            int nmt = 0;                                     // my tiles are warp, warp + NW, ...: the valid ones are a prefix
#pragma unroll
            for (int m = 0; m < MTW; ++m) {
                const int r = 8 * (warp + m * NW) + lg;
                const int rc = (r < d) ? r : (d - 1);
                aaddr[m] = smem_addr(vsm) + (unsigned)((rc * LDV + lt) * 8);
                wrow[m] = (r < d) ? pv.occ[r * f + a.site] : 0.0;
                if (8 * (warp + m * NW) < d) nmt = m + 1;
            }
            const unsigned baddr = smem_addr(zm) + (unsigned)((lg * LDZ + lt) * 8);
            for (int k0 = 0; k0 < T; k0 += kNC / 2) {
                __syncthreads();                             // V is in place / the previous chunk is consumed
                if (row < KP) {
#pragma unroll
                    for (int kk = 0; kk < kNC / 2; ++kk) {
                        zm[(size_t)(2 * kk) * LDZ + row] = zr;
                        zm[(size_t)(2 * kk + 1) * LDZ + row] = zi;
                        const double nr_ = fma(zr, wre, -(zi * wim));
                        zi = fma(zr, wim, zi * wre);
                        zr = nr_;
                    }
                }
                __syncthreads();
                const int ks_lo = ired[8] & ~3, ks_hi = (ired[9] < 0) ? 0 : (((ired[9] & ~3) + 4 < KP) ? (ired[9] & ~3) + 4 : KP);
                double2 acc[MTW][kNC / 8];
#pragma unroll
                for (int m = 0; m < MTW; ++m)
#pragma unroll
                    for (int n = 0; n < kNC / 8; ++n) acc[m][n] = make_double2(0.0, 0.0);
                // the tile count is warp-uniform: one branch-free k loop per possible count
                static_for<MTW>([&](auto Cc) {
                    constexpr int NMT = decltype(Cc)::value + 1;
                    if (nmt == NMT) {
#pragma unroll 4
                        for (int ks = ks_lo; ks < ks_hi; ks += 4) {
                            double bf[kNC / 8], af[NMT];
#pragma unroll
                            for (int n = 0; n < kNC / 8; ++n) bf[n] = lds64(baddr + (unsigned)((8 * n * LDZ + ks) * 8));
#pragma unroll
                            for (int m = 0; m < NMT; ++m) af[m] = lds64(aaddr[m] + (unsigned)(ks * 8));
#pragma unroll
                            for (int m = 0; m < NMT; ++m)
#pragma unroll
                                for (int n = 0; n < kNC / 8; ++n) dmma884(acc[m][n], af[m], bf[n]);
                        }
                    }
                });
#pragma unroll
                for (int n = 0; n < kNC / 8; ++n) {
                    double sum = 0.0;
#pragma unroll
                    for (int m = 0; m < MTW; ++m) sum = fma(wrow[m], fma(acc[m][n].x, acc[m][n].x, acc[m][n].y * acc[m][n].y), sum);
                    sum += __shfl_xor_sync(0xffffffffu, sum, 4);
                    sum += __shfl_xor_sync(0xffffffffu, sum, 8);
                    sum += __shfl_xor_sync(0xffffffffu, sum, 16);
                    if (lane < 4) partm[warp * (kNC / 2) + 4 * n + lane] = sum;      // sample 4 n + lane of the chunk
                }
                __syncthreads();
                if (tid < kNC / 2 && k0 + tid < T) {
                    double s0 = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) s0 += partm[w * (kNC / 2) + tid];
                    tr[k0 + tid] = s0;
                }
            }
            // my row of V back into registers (rows >= d are padding: zero)
#pragma unroll
            for (int i = 0; i < D; ++i) v[i] = (row < d) ? vsm[(size_t)row * LDV + i] : 0.0;
        }
#else
        // ---- time evolution in chunks of kTC samples
        for (int k0 = 0; k0 < T; k0 += kTC) {
            if (row < D) {
#pragma unroll 4
                for (int kk = 0; kk < kTC; ++kk) {
                    zt[kk * DP + row] = make_double2(zr, zi);
                    const double nr_ = fma(zr, wre, -(zi * wim));
                    zi = fma(zr, wim, zi * wre);
                    zr = nr_;
                }
            }
            __syncthreads();
            const int kend = (T - k0 < kTC) ? (T - k0) : kTC;
#pragma unroll 1
            for (int kk = 0; kk < kend; kk += 2) {
                const double2* z0 = zt + kk * DP;
                const double2* z1 = z0 + DP;
                double pr0 = 0.0, pi0 = 0.0, pr1 = 0.0, pi1 = 0.0;
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    const double2 za = z0[i], zb = z1[i];
                    pr0 = fma(v[i], za.x, pr0); pi0 = fma(v[i], za.y, pi0);
                    pr1 = fma(v[i], zb.x, pr1); pi1 = fma(v[i], zb.y, pi1);
                }
                part[kk * (NT + 1) + row] = wsite * fma(pr0, pr0, pi0 * pi0);
                part[(kk + 1) * (NT + 1) + row] = wsite * fma(pr1, pr1, pi1 * pi1);
            }
            __syncthreads();
            if (tid < kTC && k0 + tid < T) {
                const double* pp = part + tid * (NT + 1);
                double s0 = 0.0, s1 = 0.0;
                int j = 0;
                for (; j + 2 <= d; j += 2) { s0 += pp[j]; s1 += pp[j + 1]; }
                if (j < d) s0 += pp[j];
                tr[k0 + tid] = s0 + s1;
            }
        }
#endif
        __syncthreads();
        BPHASE_MARK(6);
        if (a.trace != nullptr)
            for (int k = tid; k < T; k += NT) a.trace[b * T + k] = tr[k];
        if (warp == 0) {
            int best = -1;
            double bv = 0.0;
            for (int k = lane; k < T; k += 32) {
                const double key = a.acceptor ? tr[k] : -tr[k];
                if (best < 0 || key > bv) { bv = key; best = k; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int ob = __shfl_xor_sync(0xffffffffu, best, o);
                if (ob >= 0 && (best < 0 || ov > bv || (ov == bv && ob < best))) { bv = ov; best = ob; }
            }
            if (lane == 0) {
                ired[1] = best;
                if (a.loss) a.loss[b] = a.acceptor ? ((double)pv.N - bv) : -bv;
                if (a.tstar) a.tstar[b] = best;
            }
        }
        __syncthreads();
        if (a.grad == nullptr) continue;

        // ---- gradient at t* : f_i = w_i^kbest, z*_i = c_i f_i
        const int kbest = ired[1];
        const double ts = pv.tgrid[kbest];
        double fr = 1.0, fi = 0.0;
        {
            double br = wre, bi = wim;
            for (int e = kbest; e > 0; e >>= 1) {
                if (e & 1) { const double tt = fma(fr, br, -(fi * bi)); fi = fma(fr, bi, fi * br); fr = tt; }
                const double t2 = fma(br, br, -(bi * bi)); bi = 2.0 * br * bi; br = t2;
            }
        }
        if (row < D) zt[row] = make_double2(ci * fr, ci * fi);
        __syncthreads();
        {
            double pr0 = 0.0, pi0 = 0.0, pr1 = 0.0, pi1 = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const double2 z = zt[i];
                if (i & 1) { pr1 = fma(v[i], z.x, pr1); pi1 = fma(v[i], z.y, pi1); }
                else { pr0 = fma(v[i], z.x, pr0); pi0 = fma(v[i], z.y, pi0); }
            }
            wv[row] = make_double2(wsite * (pr0 + pr1), -wsite * (pi0 + pi1));
        }
#if !QTET_EVO_MMA
        __syncthreads();                                 // zt is dead: V goes to shared memory for the column sums
        if (row < d) {
#pragma unroll
            for (int i = 0; i < D; ++i) vsm[(size_t)row * LDV + i] = v[i];
        }
#endif
        __syncthreads();
        double ar_ = 0.0, ai_ = 0.0;                       // a_row = sum_j V[j][row] w_j, kept for the pair kernel
        if (row < D) {
            double ar = 0.0, ai = 0.0;
            if (row < d) {
                double ar1 = 0.0, ai1 = 0.0;
                int j = 0;
                for (; j + 2 <= d; j += 2) {
                    const double v0 = vsm[(size_t)j * LDV + row], v1 = vsm[(size_t)(j + 1) * LDV + row];
                    const double2 w0 = wv[j], w1 = wv[j + 1];
                    ar = fma(v0, w0.x, ar); ai = fma(v0, w0.y, ai);
                    ar1 = fma(v1, w1.x, ar1); ai1 = fma(v1, w1.y, ai1);
                }
                if (j < d) { const double v0 = vsm[(size_t)j * LDV + row]; const double2 w0 = wv[j]; ar = fma(v0, w0.x, ar); ai = fma(v0, w0.y, ai); }
                ar += ar1; ai += ai1;
            }
            const double bre = fma(ar, fr, -(ai * fi)), bim = fma(ar, fi, ai * fr);
            ar_ = ar; ai_ = ai;
            vec[0 * DP + row] = ei; vec[1 * DP + row] = ci; vec[2 * DP + row] = fr; vec[3 * DP + row] = fi;
            vec[4 * DP + row] = ar; vec[5 * DP + row] = ai; vec[6 * DP + row] = bre; vec[7 * DP + row] = bim;
        }
        __syncthreads();                                 // V in shared memory is dead: the kernel S takes its place
        BPHASE_MARK(7);
        // Every term of the gradient kernel carries c_i or c_j (K_ij = c_j (...)), and c decays fast away from the initial state's
        // energy: with W = [wlo, whi] the window of |c_i| > 1e-18, only the pairs with a member in W count -- a D x |W| STRIP
        //   T[i][j] = K_jj (i = j),  S_ij / 2 (i in W),  S_ij (i outside W)          G_row = 2 sum_{j in W} V_j sum_i T[i][j] V_i
        // instead of the D^2 / 2 pairs (dimer N=100: |W| = 10 on average: a fifth of the divisions and of the row-sum work).
        // V stays in shared memory (the strip lives in the dead phase-table region), V_j comes from there.
        const int wlo = ired[10], nW = ired[11] - ired[10] + 1;
        const bool use_strip = L::CYC && nW >= 1 && nW <= kWmax;
        if constexpr (L::CYC) {
            if (use_strip) {
                double* strip = ph + L::ZM;                  // [kWmax][DP]
                if (row < D) {
                    const double br_i = fma(ar_, fr, -(ai_ * fi)), bim_i = vec[7 * DP + row];
                    const bool in_w = (row >= wlo) && (row < wlo + nW);
#pragma unroll 2
                    for (int jw = 0; jw < nW; ++jw) {
                        const int j = wlo + jw;
                        const double dl = ei - vec[j];
                        const double c_j = vec[DP + j], fjr = vec[2 * DP + j], fji = vec[3 * DP + j];
                        const double kij = c_j * (br_i - ar_ * fjr + ai_ * fji);
                        const double kji = ci * (vec[6 * DP + j] - vec[4 * DP + j] * fr + vec[5 * DP + j] * fi);
                        double val;
                        if (j == row) val = ci * ts * bim_i;                       // K_ii = c_i t* Im(b_i)
                        else {
                            if (fabs(dl * ts) < 1e-2) {
                                // near-degenerate: Phi_ij = ts f_j (exp(-i x) - 1)/x ; S = Re[(a_i c_j + a_j c_i) Phi_ij]
                                double gr, gi;
                                expm1_over_x(dl * ts, gr, gi);
                                const double phr = ts * (fjr * gr - fji * gi), phi = ts * (fjr * gi + fji * gr);
                                const double mr = ar_ * c_j + vec[4 * DP + j] * ci;
                                const double mi = ai_ * c_j + vec[5 * DP + j] * ci;
                                val = mr * phr - mi * phi;
                            } else {
                                val = (kij - kji) * fast_rcp(dl);
                            }
                            if (in_w) val *= 0.5;                                 // pairs inside the window appear twice
                        }
                        strip[(size_t)jw * DP + row] = (row < d) ? val : 0.0;
                    }
                }
            }
        }
        if (!use_strip) {
        if (row < D) sp[row] = (row < d) ? ci * ts * vec[7 * DP + row] : 0.0;      // K_ii = c_i t* Im(b_i)
        // S_ij = K_ij + K_ji.  Thread i pairs its own eigen-index (e, c, f, a, b in registers) with j = (i + o) mod D,
        // o = 1 .. D/2: every index pair once, the partner's values come from consecutive addresses across the warp
        // (no pair table, no gather), and S lands in the [i][o] order the row sums below walk through statically.
        // D <= 66 keeps the packed upper triangle dealt out through the pair table (see RpLayout::CYC).
        if constexpr (L::CYC) { if (row < D) {
            constexpr int HH = L::HH, U = 4;
            const int i = row;
            const double br_i = fma(ar_, fr, -(ai_ * fi));
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
                    dl[u] = ei - vec[j];
                    const double kij = vec[DP + j] * (br_i - ar_ * vec[2 * DP + j] + ai_ * vec[3 * DP + j]);
                    const double kji = ci * (vec[6 * DP + j] - vec[4 * DP + j] * fr + vec[5 * DP + j] * fi);
                    kd[u] = kij - kji;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) kd[u] *= fast_rcp(dl[u]);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (o0 + u <= omax) {
                        const int j = jj[u];
                        double sij = kd[u];
                        if (i >= d || j >= d) sij = 0.0;
                        else if (fabs(dl[u] * ts) < 1e-2) {
                            // near-degenerate: Phi_ij = ts f_j (exp(-i x) - 1)/x ; S = Re[(a_i c_j + a_j c_i) Phi_ij]
                            double gr, gi;
                            expm1_over_x(dl[u] * ts, gr, gi);
                            const double fjr = vec[2 * DP + j], fji = vec[3 * DP + j];
                            const double phr = ts * (fjr * gr - fji * gi), phi = ts * (fjr * gi + fji * gr);
                            const double c_j = vec[DP + j];
                            const double mr = ar_ * c_j + vec[4 * DP + j] * ci;
                            const double mi = ai_ * c_j + vec[5 * DP + j] * ci;
                            sij = mr * phr - mi * phi;
                        }
                        sp[DP + i * L::HHP + (o0 + u - 1)] = sij;
                    }
                }
            }
        } } else {
        // S_ij = K_ij + K_ji for i < j
        for (int q0 = tid; q0 < NP; q0 += 2 * NT) {
            int pi_[2], pj_[2];
            double dl[2], kd[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int qq = q0 + u * NT;
                const unsigned ij = a.pair_tab[(qq < NP) ? qq : 0];
                pi_[u] = ij & 0xff; pj_[u] = ij >> 8;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int i = pi_[u], j = pj_[u];
                dl[u] = vec[i] - vec[j];
                const double kij = vec[DP + j] * (vec[6 * DP + i] - vec[4 * DP + i] * vec[2 * DP + j] + vec[5 * DP + i] * vec[3 * DP + j]);
                const double kji = vec[DP + i] * (vec[6 * DP + j] - vec[4 * DP + j] * vec[2 * DP + i] + vec[5 * DP + j] * vec[3 * DP + i]);
                kd[u] = kij - kji;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) kd[u] *= fast_rcp(dl[u]);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int qq = q0 + u * NT;
                if (qq < NP) {
                    const int i = pi_[u], j = pj_[u];
                    double sij = kd[u];
                    if (j >= d) sij = 0.0;
                    else if (fabs(dl[u] * ts) < 1e-2) {
                        // near-degenerate: Phi_ij = ts f_j (exp(-i x) - 1)/x ; S = Re[(a_i c_j + a_j c_i) Phi_ij]
                        double gr, gi;
                        expm1_over_x(dl[u] * ts, gr, gi);
                        const double fjr = vec[2 * DP + j], fji = vec[3 * DP + j];
                        const double phr = ts * (fjr * gr - fji * gi), phi = ts * (fjr * gi + fji * gr);
                        const double c_i = vec[DP + i], c_j = vec[DP + j];
                        const double mr = vec[4 * DP + i] * c_j + vec[4 * DP + j] * c_i;
                        const double mi = vec[5 * DP + i] * c_j + vec[5 * DP + j] * c_i;
                        sij = mr * phr - mi * phi;
                    }
                    sp[DP + qq] = sij;
                }
            }
        }
        }
        }
        __syncthreads();
        BPHASE_MARK(8);
        // G_row = 2 [sum_i K_ii V_i^2 + sum_{i<j} S_ij V_i V_j], every index static
        double Gm = 0.0;
        if (use_strip) {
            if constexpr (L::CYC) {
                const unsigned ts_a = smem_addr(ph + L::ZM);
                const double* vrow = vsm + (size_t)(row < d ? row : 0) * LDV + wlo;
#pragma unroll 1
                for (int jw = 0; jw < nW; ++jw) {
                    const unsigned ta = ts_a + (unsigned)(jw * DP * 8);
                    double acc0 = 0.0, acc1 = 0.0;
                    static_for<(D + 1) / 2>([&](auto Ic) {
                        constexpr int i = 2 * decltype(Ic)::value;
                        const double2 t2 = lds128(ta + i * 8);
                        acc0 = fma(t2.x, v[i], acc0);
                        if constexpr (i + 1 < D) acc1 = fma(t2.y, v[i + 1], acc1);
                    });
                    Gm = fma(vrow[jw], acc0 + acc1, Gm);
                }
            }
        } else {
            const double2* sp2 = reinterpret_cast<const double2*>(sp + DP);
            double2 cur = make_double2(0.0, 0.0);
            if constexpr (L::CYC) {
            static_for<D>([&](auto Ic) {
                constexpr int i = decltype(Ic)::value;
                constexpr int HH = L::HH;
                constexpr int no = ((D & 1) || i < D / 2) ? HH : HH - 1;        // offsets of index i
                double acc0 = sp[i] * v[i], acc1 = 0.0;
                static_for<no>([&](auto Oc) {
                    constexpr int o = 1 + decltype(Oc)::value;
                    constexpr int j = (i + o) % D;
                    constexpr int qq = i * L::HHP + (o - 1);                // rows start on even entries: read in pairs
                    double sv;
                    if constexpr ((qq & 1) == 0) { cur = sp2[qq >> 1]; sv = cur.x; } else sv = cur.y;
                    if constexpr ((o & 1) != 0) acc0 = fma(sv, v[j], acc0); else acc1 = fma(sv, v[j], acc1);
                });
                Gm = fma(v[i], acc0 + acc1, Gm);
            });
            } else {
            static_for<D>([&](auto Ic) {
                constexpr int i = decltype(Ic)::value;
                double acc0 = sp[i] * v[i], acc1 = 0.0;
                static_for<D - 1 - i>([&](auto Jc) {
                    constexpr int j = i + 1 + decltype(Jc)::value;
                    constexpr int qq = i * D - (i * (i + 1)) / 2 + (j - i - 1);
                    double sv;
                    if constexpr ((qq & 1) == 0) { cur = sp2[qq >> 1]; sv = cur.x; } else sv = cur.y;
                    if constexpr (((j - i) & 1) != 0) acc0 = fma(sv, v[j], acc0); else acc1 = fma(sv, v[j], acc1);
                });
                Gm = fma(v[i], acc0 + acc1, Gm);
            });
            }
        }
        for (int k = 0; k < f; ++k) {
            const double qk = (row < d) ? pv.quad[row * f + k] : 0.0;
            const double s = cta_sum<NW>(qk * (2.0 * Gm), red + (k & 1) * 8, lane, warp);
            if (tid == 0) a.grad[b * f + k] = a.acceptor ? -s : s;
        }
        BPHASE_MARK(9);
    }
}

template <int D, int NW>
size_t replay_smem_bytes(int maxr, int T) {
    using L = RpLayout<D, NW>;
    const int r1 = L::R1 + 2 * maxr;
    const int phased = (r1 > L::R23 ? r1 : L::R23);
    return (size_t)(L::PHASED + ((phased + 1) & ~1) + ((T + 1) & ~1)) * sizeof(double);
}

template <int D, int NW, int MINB, int GS, bool DIRECT = false>
int launch_replay_t(const RegArgs& a, int sms, cudaStream_t st) {
    const size_t smem = replay_smem_bytes<D, NW>(a.bs.maxr, a.pv.T);
    auto kern = replay_reg_kernel<D, NW, MINB, GS, DIRECT>;
    QTET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QTET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int per_sm = 0;
    QTET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, smem));
    if (per_sm < 1) per_sm = 1;
    if (const char* env = std::getenv("QTET_RP_CTAS")) { const int v = std::atoi(env); if (v > 0 && v < per_sm) per_sm = v; }
    long long grid = (long long)sms * per_sm;
    if (grid > a.n) grid = a.n;
    kern<<<(unsigned)grid, NW * 32, smem, st>>>(a);
    count_launch();
    QTET_CUDA(cudaGetLastError());
    return QTET_OK;
}

template <int D, int NW, int MINB>
int launch_hh_t(const RegArgs& a, int sms, cudaStream_t st) {
    const size_t smem = (size_t)HhLayout<D>::TOTAL * sizeof(double);
    auto kern = hh_reg_kernel<D, NW, MINB>;
    QTET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QTET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int per_sm = 0;
    QTET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sms * per_sm;
    if (grid > a.n) grid = a.n;
    kern<<<(unsigned)grid, NW * 32, smem, st>>>(a);
    count_launch();
    QTET_CUDA(cudaGetLastError());
    return QTET_OK;
}

}  // namespace

#ifdef QTET_PHASE_TIMING
extern "C" void qtet_debug_big_phase_cycles(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_big_phase_cycles, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_big_phase_cycles, z, sizeof(z)); }
}
#endif

// Padded dimension of the register-resident kernels for Fock dimension d, or 0 when d is out of their range.
int bigreg_padded_dim(int d) {
    const int sizes[] = {16, 32, 48, 66, 84, 101};
    for (int s : sizes) if (d <= s) return s;
    return 0;
}

int bigreg_max_timesteps() { return 2048; }

static RegArgs make_args(const ProblemView& pv, const double* chis, int64_t n, int site, const BigStream& bs,
                         const unsigned short* pair_tab, double* loss, double* grad, int32_t* tstar, double* trace) {
    RegArgs a;
    a.pv = pv; a.chis = chis; a.n = n; a.site = site; a.acceptor = (site == pv.f - 1) ? 1 : 0; a.bs = bs;
    a.pair_tab = pair_tab; a.loss = loss; a.grad = grad; a.tstar = tstar; a.trace = trace;
    return a;
}

int launch_hh_reg(const ProblemView& pv, const double* chis, int64_t n, const BigStream& bs, int sms, cudaStream_t st) {
    static const bool use_frag = [] { const char* e = std::getenv("QTET_HH_FRAG"); return !(e && std::atoi(e) == 0); }();
    if (use_frag) return launch_hh_frag(pv, chis, n, bs, sms, st);
    const RegArgs a = make_args(pv, chis, n, 0, bs, nullptr, nullptr, nullptr, nullptr, nullptr);
    switch (bigreg_padded_dim(pv.d)) {
        case 16: return launch_hh_t<16, 1, 8>(a, sms, st);
        case 32: return launch_hh_t<32, 1, 8>(a, sms, st);
        case 48: return launch_hh_t<48, 2, 4>(a, sms, st);
        case 66: return launch_hh_t<66, 3, 4>(a, sms, st);
        case 84: return launch_hh_t<84, 3, 2>(a, sms, st);
        case 101: return launch_hh_t<101, 4, 2>(a, sms, st);
        default: break;
    }
    set_error("register-resident Householder kernel: dimension out of range");
    return QTET_EUNSUPPORTED;
}

int bigreg_direct_max_dim(bool dense) { return dense ? 84 : 101; }

int launch_replay_reg(const ProblemView& pv, const double* chis, int64_t n, int site, const BigStream& bs,
                      const unsigned short* pair_tab, double* loss, double* grad, int32_t* tstar, double* trace, int sms,
                      cudaStream_t st, bool direct, int* fb_list, int* fb_count) {
    RegArgs a = make_args(pv, chis, n, site, bs, pair_tab, loss, grad, tstar, trace);
    a.fb_list = fb_list; a.fb_count = fb_count;
    if (direct) {
        switch (bigreg_padded_dim(pv.d)) {
            case 16: return launch_replay_t<16, 1, 8, 4, true>(a, sms, st);
            case 32: return launch_replay_t<32, 1, 8, 4, true>(a, sms, st);
            case 48: return launch_replay_t<48, 2, 4, 4, true>(a, sms, st);
            case 66: return launch_replay_t<66, 3, 4, 4, true>(a, sms, st);
            case 84: return launch_replay_t<84, 3, 2, 4, true>(a, sms, st);
            case 101: return launch_replay_t<101, 4, 2, 2, true>(a, sms, st);
            default: break;
        }
        set_error("register-resident direct kernel: dimension out of range");
        return QTET_EUNSUPPORTED;
    }
    switch (bigreg_padded_dim(pv.d)) {
        case 16: return launch_replay_t<16, 1, 8, 4>(a, sms, st);
        case 32: return launch_replay_t<32, 1, 8, 4>(a, sms, st);
        case 48: return launch_replay_t<48, 2, 4, 4>(a, sms, st);
        case 66: return launch_replay_t<66, 3, 4, 4>(a, sms, st);
        case 84: return launch_replay_t<84, 3, 2, 4>(a, sms, st);
        case 101: return launch_replay_t<101, 4, 2, 2>(a, sms, st);
        default: break;
    }
    set_error("register-resident replay kernel: dimension out of range");
    return QTET_EUNSUPPORTED;
}

}  // namespace qtet
