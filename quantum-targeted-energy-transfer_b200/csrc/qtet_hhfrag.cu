// Householder tridiagonalisation of a dense Hamiltonian (trimers, chains) and the orthogonal factor Q, one CTA per
// parameter point, the matrix distributed over the CTA in 8 x 8 TILES held in registers (the accumulator layout of
// mma.m8n8k4: lane l of a warp holds row l / 4 and the column pair 2 (l % 4), 2 (l % 4) + 1 of every tile it owns).
//
// Why tiles.  The kernel this one replaces (hh_reg_kernel, qtet_bigreg.cu) gives every thread one ROW: each element of the
// reflector v and of p = beta A v is then a broadcast operand that feeds one or two DFMAs, i.e. ~100 LDS.128 per warp and
// column, and the shared-memory return path (l1tex data-pipe wavefronts 68 %) saturates long before the FP64 pipe (34 %).
// With a lane owning SL x 2 MT elements (3 rows x 18 columns at d = 66) a column pair loaded once serves all SL rows and a
// row value all 2 MT columns: ~30 LDS.128 per warp and column, and the two reductions the algorithm needs run over the
// four lanes of a quad (matrix-vector product) or stay inside the quad that owns row k (column norm: A is symmetric, so
// row k IS column k and sits in ONE quad) -- no CTA-wide reduction for the norm, two barriers per column.
//
//   build      H = diag(sum_k omega_k n_k + chi_k n_k^2 / 2) + hopping                       HamiltonianLoss.py:124-176
//   reduce     for k = 0 .. d-3:  v_k from row k;  p = beta A v;  A -= v w^T + w v^T,  w = p - (beta p.v / 2) v
//   Q          X = Q^T = H_{d-3} ... H_0 accumulated BACKWARDS by right-multiplications (X <- X H_k touches only rows and
//              columns > k, so dead tiles are skipped and the work is 2/3 d^3 instead of d^3), stored so that the consumer
//              (replay_reg_kernel) finds Q column-major.
//
// Tiles that lie entirely above / left of the current column are skipped STATICALLY: the column loop is unrolled over
// blocks of 8 columns (`kb`), inside a block the column index is dynamic and only selects lanes.  A warp's row tiles are
// w, w + NW, ... (interleaved, so that all warps lose tiles at the same rate); the first live slot of a warp takes one of
// two values per block, each with its own fully unrolled code.
#include <cstdlib>
#include <type_traits>
#include <utility>

#include "qtet_big.cuh"

namespace qtet {

namespace {

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
__device__ __forceinline__ void sts128(unsigned addr, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}
__device__ __forceinline__ double quad_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

template <class F, int... Is>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, Is...>) { (f(std::integral_constant<int, Is>{}), ...); }
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) { static_for_impl(f, std::make_integer_sequence<int, N>{}); }

#ifdef QTET_PHASE_TIMING
__device__ unsigned long long g_hf_phase_cycles[8];
#define HPHASE_DECL() long long _t0 = 0
#define HPHASE_START() _t0 = clock64()
#define HPHASE_MARK(idx) do { const long long _t = clock64(); if (tid == 0) atomicAdd(&g_hf_phase_cycles[idx], (unsigned long long)(_t - _t0)); _t0 = _t; } while (0)
#else
#define HPHASE_DECL() do {} while (0)
#define HPHASE_START() do {} while (0)
#define HPHASE_MARK(idx) do {} while (0)
#endif

struct HfArgs {
    ProblemView pv;
    const double* chis;
    long long n;
    double* tri_d; double* tri_e; double* qmat;
    // staged reduction (see MODE below)
    int koff = 0;                 // first row / column of the trailing block the stage works on
    double* rfl = nullptr;        // [n][koff][DP + 1]  reflectors 0 .. koff-1 of stage A (+ beta in the last entry of the row)
};

// Stage sizes.  The reduction is a latency chain per column whatever the number of live tiles, so its throughput is the number
// of matrices resident per SM: 4 at d = 66 (108 registers of matrix per lane).  Once 8 (MT - 6) columns are done the live
// block fits the 48-wide kernel (72 registers of matrix, two warps): the rest of the reduction AND the accumulation of its
// part of Q run there with one warp per point and 10 CTAs per SM.
//   MODE 0  everything in one kernel (any size)
//   MODE 1  stage A: build H, columns 0 .. koff-1, then hand over: trailing block -> scratch (the point's Q slot), reflectors -> rfl
//   MODE 2  stage B (48-wide kernel): the trailing block as its whole problem; X_B = Q_B^T lands in the trailing block of the Q slot
//   MODE 3  stage C: X = (I + X_B) H_{koff-1} ... H_0 with the reflectors of stage A, Q stored
constexpr int kStageD = 48;

template <int D, int NW>
struct HfLayout {
    static constexpr int MT = (D + 7) / 8;          // 8 x 8 tiles per side
    static constexpr int DP = 8 * MT;
    static constexpr int SL = (MT + NW - 1) / NW;   // row tiles ("slots") per warp: slot j of warp w is tile w + NW j
    // doubles
    static constexpr int PS = 0;                    // [2][DP]  p = beta A v        (by parity of the column)
    static constexpr int ROW = PS + 2 * DP;         // [2][DP]  row k+1 of the matrix before update k (pipelined reduction)
    static constexpr int DIAG = ROW + 2 * DP;       // [DP]
    static constexpr int OFF = DIAG + DP;           // [DP]
    static constexpr int BETA = OFF + DP;           // [DP]
    static constexpr int RED = BETA + DP;           // [2][2 NW <= 16]  per-warp partial sums
    static constexpr int REFL = RED + 32;           // [D + 1][DP]  row k + 1: reflector v_k (zeros up to column k); row 0: zeros
    static constexpr int TOTAL = REFL + (D + 1) * DP;
};

constexpr int reg_cap(int nw, int minb) {
    int r = 65536 / (minb * nw * 32);
    r = r / 8 * 8;
    return r > 255 ? 255 : r;
}

// PIPE: the one-barrier-per-column reduction (see the comment at its top), otherwise the plain two-barrier one.
template <int D, int NW, int MINB, bool PIPE, int MODE = 0>
__global__ void __launch_bounds__(NW * 32) __maxnreg__(reg_cap(NW, MINB)) hh_frag_kernel(HfArgs a_) {
    using L = HfLayout<D, NW>;
    constexpr int MT = L::MT, DP = L::DP, SL = L::SL, NT = NW * 32;
    extern __shared__ __align__(16) double sm[];
    const ProblemView& pv = a_.pv;
    const int dfull = pv.d, f = pv.f;
    const int koff = (MODE == 0) ? 0 : a_.koff;
    const int d = (MODE == 2) ? dfull - koff : dfull;          // the dimension this kernel's index logic sees
    constexpr int KBS = (MT > kStageD / 8) ? MT - kStageD / 8 : 0;   // MODE 1 / 3: blocks of eight columns done by stage A
    static_assert(!PIPE || MODE == 0, "the staged reduction uses the two-barrier column step");
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // tells the compiler the warp index is warp-uniform: branches on it
    const int lg = lane >> 2, lt = lane & 3;                     // need no convergence bookkeeping around the shuffles inside
    double* ps = sm + L::PS;
    double* s_diag = sm + L::DIAG;
    double* s_off = sm + L::OFF;
    double* s_beta = sm + L::BETA;
    double* red = sm + L::RED;
    double* refl = sm + L::REFL;
    const unsigned ps_a = smem_addr(ps), row_a = smem_addr(sm + L::ROW), refl_a = smem_addr(refl) + (PIPE ? 0u : (unsigned)(DP * 8));
    (void)row_a;
    const bool even_d = (d & 1) == 0;
    const bool vec_ok = even_d && ((reinterpret_cast<size_t>(pv.offd) & 15) == 0);   // 16-byte loads of the hopping rows

    for (long long b = blockIdx.x; b < a_.n; b += gridDim.x) {
        const double* chi = a_.chis + b * f;
        __syncthreads();                                 // the previous point is done with every shared region
        HPHASE_DECL();
        HPHASE_START();
        double2 a[SL][MT];
        if constexpr (MODE == 0 || MODE == 1) {
        // ---- my tiles of H.  The hopping part is the same for every point (L1-resident after the first one): all loads are
        // issued unconditionally from clamped addresses, the masks and the diagonal follow.
#pragma unroll
        for (int j = 0; j < SL; ++j) {
            const int r = 8 * (warp + NW * j) + lg;
            const double* hrow = pv.offd + (size_t)(r < d ? r : d - 1) * d;
#pragma unroll
            for (int nt = 0; nt < MT; ++nt) {
                const int c0 = 8 * nt + 2 * lt;
                if (vec_ok) {
                    a[j][nt] = __ldg(reinterpret_cast<const double2*>(hrow + (c0 + 1 < d ? c0 : 0)));
                } else {
                    a[j][nt].x = __ldg(hrow + (c0 < d ? c0 : 0));
                    a[j][nt].y = __ldg(hrow + (c0 + 1 < d ? c0 + 1 : 0));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < SL; ++j) {
            const int r = 8 * (warp + NW * j) + lg;
            double dg = 0.0;
            if (r < d)
                for (int k = 0; k < f; ++k)
                    dg = dg + (pv.omegas[k] * pv.occ[r * f + k] + (0.5 * chi[k]) * (pv.occ[r * f + k] * pv.occ[r * f + k]));
#pragma unroll
            for (int nt = 0; nt < MT; ++nt) {
                const int c0 = 8 * nt + 2 * lt;
                double x = (r < d && c0 < d) ? a[j][nt].x : 0.0, y = (r < d && c0 + 1 < d) ? a[j][nt].y : 0.0;
                if (c0 == r) x = dg;
                if (c0 + 1 == r) y = dg;
                a[j][nt] = make_double2(x, y);
            }
        }
        } else if constexpr (MODE == 2) {
            // the trailing block stage A left in the point's Q slot: [kStageD][kStageD], zero padded
            const double* tm = a_.qmat + (size_t)b * dfull * dfull;
#pragma unroll
            for (int j = 0; j < SL; ++j) {
                const int r = 8 * (warp + NW * j) + lg;
#pragma unroll
                for (int nt = 0; nt < MT; ++nt) {
                    const int c0 = 8 * nt + 2 * lt;
                    const bool in = (r < kStageD) && (c0 + 1 < kStageD);
                    a[j][nt].x = in ? tm[r * kStageD + c0] : 0.0;
                    a[j][nt].y = in ? tm[r * kStageD + c0 + 1] : 0.0;
                }
            }
        }
        HPHASE_MARK(0);


        if constexpr (PIPE) {
        // ---- reduction, ONE CTA barrier per column.  The classical column step is  reflector -> barrier -> p = beta A v
        // -> barrier -> rank-2 update, a chain of ~2400 cycles per column of which a third is one quad building the
        // reflector while everybody else waits.  Here iteration k does, in ONE sweep over a lane's tiles,
        //   * the update of column k-1:   A <- A - v w^T - w v^T                         (v = v_{k-1}, w = p - kq v)
        //   * row k of the UPDATED matrix at the lane's own columns, x_c = A'[k][c], from the row the owner quad published
        //     one iteration earlier (before the update) -- every quad redoes these ~20 flops instead of waiting for one
        //   * the products of the next step with the RAW row:  acc_r = sum_c A'[r][c] x_c,  y_r = A'[r][k+1],
        //     since v_k = x - alpha e_{k+1}:  p = beta (A' v_k) = beta (acc - alpha y)
        // followed by quad sums (sigma = |x|^2 beyond k+1, acc, y), the scalars alpha / beta / v0 in every thread, and warp
        // sums of S1 = sum_r x_r acc_r, S2 = sum_r x_r y_r that run beside the quad sums: v.p = beta (S1 - alpha S2) - alpha p_{k+1}.
        // p, (S1, S2) and the next published row cross the single barrier.
        const unsigned red_a = smem_addr(red);
        for (int i = tid; i < 2 * DP; i += NT) ps[i] = 0.0;
        for (int i = tid; i < DP; i += NT) refl[i] = 0.0;                       // "reflector -1"
        if (tid < 32) red[tid] = 0.0;
        if (warp == 0 && lg == 0) {                                             // row 0 of H for iteration 0
#pragma unroll
            for (int nt = 0; nt < MT; ++nt) sts128(row_a + (unsigned)((DP + 8 * nt + 2 * lt) * 8), a[0][nt].x, a[0][nt].y);
        }
        __syncthreads();
        double beta_p = 0.0, alpha_p = 0.0, v0p = 0.0;
        double pj[SL], vr[SL];
#pragma unroll
        for (int j = 0; j < SL; ++j) { pj[j] = 0.0; vr[j] = 0.0; }
        bool done = false;
        static_for<MT>([&](auto KBc) {
            constexpr int kb = decltype(KBc)::value;
            constexpr int wk = kb % NW, jk = kb / NW;     // warp / slot that own rows 8 kb .. 8 kb + 7
            constexpr int q0 = kb / NW, r0 = kb % NW;     // live slots: j >= q0 (warps >= r0), j >= q0 + 1 (warps < r0)
            if (!done && 8 * kb <= d - 2) {
                const bool lowj = (warp >= r0);
#pragma unroll 1
                for (int kk = 0; kk < 8; ++kk) {
                    const int k = 8 * kb + kk;
                    if (k > d - 2) { done = true; break; }
                    const bool last = (k == d - 2);           // only the update of column d-3 is left
                    const unsigned par = (unsigned)(k & 1), pp = par ^ 1u;
                    const unsigned vprev_a = refl_a + (unsigned)(k * DP * 8), vcur_a = vprev_a + (unsigned)(DP * 8);
                    const unsigned psp_a = ps_a + pp * (DP * 8);
                    const unsigned rowp_a = row_a + pp * (DP * 8), rowc_a = row_a + par * (DP * 8);
                    auto body = [&](auto J0c) {
                        constexpr int J0 = decltype(J0c)::value;
                        if constexpr (J0 < SL) {
                            // -- scalars of the update of column k-1
                            double S1 = 0.0, S2 = 0.0;
#pragma unroll
                            for (int w = 0; w < NW; ++w) {
                                const double2 s2 = lds128(red_a + (pp * 16 + 2 * w) * 8);
                                S1 += s2.x; S2 += s2.y;
                            }
                            const double pk = lds64(psp_a + (unsigned)(k * 8));
                            const double vp = fma(beta_p, fma(-alpha_p, S2, S1), -(alpha_p * pk));
                            const double kq = 0.5 * beta_p * vp;
                            const double wk_ = fma(-kq, v0p, pk);                 // w_{k-1}[k]
                            double wr[SL], xr[SL];
#pragma unroll
                            for (int j = J0; j < SL; ++j) {
                                const int mt = warp + NW * j, r = 8 * mt + lg;
                                wr[j] = fma(-kq, vr[j], pj[j]);
                                const double rw = (mt < MT) ? lds64(rowp_a + (unsigned)(r * 8)) : 0.0;
                                const double x = fma(-wk_, vr[j], fma(-v0p, wr[j], rw));      // row k of the updated matrix at index r (symmetry)
                                xr[j] = (mt < MT && r > k && r < d) ? x : 0.0;
                            }
                            // -- the sweep
                            double acc0[SL], acc1[SL], yy[SL];
#pragma unroll
                            for (int j = J0; j < SL; ++j) { acc0[j] = 0.0; acc1[j] = 0.0; yy[j] = 0.0; }
                            double s0 = 0.0, s1 = 0.0, x0 = 0.0, dgk = 0.0;
                            static_for<MT - kb>([&](auto Nc) {
                                constexpr int nt = kb + decltype(Nc)::value;
                                const int c0 = 8 * nt + 2 * lt;
                                const double2 v2 = lds128(vprev_a + (unsigned)(c0 * 8));
                                const double2 p2 = lds128(psp_a + (unsigned)(c0 * 8));
                                const double2 r2 = lds128(rowp_a + (unsigned)(c0 * 8));
                                const double wx = fma(-kq, v2.x, p2.x), wy = fma(-kq, v2.y, p2.y);
                                double xx = fma(-wk_, v2.x, fma(-v0p, wx, r2.x)), xy = fma(-wk_, v2.y, fma(-v0p, wy, r2.y));
                                if constexpr (nt <= kb + 1) {          // only the two tiles around columns k, k+1 need masks
                                    if (c0 == k) dgk = xx;
                                    if (c0 + 1 == k) dgk = xy;
                                    if (c0 == k + 1) x0 = xx;
                                    if (c0 + 1 == k + 1) x0 = xy;
                                    xx = (c0 > k) ? xx : 0.0; xy = (c0 + 1 > k) ? xy : 0.0;
                                    const double sx = (c0 > k + 1) ? xx : 0.0, sy = (c0 + 1 > k + 1) ? xy : 0.0;
                                    s0 = fma(sx, sx, s0); s1 = fma(sy, sy, s1);
                                } else {
                                    s0 = fma(xx, xx, s0); s1 = fma(xy, xy, s1);
                                }
                                if (warp == wk && lg == 0) sts128(vcur_a + (unsigned)(c0 * 8), xx, xy);   // raw row; entry k+1 is fixed below
#pragma unroll
                                for (int j = J0; j < SL; ++j) {
                                    a[j][nt].x = fma(-vr[j], wx, fma(-wr[j], v2.x, a[j][nt].x));
                                    a[j][nt].y = fma(-vr[j], wy, fma(-wr[j], v2.y, a[j][nt].y));
                                    acc0[j] = fma(a[j][nt].x, xx, acc0[j]);
                                    acc1[j] = fma(a[j][nt].y, xy, acc1[j]);
                                    if constexpr (nt <= kb + 1) {
                                        if (c0 == k + 1) yy[j] = a[j][nt].x;
                                        if (c0 + 1 == k + 1) yy[j] = a[j][nt].y;
                                    }
                                }
                            });
                            if (!last) {
                                // -- publish row k+1 of the matrix as it is now (update k-1 applied, update k not yet)
                                if (kk < 7) {
                                    if (warp == wk && lg == kk + 1) {
                                        static_for<MT - kb>([&](auto Nc) {
                                            constexpr int nt = kb + decltype(Nc)::value;
                                            sts128(rowc_a + (unsigned)((8 * nt + 2 * lt) * 8), a[jk][nt].x, a[jk][nt].y);
                                        });
                                    }
                                } else {
                                    constexpr int kb1 = (kb + 1 < MT) ? kb + 1 : kb;
                                    constexpr int w1 = kb1 % NW, j1 = kb1 / NW;
                                    if (warp == w1 && lg == 0) {
                                        static_for<MT - kb>([&](auto Nc) {
                                            constexpr int nt = kb + decltype(Nc)::value;
                                            sts128(rowc_a + (unsigned)((8 * nt + 2 * lt) * 8), a[j1][nt].x, a[j1][nt].y);
                                        });
                                    }
                                }
                                // -- sums: S1, S2 over the warp (partials are linear in the lanes' pieces), the rest over the quad
                                double S1p = 0.0, S2p = 0.0;
#pragma unroll
                                for (int j = J0; j < SL; ++j) {
                                    acc0[j] += acc1[j];
                                    S1p = fma(xr[j], acc0[j], S1p);
                                    S2p = fma(xr[j], yy[j], S2p);
                                }
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) {
                                    S1p += __shfl_xor_sync(0xffffffffu, S1p, o);
                                    S2p += __shfl_xor_sync(0xffffffffu, S2p, o);
                                }
                                const double sigma = quad_sum(s0 + s1);
                                x0 = quad_sum(x0);
                                dgk = quad_sum(dgk);
#pragma unroll
                                for (int j = J0; j < SL; ++j) { acc0[j] = quad_sum(acc0[j]); yy[j] = quad_sum(yy[j]); }
                                // -- reflector k (every thread: same inputs, same operations, same result)
                                double alpha = x0, beta = 0.0, v0 = 0.0;
                                if (sigma != 0.0) {
                                    const double nrm2 = fma(x0, x0, sigma);
                                    const double mu = nrm2 * rsqrt(nrm2);
                                    alpha = (x0 >= 0.0) ? -mu : mu;
                                    v0 = x0 - alpha;
                                    beta = 2.0 * fast_rcp(fma(v0, v0, sigma));
                                }
#pragma unroll
                                for (int j = J0; j < SL; ++j) {
                                    const int mt = warp + NW * j, r = 8 * mt + lg;
                                    const bool in = (mt < MT) && (r > k) && (r < d);
                                    pj[j] = in ? beta * fma(-alpha, yy[j], acc0[j]) : 0.0;
                                    vr[j] = (beta == 0.0) ? 0.0 : ((r == k + 1) ? v0 : ((r > k + 1) ? xr[j] : 0.0));
                                    if (lt == 0 && mt < MT) ps[par * DP + r] = pj[j];
                                }
                                if (lane == 0) { red[par * 16 + 2 * warp] = S1p; red[par * 16 + 2 * warp + 1] = S2p; }
                                if (warp == wk && lane == 0) {
                                    refl[(size_t)(k + 1) * DP + k + 1] = v0;
                                    s_diag[k] = dgk; s_off[k] = alpha; s_beta[k] = beta;
                                }
                                beta_p = beta; alpha_p = alpha; v0p = v0;
                            }
                        } else {
                            if (lane == 0) { red[par * 16 + 2 * warp] = 0.0; red[par * 16 + 2 * warp + 1] = 0.0; }
                        }
                    };
                    if (lowj) body(std::integral_constant<int, q0>{});
                    else body(std::integral_constant<int, (q0 + 1 < SL ? q0 + 1 : SL)>{});
                    if (last) { done = true; break; }
                    __syncthreads();
                }
            }
        });
        } else {
        // ---- reduction, two CTA barriers per column
        static_for<MT>([&](auto KBc) {
            constexpr int kb = decltype(KBc)::value;
            constexpr int wk = kb % NW, jk = kb / NW;     // warp / slot that own rows 8 kb .. 8 kb + 7
            constexpr int q0 = kb / NW, r0 = kb % NW;     // live slots: j >= q0 (warps >= r0), j >= q0 + 1 (warps < r0)
            if constexpr (MODE != 3 && (MODE != 1 || kb < KBS))
            if (8 * kb < d - 2) {
                const bool lowj = (warp >= r0);
#pragma unroll 1
                for (int kk = 0; kk < 8; ++kk) {
                    const int k = 8 * kb + kk;
                    if (k >= d - 2) break;
                    const unsigned rk_a = refl_a + (unsigned)(k * DP * 8);
                    if (warp == wk) {
                        // row 8 kb + lg of the current matrix; the quad lg == kk holds row k = column k (symmetry)
                        double s0 = 0.0, s1 = 0.0, x0 = 0.0, dgk = 0.0;
                        static_for<MT - kb>([&](auto Nc) {
                            constexpr int nt = kb + decltype(Nc)::value;
                            const int c0 = 8 * nt + 2 * lt;
                            const double x = a[jk][nt].x, y = a[jk][nt].y;
                            if constexpr (nt <= kb + 1) {          // only the two tiles around column k + 1 need masks
                                const double xs_ = (c0 > k + 1) ? x : 0.0, ys_ = (c0 + 1 > k + 1) ? y : 0.0;
                                s0 = fma(xs_, xs_, s0); s1 = fma(ys_, ys_, s1);
                            } else {
                                s0 = fma(x, x, s0); s1 = fma(y, y, s1);
                            }
                            if constexpr (nt <= kb + 1) {
                                if (c0 == k + 1) x0 = x;
                                if (c0 + 1 == k + 1) x0 = y;
                                if (c0 == k) dgk = x;
                                if (c0 + 1 == k) dgk = y;
                            }
                        });
                        const double sigma = quad_sum(s0 + s1);
                        x0 = quad_sum(x0);
                        dgk = quad_sum(dgk);
                        if (lg == kk) {
                            double alpha = x0, beta = 0.0, v0 = 0.0;
                            if (sigma != 0.0) {
                                const double nrm2 = fma(x0, x0, sigma);
                                const double mu = nrm2 * rsqrt(nrm2);     // a few ulp off sqrt: only moves the tridiagonal entry by that much;
                                alpha = (x0 >= 0.0) ? -mu : mu;           // H stays orthogonal because beta is formed from the v actually used
                                v0 = x0 - alpha;
                                beta = 2.0 * fast_rcp(fma(v0, v0, sigma));
                            }
                            static_for<MT - kb>([&](auto Nc) {
                                constexpr int nt = kb + decltype(Nc)::value;
                                const int c0 = 8 * nt + 2 * lt;
                                double x = a[jk][nt].x, y = a[jk][nt].y;
                                if constexpr (nt <= kb + 1) {
                                    x = (c0 > k + 1) ? x : ((c0 == k + 1) ? v0 : 0.0);
                                    y = (c0 + 1 > k + 1) ? y : ((c0 + 1 == k + 1) ? v0 : 0.0);
                                }
                                sts128(rk_a + (unsigned)(c0 * 8), x, y);
                            });
                            if (lt == 0) { s_diag[k] = dgk; s_off[k] = alpha; s_beta[k] = beta; }
                        }
                    }
                    HPHASE_MARK(4);
                    __syncthreads();                             // #1: reflector k is in shared memory
                    HPHASE_MARK(5);
                    const double beta = s_beta[k];
                    if (beta == 0.0) continue;                   // nothing to annihilate in this column (CTA-uniform)
                    double pj[SL], vr[SL];
#pragma unroll
                    for (int j = 0; j < SL; ++j) { pj[j] = 0.0; vr[j] = 0.0; }
                    // p = beta A v over the live tiles; quad sums finish the rows
                    auto matvec = [&](auto J0c) {
                        constexpr int J0 = decltype(J0c)::value;
                        if constexpr (J0 < SL) {
                            double acc0[SL], acc1[SL];
#pragma unroll
                            for (int j = J0; j < SL; ++j) { acc0[j] = 0.0; acc1[j] = 0.0; }
                            static_for<MT - kb>([&](auto Nc) {
                                constexpr int nt = kb + decltype(Nc)::value;
                                const double2 v2 = lds128(rk_a + (unsigned)((8 * nt + 2 * lt) * 8));
#pragma unroll
                                for (int j = J0; j < SL; ++j) {
                                    acc0[j] = fma(a[j][nt].x, v2.x, acc0[j]);
                                    acc1[j] = fma(a[j][nt].y, v2.y, acc1[j]);
                                }
                            });
                            double t = 0.0;
#pragma unroll
                            for (int j = J0; j < SL; ++j) {
                                const int mt = warp + NW * j, r = 8 * mt + lg;
                                const double s = quad_sum(acc0[j] + acc1[j]);
                                const bool in = (mt < MT) && (r > k) && (r < d);
                                pj[j] = in ? beta * s : 0.0;
                                vr[j] = (mt < MT) ? lds64(rk_a + (unsigned)(r * 8)) : 0.0;
                                if (lt == 0 && mt < MT) ps[r] = pj[j];
                                t = fma(vr[j], pj[j], t);
                            }
                            t = (lt == 0) ? t : 0.0;
                            t += __shfl_xor_sync(0xffffffffu, t, 4);
                            t += __shfl_xor_sync(0xffffffffu, t, 8);
                            t += __shfl_xor_sync(0xffffffffu, t, 16);
                            if (lane == 0) red[warp] = t;
                        } else {
                            if (lane == 0) red[warp] = 0.0;
                        }
                    };
                    if (lowj) matvec(std::integral_constant<int, q0>{});
                    else matvec(std::integral_constant<int, (q0 + 1 < SL ? q0 + 1 : SL)>{});
                    HPHASE_MARK(6);
                    __syncthreads();                             // #2: p and the partial sums of v . p are in shared memory
                    HPHASE_MARK(7);
                    double vp = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) vp += red[w];
                    const double kq = 0.5 * beta * vp;
                    // A -= v w^T + w v^T with w = p - kq v
                    auto update = [&](auto J0c) {
                        constexpr int J0 = decltype(J0c)::value;
                        if constexpr (J0 < SL) {
                            double wr[SL];
#pragma unroll
                            for (int j = J0; j < SL; ++j) wr[j] = fma(-kq, vr[j], pj[j]);
                            static_for<MT - kb>([&](auto Nc) {
                                constexpr int nt = kb + decltype(Nc)::value;
                                const double2 v2 = lds128(rk_a + (unsigned)((8 * nt + 2 * lt) * 8));
                                const double2 p2 = lds128(ps_a + (unsigned)((8 * nt + 2 * lt) * 8));
                                const double wx = fma(-kq, v2.x, p2.x), wy = fma(-kq, v2.y, p2.y);
#pragma unroll
                                for (int j = J0; j < SL; ++j) {
                                    a[j][nt].x = fma(-vr[j], wx, fma(-wr[j], v2.x, a[j][nt].x));
                                    a[j][nt].y = fma(-vr[j], wy, fma(-wr[j], v2.y, a[j][nt].y));
                                }
                            });
                        }
                    };
                    if (lowj) update(std::integral_constant<int, q0>{});
                    else update(std::integral_constant<int, (q0 + 1 < SL ? q0 + 1 : SL)>{});
                }
            }
        });
        }
        HPHASE_MARK(1);
        if constexpr (MODE == 1) {
            // ---- stage A hands over: tridiagonal entries of its columns, reflectors (+ beta), trailing block
            __syncthreads();
            for (int i = tid; i < koff; i += NT) {
                a_.tri_d[(size_t)b * d + i] = s_diag[i];
                a_.tri_e[(size_t)b * d + i] = s_off[i];
            }
            {
                double* rf = a_.rfl + (size_t)b * koff * (DP + 1);
                for (int i = tid; i < koff * (DP + 1); i += NT) {
                    const int k = i / (DP + 1), c = i - k * (DP + 1);
                    rf[i] = (c < DP) ? refl[(size_t)(k + 1) * DP + c] : s_beta[k];      // (reflector k sits in row k + 1 of the buffer)
                }
                double* tm = a_.qmat + (size_t)b * d * d;
#pragma unroll
                for (int j = 0; j < SL; ++j) {
                    const int r = 8 * (warp + NW * j) + lg - 8 * KBS;
#pragma unroll
                    for (int nt = KBS; nt < MT; ++nt) {
                        const int c0 = 8 * (nt - KBS) + 2 * lt;
                        if (r >= 0 && r < kStageD) { tm[r * kStageD + c0] = a[j][nt].x; tm[r * kStageD + c0 + 1] = a[j][nt].y; }
                    }
                }
            }
            HPHASE_MARK(3);
            continue;
        }
        if constexpr (MODE == 0 || MODE == 2) {
        // ---- the last 2 x 2 block (rows d-2, d-1)
#pragma unroll
        for (int j = 0; j < SL; ++j) {
            const int r = 8 * (warp + NW * j) + lg;
#pragma unroll
            for (int nt = 0; nt < MT; ++nt) {
                const int c0 = 8 * nt + 2 * lt;
                if (r < d && r >= d - 2) {
                    if (c0 == r) s_diag[r] = a[j][nt].x;
                    if (c0 + 1 == r) s_diag[r] = a[j][nt].y;
                    if (r == d - 1 && d >= 2) {
                        if (c0 == d - 2) s_off[d - 2] = a[j][nt].x;
                        if (c0 + 1 == d - 2) s_off[d - 2] = a[j][nt].y;
                    }
                }
            }
        }
        if (tid == 0) s_off[d - 1] = 0.0;
        __syncthreads();
        for (int i = tid; i < d; i += NT) {
            a_.tri_d[(size_t)b * dfull + koff + i] = s_diag[i];
            a_.tri_e[(size_t)b * dfull + koff + i] = s_off[i];
        }
        }

        // ---- X = Q^T = H_{d-3} ... H_0, accumulated backwards (no barrier: the reflectors are all in shared memory)
#pragma unroll
        for (int j = 0; j < SL; ++j) {
            const int r = 8 * (warp + NW * j) + lg;
#pragma unroll
            for (int nt = 0; nt < MT; ++nt) {
                const int c0 = 8 * nt + 2 * lt;
                a[j][nt] = make_double2((c0 == r) ? 1.0 : 0.0, (c0 + 1 == r) ? 1.0 : 0.0);
            }
        }
        if constexpr (MODE == 3) {
            // reflectors of stage A back into shared memory, X_B (stage B) into the trailing block of X
            const double* rf = a_.rfl + (size_t)b * koff * (DP + 1);
            constexpr int KOFF = 8 * KBS;                        // = koff: all loads of a thread are independent and issue together
            for (int c = tid; c < DP + 1; c += NT) {
                double tmp[KOFF > 0 ? KOFF : 1];
#pragma unroll
                for (int k = 0; k < KOFF; ++k) tmp[k] = __ldg(rf + k * (DP + 1) + c);
#pragma unroll
                for (int k = 0; k < KOFF; ++k) {
                    if (c < DP) refl[(size_t)(k + 1) * DP + c] = tmp[k]; else s_beta[k] = tmp[k];
                }
            }
            const double* q = a_.qmat + (size_t)b * d * d;
#pragma unroll
            for (int j = 0; j < SL; ++j) {
                const int r = 8 * (warp + NW * j) + lg;
#pragma unroll
                for (int nt = KBS; nt < MT; ++nt) {
                    const int c0 = 8 * nt + 2 * lt;
                    // unconditional loads from clamped addresses (they all issue at once), then the masks
                    const int rc = (r >= koff && r < d) ? r : koff;
                    const double x = __ldg(q + (size_t)rc * d + (c0 < d ? c0 : koff)), y = __ldg(q + (size_t)rc * d + (c0 + 1 < d ? c0 + 1 : koff));
                    if (r >= koff && r < d) {
                        if (c0 < d) a[j][nt].x = x;
                        if (c0 + 1 < d) a[j][nt].y = y;
                    }
                }
            }
            __syncthreads();
        }
        static_for<MT>([&](auto KBc) {
            constexpr int kb = MT - 1 - decltype(KBc)::value;
            constexpr int q0 = kb / NW, r0 = kb % NW;
            if constexpr (MODE != 3 || kb < KBS)
            if (8 * kb < d - 2) {
                const bool lowj = (warp >= r0);
#pragma unroll 1
                for (int kk = 7; kk >= 0; --kk) {
                    const int k = 8 * kb + kk;
                    if (k > d - 3) continue;
                    const double beta = s_beta[k];
                    if (beta == 0.0) continue;
                    const unsigned rk_a = refl_a + (unsigned)((k + (PIPE ? 1 : 0)) * DP * 8);
                    auto apply = [&](auto J0c) {
                        constexpr int J0 = decltype(J0c)::value;
                        if constexpr (J0 < SL) {
                            double acc0[SL], acc1[SL];
#pragma unroll
                            for (int j = J0; j < SL; ++j) { acc0[j] = 0.0; acc1[j] = 0.0; }
                            static_for<MT - kb>([&](auto Nc) {
                                constexpr int nt = kb + decltype(Nc)::value;
                                const double2 v2 = lds128(rk_a + (unsigned)((8 * nt + 2 * lt) * 8));
#pragma unroll
                                for (int j = J0; j < SL; ++j) {
                                    acc0[j] = fma(a[j][nt].x, v2.x, acc0[j]);
                                    acc1[j] = fma(a[j][nt].y, v2.y, acc1[j]);
                                }
                            });
                            double u[SL];
#pragma unroll
                            for (int j = J0; j < SL; ++j) u[j] = beta * quad_sum(acc0[j] + acc1[j]);
                            static_for<MT - kb>([&](auto Nc) {
                                constexpr int nt = kb + decltype(Nc)::value;
                                const double2 v2 = lds128(rk_a + (unsigned)((8 * nt + 2 * lt) * 8));
#pragma unroll
                                for (int j = J0; j < SL; ++j) {
                                    a[j][nt].x = fma(-u[j], v2.x, a[j][nt].x);
                                    a[j][nt].y = fma(-u[j], v2.y, a[j][nt].y);
                                }
                            });
                        }
                    };
                    if (lowj) apply(std::integral_constant<int, q0>{});
                    else apply(std::integral_constant<int, (q0 + 1 < SL ? q0 + 1 : SL)>{});
                }
            }
        });
        HPHASE_MARK(2);
        // ---- Q column-major = X row-major: the column pair of a lane is contiguous  (stage B: into the trailing block)
        {
            const int soff = (MODE == 2) ? koff : 0;         // stage B writes the trailing block, everybody else the whole matrix
            double* q = a_.qmat + (size_t)b * dfull * dfull + (size_t)soff * dfull + soff;
            const bool vec = ((dfull | soff) & 1) == 0;
            if constexpr (MODE == 2) __syncthreads();        // every thread has read the trailing block that lives in this slot
#pragma unroll
            for (int j = 0; j < SL; ++j) {
                const int r = 8 * (warp + NW * j) + lg;
                if (r < d) {
#pragma unroll
                    for (int nt = 0; nt < MT; ++nt) {
                        const int c0 = 8 * nt + 2 * lt;
                        if (vec && c0 + 1 < d) {
                            *reinterpret_cast<double2*>(q + (size_t)r * dfull + c0) = a[j][nt];
                        } else {
                            if (c0 < d) q[(size_t)r * dfull + c0] = a[j][nt].x;
                            if (c0 + 1 < d) q[(size_t)r * dfull + c0 + 1] = a[j][nt].y;
                        }
                    }
                }
            }
        }
        HPHASE_MARK(3);
    }
}

template <int D, int NW, int MINB, bool PIPE, int MODE = 0>
int launch_hf_t(const HfArgs& a, int sms, cudaStream_t st) {
    const size_t smem = (size_t)HfLayout<D, NW>::TOTAL * sizeof(double);
    auto kern = hh_frag_kernel<D, NW, MINB, PIPE, MODE>;
    static int per_sm_cached[64] = {};               // per instantiation and device: the function attributes are set once
    int dev = 0;
    QTET_CUDA(cudaGetDevice(&dev));
    dev &= 63;
    if (per_sm_cached[dev] == 0) {
        QTET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        QTET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int per_sm = 0;
        QTET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, smem));
        per_sm_cached[dev] = per_sm < 1 ? 1 : per_sm;
    }
    int per_sm = per_sm_cached[dev];
    if (const char* env = std::getenv("QTET_HH_CTAS")) { const int v = std::atoi(env); if (v > 0 && v < per_sm) per_sm = v; }
    long long grid = (long long)sms * per_sm;
    if (grid > a.n) grid = a.n;
    kern<<<(unsigned)grid, NW * 32, smem, st>>>(a);
    count_launch();
    QTET_CUDA(cudaGetLastError());
    return QTET_OK;
}

}  // namespace

#ifdef QTET_PHASE_TIMING
extern "C" void qtet_debug_hf_phase_cycles(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_hf_phase_cycles, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_hf_phase_cycles, z, sizeof(z)); }
}
#endif

// Stage A / B / C of the staged reduction (see MODE at the top): the wide kernel for the first koff columns, the 48-wide one for
// the rest of the reduction and its part of Q, the wide one again for the reflectors of stage A.
template <int D, int NW, int MINB>
int launch_hf_staged(HfArgs a, int sms, cudaStream_t st) {
    a.koff = 8 * (HfLayout<D, NW>::MT - kStageD / 8);
    int rc = launch_hf_t<D, NW, MINB, false, 1>(a, sms, st);
    if (rc) return rc;
    // stage B with ONE warp per point (six row tiles per lane, 10 CTAs per SM): no cross-warp sums, cheap barriers.  Measured on
    // the trimer N=10 (whole Householder stage per 65536 points): 10.1 ms, against 10.8 ms with two warps x 8 CTAs and x 6 CTAs.
    rc = launch_hf_t<kStageD, 1, 10, false, 2>(a, sms, st);
    if (rc) return rc;
    return launch_hf_t<D, NW, MINB, false, 3>(a, sms, st);
}

// bytes of hand-over scratch (reflectors of stage A) per point, 0 when the dimension is reduced in one kernel
size_t hh_frag_scratch_bytes(int d) {
    const int D = bigreg_padded_dim(d);
    if (D <= kStageD) return 0;
    const int MT = (D + 7) / 8, koff = 8 * (MT - kStageD / 8);
    return (size_t)koff * (8 * MT + 1) * sizeof(double);
}

int launch_hh_frag(const ProblemView& pv, const double* chis, int64_t n, const BigStream& bs, int sms, cudaStream_t st) {
    HfArgs a;
    a.pv = pv; a.chis = chis; a.n = n; a.tri_d = bs.tri_d; a.tri_e = bs.tri_e; a.qmat = bs.qmat; a.rfl = bs.rfl;
    static const bool staged_on = [] { const char* e = std::getenv("QTET_HH_STAGED"); return !(e && std::atoi(e) == 0); }();
    if (staged_on && bs.rfl != nullptr) {
        switch (bigreg_padded_dim(pv.d)) {
            case 66: return launch_hf_staged<66, 3, 4>(a, sms, st);
            case 84: return launch_hf_staged<84, 3, 2>(a, sms, st);
            case 101: return launch_hf_staged<101, 4, 2>(a, sms, st);
            default: break;
        }
    }
    // QTET_HH_FRAG=2 selects the one-barrier (pipelined) reduction: correct, but measured SLOWER (trimer N=10: 22.7 ms against 13.0 ms
    // per 65536 points -- its longer straight-line body stalls on instruction fetch, 17 % no_instruction, and spills)
    static const bool pipe = [] { const char* e = std::getenv("QTET_HH_FRAG"); return e && std::atoi(e) == 2; }();
    switch (bigreg_padded_dim(pv.d)) {
        case 16: return pipe ? launch_hf_t<16, 1, 8, true>(a, sms, st) : launch_hf_t<16, 1, 8, false>(a, sms, st);
        case 32: return pipe ? launch_hf_t<32, 1, 8, true>(a, sms, st) : launch_hf_t<32, 1, 8, false>(a, sms, st);
        case 48: return pipe ? launch_hf_t<48, 2, 4, true>(a, sms, st) : launch_hf_t<48, 2, 4, false>(a, sms, st);
        case 66: return pipe ? launch_hf_t<66, 3, 4, true>(a, sms, st) : launch_hf_t<66, 3, 4, false>(a, sms, st);
        case 84: return pipe ? launch_hf_t<84, 3, 2, true>(a, sms, st) : launch_hf_t<84, 3, 2, false>(a, sms, st);
        case 101: return pipe ? launch_hf_t<101, 4, 2, true>(a, sms, st) : launch_hf_t<101, 4, 2, false>(a, sms, st);
        default: break;
    }
    set_error("tiled Householder kernel: dimension out of range");
    return QTET_EUNSUPPORTED;
}

}  // namespace qtet

#ifdef QTET_DEBUG_EXPORTS
// debug builds only: run the Householder stage alone on caller-provided device buffers (tools/debug_hh.py)
extern "C" int qtet_debug_hh(const qtet_problem* p, const double* chis, long long n, double* tri_d, double* tri_e, double* qmat, double* rfl) {
    qtet::BigStream bs;
    bs.tri_d = tri_d; bs.tri_e = tri_e; bs.qmat = qmat; bs.rfl = rfl;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
    const int rc = qtet::launch_hh_frag(p->view, chis, n, bs, sms, 0);
    cudaDeviceSynchronize();
    return rc;
}
#endif
