// Fused generic kernel: one CTA per parameter point, any Fock dimension that fits shared memory.
//
//   H build (HamiltonianLoss.py:124-176) -> [Householder tridiagonalisation if H is not tridiagonal]
//   -> implicit-shift QL with accumulated eigenvectors (replaces tf.linalg.eigh, :182)
//   -> psi_j(t) = sum_i V_ji c_i exp(-i e_i t), <n(t)> = sum_j n_site(j) |psi_j|^2   (:204-215)
//   -> loss = N - max_t / min_t (:227-231) -> analytic gradient (replaces Optimizer.py:85-91).
//
// Everything lives in shared memory: V [d x ldv] (eigenvectors) and M [d x ldv] (the dense matrix
// during tridiagonalisation, later the Daleckii-Krein kernel K).  Every warp keeps a private copy of
// the tridiagonal (diag/off) and runs the scalar QL recurrence redundantly, so the QL phase needs no
// block-level barrier: each thread applies the rotations to its own rows of V.
//
// This kernel is the general/robust path (trimer, large dimers) and the on-GPU cross-check of the
// fast split path in qtet_split.cu.
#include <mutex>

#include "qtet_common.cuh"
#include "qtet_big.cuh"

namespace qtet {

namespace {

constexpr double kEps = 1.1102230246251565e-16;   // 2^-53
constexpr int kMaxQlIter = 60;

struct GenericArgs {
    ProblemView pv;
    const double* chis;
    long long B;
    int site;
    int acceptor;
    int ldv;
    double* loss;
    double* grad;
    int* tstar;
    double* trace;
    const int* list;      // optional: evaluate only points list[0 .. *count)
    const int* count;
    // mode 0: fused (H build, [Householder], QL, evolution, gradient)
    // mode 1: H build + Householder only -> tri_d / tri_e / qmat           (dense H, feeds the scalar QL kernel)
    // mode 2: eigenvectors by REPLAYING the rotation stream of the scalar QL kernel on V = I (or Q), then
    //         evolution + gradient as in mode 0
    int mode;
    BigStream bs;
    // Fock dimensions whose V and M do not fit shared memory (d > 115): the two matrices live in a per-CTA slab of global
    // memory (L1 / L2 cached) instead -- slow, but every size the reference accepts returns a result
    double* slab = nullptr;       // [gridDim.x][2][d][ldv]
};

__device__ __forceinline__ double block_sum(double v, double* red, int tid, int nwarps) {
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += red[w];
    return s;
}

__global__ void generic_kernel(GenericArgs a) {
    extern __shared__ double smem[];
    const ProblemView& pv = a.pv;
    const int d = pv.d, f = pv.f, T = pv.T, ld = a.ldv;
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nth >> 5;

    double* V = a.slab ? a.slab + (size_t)blockIdx.x * 2 * d * ld : smem;      // [d][ld]
    double* M = V + (size_t)d * ld;         // [d][ld]
    double* wdiag = a.slab ? smem : M + (size_t)d * ld;     // [nwarps][d]
    double* woff = wdiag + nwarps * d;      // [nwarps][d]
    double* cvec = woff + nwarps * d;       // [d]   c_i = V[0][i]
    double* zr = cvec + d;                  // [d]
    double* zi = zr + d;
    double* fr = zi + d;                    // [d]   exp(-i e_i t*)
    double* fi = fr + d;
    double* ar = fi + d;                    // [d]   a = V^T (w conj psi)
    double* ai = ar + d;
    double* wr = ai + d;                    // [d]   w_j conj(psi_j) ; also Householder v
    double* wi = wr + d;                    // [d]   also Householder w
    double* tr = wi + d;                    // [T]
    double* red = tr + T;                   // [32]
    int* ired = (int*)(red + 32);           // [4]

    double* mydiag = wdiag + warp * d;
    double* myoff = woff + warp * d;

    const long long limit = a.list ? (long long)(*a.count) : a.B;
    for (long long q = blockIdx.x; q < limit; q += gridDim.x) {
        const long long b = a.list ? (long long)a.list[q] : q;
        const double* chi = a.chis + b * f;
        __syncthreads();
        // ---------------------------------------------------------------- H build
        // V <- identity   (mode 2, dense H: V <- Q of the Householder reduction done by the mode-1 launch)
        if (a.mode == 2 && !pv.tridiag) {
            const double* q = a.bs.qmat + (size_t)b * d * d;
            for (int idx = tid; idx < d * d; idx += nth) {
                const int r = idx / d, c = idx - r * d;
                V[r * ld + c] = q[idx];
            }
        } else {
            for (int idx = tid; idx < d * d; idx += nth) {
                const int r = idx / d, c = idx - r * d;
                V[r * ld + c] = (r == c) ? 1.0 : 0.0;
            }
        }
        if (a.mode == 2) {
            // ------------------------------------------------------------ replay of the recorded rounds
            // thread = row of V.  A round's (c, s) of my point are staged in shared memory (double-buffered,
            // next round prefetched into a register); every row then walks the planes top-down keeping the
            // chain element in a register: one LDS + one STS + four FP64 per plane.
            __syncthreads();
            const long long batch = b >> 5;
            const int blane = (int)(b & 31);
            const int nr = a.bs.nrounds[batch];
            const double2* rec = a.bs.stream + (size_t)batch * a.bs.capit * 32 + blane;
            const unsigned* hdr = a.bs.hdr + (size_t)batch * a.bs.maxr * 32 + blane;
            const int* it0s = a.bs.it0 + (size_t)batch * a.bs.maxr;
            double2* csb = reinterpret_cast<double2*>(M + (((size_t)d * ld) & 1));   // [2][d], 16 B aligned (M is free until the gradient)
            unsigned h = (nr > 0) ? hdr[0] : 0u;
            int it0 = (nr > 0) ? it0s[0] : 0;
            double2 pre = make_double2(1.0, 0.0);
            {
                const int l0 = h & 0xff, mt0 = (h >> 8) & 0xff;
                if (nr > 0 && l0 != 0xff && tid < mt0 - l0) pre = rec[(size_t)(it0 + tid) * 32];
            }
            for (int r = 0; r < nr; ++r) {
                const int lp = h & 0xff, mtop = (h >> 8) & 0xff;
                const int cnt = (lp == 0xff) ? 0 : (mtop - lp);
                double2* cs = csb + (r & 1) * d;
                if (tid < cnt) cs[tid] = pre;
                // prefetch the next round
                unsigned hn = 0u; int it0n = 0;
                pre = make_double2(1.0, 0.0);
                if (r + 1 < nr) {
                    hn = hdr[(size_t)(r + 1) * 32];
                    it0n = it0s[r + 1];
                    const int ln = hn & 0xff, mtn = (hn >> 8) & 0xff;
                    if (ln != 0xff && tid < mtn - ln) pre = rec[(size_t)(it0n + tid) * 32];
                }
                __syncthreads();
                if (cnt > 0) {
                    for (int row = tid; row < d; row += nth) {
                        double* vr = V + (size_t)row * ld;
                        double w = vr[mtop];
                        int j = 0;
                        // four planes per trip: loads first, then the four dependent steps (one fma each on the chain)
                        for (; j + 4 <= cnt; j += 4) {
                            const int i = mtop - 1 - j;
                            const double2 c0 = cs[j], c1 = cs[j + 1], c2 = cs[j + 2], c3 = cs[j + 3];
                            const double x0 = vr[i], x1 = vr[i - 1], x2 = vr[i - 2], x3 = vr[i - 3];
                            vr[i + 1] = fma(c0.x, w, c0.y * x0); w = fma(-c0.y, w, c0.x * x0);
                            vr[i] = fma(c1.x, w, c1.y * x1);     w = fma(-c1.y, w, c1.x * x1);
                            vr[i - 1] = fma(c2.x, w, c2.y * x2); w = fma(-c2.y, w, c2.x * x2);
                            vr[i - 2] = fma(c3.x, w, c3.y * x3); w = fma(-c3.y, w, c3.x * x3);
                        }
                        for (; j < cnt; ++j) {
                            const int i = mtop - 1 - j;
                            const double2 c_s = cs[j];
                            const double x = vr[i];
                            vr[i + 1] = fma(c_s.x, w, c_s.y * x);
                            w = fma(-c_s.y, w, c_s.x * x);
                        }
                        vr[lp] = w;
                    }
                }
                h = hn; it0 = it0n;
            }
            __syncthreads();
            for (int i = tid; i < d; i += nth) wdiag[i] = a.bs.ev[(size_t)b * d + i];
        } else if (pv.tridiag) {
            for (int m = lane; m < d; m += 32) {
                double acc = 0.0;
                for (int k = 0; k < f; ++k)
                    acc = acc + (pv.omegas[k] * pv.occ[m * f + k] + (0.5 * chi[k]) * (pv.occ[m * f + k] * pv.occ[m * f + k]));
                mydiag[m] = acc;
                myoff[m] = (m < d - 1) ? pv.offd[m] : 0.0;
            }
            __syncwarp();
        } else {
            // dense symmetric H into M, then Householder reduction to tridiagonal form with the
            // orthogonal factor accumulated (forward) into V.  Row/col 0 of Q stay e_0.
            for (int idx = tid; idx < d * d; idx += nth) {
                const int r = idx / d, c = idx - r * d;
                double h = pv.offd[idx];
                if (r == c) {
                    double acc = 0.0;
                    for (int k = 0; k < f; ++k)
                        acc = acc + (pv.omegas[k] * pv.occ[r * f + k] + (0.5 * chi[k]) * (pv.occ[r * f + k] * pv.occ[r * f + k]));
                    h = acc;
                }
                M[r * ld + c] = h;
            }
            __syncthreads();
            for (int k = 0; k < d - 2; ++k) {
                const int n = d - k - 1;            // length of the column below the diagonal
                const int o = k + 1;
                double part = 0.0;
                for (int i = 1 + tid; i < n; i += nth) { const double x = M[(o + i) * ld + k]; part += x * x; }
                const double sigma = block_sum(part, red, tid, nwarps);
                const double x0 = M[o * ld + k];
                if (sigma == 0.0) {                 // nothing to annihilate
                    for (int w = tid; w < nwarps; w += nth) { wdiag[w * d + k] = M[k * ld + k]; woff[w * d + k] = x0; }
                    __syncthreads();
                    continue;
                }
                const double mu = sqrt(x0 * x0 + sigma);
                const double alpha = (x0 >= 0.0) ? -mu : mu;
                const double v0 = x0 - alpha;
                const double beta = 2.0 / (v0 * v0 + sigma);
                for (int i = tid; i < n; i += nth) wr[i] = (i == 0) ? v0 : M[(o + i) * ld + k];
                __syncthreads();
                // p = beta * A22 v
                double vp = 0.0;
                for (int i = tid; i < n; i += nth) {
                    const double* row = M + (size_t)(o + i) * ld + o;
                    double acc = 0.0;
                    for (int j = 0; j < n; ++j) acc = fma(row[j], wr[j], acc);
                    acc *= beta;
                    wi[i] = acc;
                    vp = fma(wr[i], acc, vp);
                }
                const double kk = 0.5 * beta * block_sum(vp, red, tid, nwarps);
                for (int i = tid; i < n; i += nth) wi[i] = wi[i] - kk * wr[i];
                __syncthreads();
                // A22 -= v w^T + w v^T ;  Q[:, o:] -= beta (Q[:, o:] v) v^T
                for (int i = tid; i < n; i += nth) {
                    double* row = M + (size_t)(o + i) * ld + o;
                    const double vi = wr[i], wi_ = wi[i];
                    for (int j = 0; j < n; ++j) row[j] -= vi * wi[j] + wi_ * wr[j];
                }
                for (int r = tid; r < d; r += nth) {
                    double* row = V + (size_t)r * ld + o;
                    double u = 0.0;
                    for (int j = 0; j < n; ++j) u = fma(row[j], wr[j], u);
                    u *= beta;
                    for (int j = 0; j < n; ++j) row[j] = fma(-u, wr[j], row[j]);
                }
                for (int w = tid; w < nwarps; w += nth) { wdiag[w * d + k] = M[k * ld + k]; woff[w * d + k] = alpha; }
                __syncthreads();
            }
            for (int w = tid; w < nwarps; w += nth) {
                if (d >= 2) {
                    wdiag[w * d + d - 2] = M[(d - 2) * ld + d - 2];
                    woff[w * d + d - 2] = M[(d - 1) * ld + d - 2];
                }
                wdiag[w * d + d - 1] = M[(d - 1) * ld + d - 1];
                woff[w * d + d - 1] = 0.0;
            }
            __syncthreads();
            if (a.mode == 1) {
                for (int i = tid; i < d; i += nth) { a.bs.tri_d[(size_t)b * d + i] = wdiag[i]; a.bs.tri_e[(size_t)b * d + i] = woff[i]; }
                double* q = a.bs.qmat + (size_t)b * d * d;
                for (int idx = tid; idx < d * d; idx += nth) {
                    const int r = idx / d, c = idx - r * d;
                    q[idx] = V[r * ld + c];
                }
                continue;
            }
        }

        // ---------------------------------------------------------------- implicit QL
        // Scalar recurrences are computed redundantly by every lane on the warp-private copy;
        // each thread rotates its own rows of V.  (Algorithm: implicit-shift QL with Wilkinson shift.)
        if (a.mode == 0) {
            double* dg = mydiag;
            double* of = myoff;
            for (int l = 0; l < d; ++l) {
                int iter = 0;
                while (true) {
                    int m = l;
                    for (; m < d - 1; ++m) {
                        const double dd = fabs(dg[m]) + fabs(dg[m + 1]);
                        if (fabs(of[m]) <= kEps * dd) break;
                    }
                    if (m == l) break;
                    if (++iter > kMaxQlIter) { if (tid == 0) atomicAdd(pv.stats + 2, 1ULL); break; }   // surfaced through qtet_stats
                    double g = (dg[l + 1] - dg[l]) / (2.0 * of[l]);
                    double r = sqrt(g * g + 1.0);
                    g = dg[m] - dg[l] + of[l] / (g + copysign(r, g));
                    double s = 1.0, c = 1.0, p = 0.0;
                    int i = m - 1;
                    bool underflow = false;
                    __syncwarp();
                    for (; i >= l; --i) {
                        const double fo = s * of[i];
                        const double bo = c * of[i];
                        r = sqrt(fo * fo + g * g);
                        __syncwarp();
                        if (lane == 0) of[i + 1] = r;
                        if (r == 0.0) {
                            if (lane == 0) { dg[i + 1] -= p; of[m] = 0.0; }
                            underflow = true;
                            break;
                        }
                        const double rinv = 1.0 / r;
                        s = fo * rinv;
                        c = g * rinv;
                        g = dg[i + 1] - p;
                        r = (dg[i] - g) * s + 2.0 * c * bo;
                        p = s * r;
                        __syncwarp();
                        if (lane == 0) dg[i + 1] = g + p;
                        g = c * r - bo;
                        for (int row = tid; row < d; row += nth) {
                            double* vr = V + (size_t)row * ld;
                            const double y = vr[i + 1], x = vr[i];
                            vr[i + 1] = fma(s, x, c * y);
                            vr[i] = fma(c, x, -(s * y));
                        }
                    }
                    __syncwarp();
                    if (underflow) continue;
                    if (lane == 0) { dg[l] -= p; of[l] = g; of[m] = 0.0; }
                    __syncwarp();
                }
            }
        }
        __syncthreads();
        const double* ev = wdiag;           // warp 0's copy (all identical)
        for (int i = tid; i < d; i += nth) cvec[i] = V[i];    // c = V[0,:]
        __syncthreads();

        // ---------------------------------------------------------------- evolution + loss
        for (int k = 0; k < T; ++k) {
            const double t = pv.tgrid[k];
            for (int i = tid; i < d; i += nth) {
                double sn, cs;
                sincos(ev[i] * t, &sn, &cs);
                zr[i] = cvec[i] * cs;
                zi[i] = -cvec[i] * sn;
            }
            __syncthreads();
            double part = 0.0;
            for (int j = tid; j < d; j += nth) {
                const double* row = V + (size_t)j * ld;
                double pr = 0.0, pi = 0.0;
                for (int i = 0; i < d; ++i) { pr = fma(row[i], zr[i], pr); pi = fma(row[i], zi[i], pi); }
                part = fma(pv.occ[j * f + a.site], pr * pr + pi * pi, part);
            }
            const double nk = block_sum(part, red, tid, nwarps);
            if (tid == 0) tr[k] = nk;
        }
        __syncthreads();
        if (a.trace != nullptr)
            for (int k = tid; k < T; k += nth) a.trace[b * T + k] = tr[k];
        if (tid == 0) {
            int best = 0;
            double bv = tr[0];
            for (int k = 1; k < T; ++k) {
                const bool better = a.acceptor ? (tr[k] > bv) : (tr[k] < bv);
                if (better) { bv = tr[k]; best = k; }
            }
            ired[0] = best;
            if (a.loss) a.loss[b] = a.acceptor ? ((double)pv.N - bv) : bv;
            if (a.tstar) a.tstar[b] = best;
        }
        __syncthreads();
        if (a.grad == nullptr) continue;

        // ---------------------------------------------------------------- gradient
        const int ks = ired[0];
        const double ts = pv.tgrid[ks];
        for (int i = tid; i < d; i += nth) {
            double sn, cs;
            sincos(ev[i] * ts, &sn, &cs);
            fr[i] = cs; fi[i] = -sn;
            zr[i] = cvec[i] * cs; zi[i] = -cvec[i] * sn;
        }
        __syncthreads();
        for (int j = tid; j < d; j += nth) {
            const double* row = V + (size_t)j * ld;
            double pr = 0.0, pi = 0.0;
            for (int i = 0; i < d; ++i) { pr = fma(row[i], zr[i], pr); pi = fma(row[i], zi[i], pi); }
            const double w = pv.occ[j * f + a.site];
            wr[j] = w * pr;            // w conj(psi)
            wi[j] = -w * pi;
        }
        __syncthreads();
        for (int i = tid; i < d; i += nth) {
            double sr = 0.0, si = 0.0;
            for (int j = 0; j < d; ++j) { const double v = V[(size_t)j * ld + i]; sr = fma(v, wr[j], sr); si = fma(v, wi[j], si); }
            ar[i] = sr; ai[i] = si;
        }
        __syncthreads();
        // K_ij = Re(a_i c_j Phi_ij)
        for (int idx = tid; idx < d * d; idx += nth) {
            const int i = idx / d, j = idx - i * d;
            double pr, pi;
            if (i == j) { pr = ts * fi[i]; pi = -ts * fr[i]; }
            else dk_phi(ev[i], fr[i], fi[i], ev[j], fr[j], fi[j], ts, pr, pi);
            M[i * ld + j] = cvec[j] * (ar[i] * pr - ai[i] * pi);
        }
        __syncthreads();
        double gacc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) gacc[k] = 0.0;
        for (int m = tid; m < d; m += nth) {
            const double* row = V + (size_t)m * ld;
            double G = 0.0;
            int i = 0;
            // four rows of K per pass over my row of V: 5 shared-memory loads per 4 fma instead of 8
            for (; i + 4 <= d; i += 4) {
                const double* k0 = M + (size_t)i * ld;
                const double* k1 = k0 + ld;
                const double* k2 = k1 + ld;
                const double* k3 = k2 + ld;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                for (int j = 0; j < d; ++j) {
                    const double vj = row[j];
                    a0 = fma(k0[j], vj, a0); a1 = fma(k1[j], vj, a1); a2 = fma(k2[j], vj, a2); a3 = fma(k3[j], vj, a3);
                }
                G = fma(row[i], a0, G); G = fma(row[i + 1], a1, G); G = fma(row[i + 2], a2, G); G = fma(row[i + 3], a3, G);
            }
            for (; i < d; ++i) {
                const double* kr = M + (size_t)i * ld;
                double acc = 0.0;
                for (int j = 0; j < d; ++j) acc = fma(kr[j], row[j], acc);
                G = fma(row[i], acc, G);
            }
            G *= 2.0;
            for (int k = 0; k < f && k < 8; ++k) gacc[k] = fma(pv.quad[m * f + k], G, gacc[k]);
        }
        for (int k = 0; k < f && k < 8; ++k) {
            const double s = block_sum(gacc[k], red, tid, nwarps);
            if (tid == 0) a.grad[b * f + k] = a.acceptor ? -s : s;
        }
    }
}

__global__ void build_h_kernel(ProblemView pv, const double* chi, double* H) {
    const int d = pv.d, f = pv.f;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < d * d; idx += gridDim.x * blockDim.x) {
        const int r = idx / d, c = idx - r * d;
        double h = 0.0;
        if (r == c) {
            double acc = 0.0;
            for (int k = 0; k < f; ++k)
                acc = acc + (pv.omegas[k] * pv.occ[r * f + k] + (0.5 * chi[k]) * (pv.occ[r * f + k] * pv.occ[r * f + k]));
            h = acc;
        } else if (pv.tridiag) {
            if (r == c + 1) h = pv.offd[c];
            else if (c == r + 1) h = pv.offd[r];
        } else {
            h = pv.offd[idx];
        }
        H[idx] = h;
    }
}

}  // namespace

int generic_threads(int d) {
    int t = ((d + 31) / 32) * 32;
    if (t < 32) t = 32;
    if (t > 128) t = 128;
    return t;
}

static int generic_ld(int d) { return d | 1; }

size_t generic_smem_bytes(int d, int nthreads) {
    const int nw = nthreads / 32;
    const size_t ld = generic_ld(d);
    size_t n = 2 * (size_t)d * ld + 2 * (size_t)nw * d + 9 * (size_t)d + 64 /*tr, up to T=64*/ + 32 + 2;
    return n * sizeof(double);
}

// Per-device slab for the global-memory mode (grow-only, lives as long as the process: the launchers here have no handle
// to hang it on, and the mode is the last-resort path of very large Fock spaces).
static double* generic_slab(int dev, size_t bytes) {
    static std::mutex mu;
    static double* ptr[64] = {};
    static size_t have[64] = {};
    std::lock_guard<std::mutex> lock(mu);
    dev &= 63;
    if (bytes > have[dev]) {
        if (ptr[dev]) { cudaDeviceSynchronize(); cudaFree(ptr[dev]); }
        ptr[dev] = nullptr; have[dev] = 0;
        if (cudaMalloc(&ptr[dev], bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        have[dev] = bytes;
    }
    return ptr[dev];
}

size_t generic_small_smem_bytes(int d, int nthreads, int T) {      // everything except V and M
    const int nw = nthreads / 32;
    return (2 * (size_t)nw * d + 9 * (size_t)d + (size_t)(T > 64 ? T : 64) + 32 + 2) * sizeof(double);
}

static int launch_generic_impl(const ProblemView& pv, const double* chis, int64_t B, const int* list, const int* count,
                               int site, double* loss, double* grad, int32_t* tstar, double* trace, cudaStream_t st,
                               int mode = 0, const BigStream* bs = nullptr) {
    if (B <= 0) return QTET_OK;
    if (pv.f > 8) { set_error("generic kernel supports at most 8 sites"); return QTET_EUNSUPPORTED; }
    const int nth = generic_threads(pv.d);
    size_t smem = generic_smem_bytes(pv.d, nth) + (pv.T > 64 ? (pv.T - 64) * sizeof(double) : 0);
    int dev = 0, max_optin = 0, sms = 0;
    QTET_CUDA(cudaGetDevice(&dev));
    QTET_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    QTET_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    bool use_slab = false;
    if (smem > (size_t)max_optin) {
        // V and M move to global memory; the vectors stay in shared memory
        smem = generic_small_smem_bytes(pv.d, nth, pv.T);
        use_slab = true;
        if (smem > (size_t)max_optin || mode != 0) {
            set_error("Fock dimension " + std::to_string(pv.d) + " needs " + std::to_string(smem) +
                      " B of shared memory per CTA even with the matrices in global memory; device limit is " + std::to_string(max_optin));
            return QTET_EUNSUPPORTED;
        }
    }
    QTET_CUDA(cudaFuncSetAttribute(generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    QTET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, generic_kernel, nth, smem));
    if (per_sm < 1) per_sm = 1;
    if (use_slab && per_sm > 4) per_sm = 4;
    long long grid = (long long)sms * per_sm;     // persistent: a whole number of waves of the 148 SMs
    if (grid > B) grid = B;
    GenericArgs a;
    a.pv = pv; a.chis = chis; a.B = B; a.site = site; a.acceptor = (site == pv.f - 1) ? 1 : 0;
    a.ldv = generic_ld(pv.d); a.loss = loss; a.grad = grad; a.tstar = tstar; a.trace = trace;
    a.list = list; a.count = count;
    a.mode = mode;
    if (bs) a.bs = *bs; else a.bs = BigStream{};
    if (list) { grid = 2LL * sms; if (grid > B) grid = B; }
    if (use_slab) {
        const size_t bytes = (size_t)grid * 2 * pv.d * a.ldv * sizeof(double);
        a.slab = generic_slab(dev, bytes);
        if (!a.slab) { set_error("cudaMalloc(generic-kernel slab, " + std::to_string(bytes) + " B) failed"); return QTET_ENOMEM; }
    }
    generic_kernel<<<(unsigned)grid, nth, smem, st>>>(a);
    count_launch();
    QTET_CUDA(cudaGetLastError());
    return QTET_OK;
}

int launch_generic(const ProblemView& pv, const double* chis, int64_t B, int site, double* loss,
                   double* grad, int32_t* tstar, double* trace, cudaStream_t st, Profiler* prof) {
    if (prof) prof->begin(2, B, st);
    const int rc = launch_generic_impl(pv, chis, B, nullptr, nullptr, site, loss, grad, tstar, trace, st);
    if (prof) prof->finish(2, st);
    return rc;
}

// Evaluate only the points list[0 .. *count) (count read on the device; at most max_count).
int launch_generic_list(const ProblemView& pv, const double* chis, const int* list, const int* count, long long max_count,
                        int site, double* loss, double* grad, int32_t* tstar, double* trace, cudaStream_t st) {
    return launch_generic_impl(pv, chis, max_count, list, count, site, loss, grad, tstar, trace, st);
}

int launch_generic_mode(const ProblemView& pv, const double* chis, int64_t B, int site, double* loss, double* grad,
                        int32_t* tstar, double* trace, cudaStream_t st, int mode, const BigStream& bs) {
    return launch_generic_impl(pv, chis, B, nullptr, nullptr, site, loss, grad, tstar, trace, st, mode, &bs);
}

int launch_build_h(const ProblemView& pv, const double* chis, double* H, cudaStream_t st) {
    build_h_kernel<<<8, 256, 0, st>>>(pv, chis, H);
    count_launch();
    QTET_CUDA(cudaGetLastError());
    return QTET_OK;
}

}  // namespace qtet
