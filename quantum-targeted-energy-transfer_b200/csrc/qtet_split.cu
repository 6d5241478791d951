// Fast path for tridiagonal Hamiltonians with Fock dimension d <= 32 (every dimer up to N = 31;
// BASELINE's headline is the dimer N = 20, d = 21).
//
// The eigendecomposition is split by the kind of parallelism it offers:
//
//   ql_scalar_kernel  (K_A)  one THREAD per parameter point.  Implicit-shift QL on the tridiagonal
//                     (diag, off) only -- O(d^2) scalar work with sqrt/div chains that would waste a
//                     whole warp if done per warp.  Emits the eigenvalues and the sequence of plane
//                     rotations (c, s), sweep by sweep, into a per-point stream in HBM/L2.
//   apply_kernel      (K_B)  one WARP per parameter point, lane j owns row j of V in REGISTERS.
//                     Replays the rotation stream on V = I (4 DFMA per rotation per lane, the
//                     O(d^3) part), then the time evolution psi = V (c . exp(-i e t)) with phases
//                     advanced by a complex rotation recurrence, <n(t)>, the loss, and the analytic
//                     gradient through the symmetrised Daleckii-Krein kernel.
//
// Points whose stream overflows its slot (pathological convergence) are appended to a list and
// recomputed by the fused generic kernel; no host synchronisation is involved.
#include "qtet_common.cuh"

namespace qtet {

namespace {

constexpr double kEps = 1.1102230246251565e-16;
constexpr int kMaxSweeps = 128;            // header slots per point
constexpr unsigned kTerm = 0xffffffffu;
constexpr int kWarpsA = 4;                 // warps per CTA, scalar kernel
constexpr int kWarpsB = 4;                 // warps per CTA, apply kernel

struct SplitArgs {
    ProblemView pv;
    const double* chis;       // [n, f] (already offset to the chunk)
    long long n;              // points in this chunk
    int site, acceptor;
    int cap;                  // rotation records per point
    double2* stream;          // [n, cap]
    unsigned* hdr;            // [n, kMaxSweeps]
    double* ev;               // [n, d]
    int* overflow_list;       // [n]
    int* overflow_count;      // [1]
    const unsigned short* pair_tab;   // [D(D-1)/2] i | j << 8
    double* loss; double* grad; int* tstar; double* trace;   // chunk-offset outputs
};

// ------------------------------------------------------------------------------------------------
// K_A: thread-per-point scalar QL.  diag/off live in shared memory as [k][lane] (bank = lane, so
// lanes may sit at different k without conflicts).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsA * 32) ql_scalar_kernel(SplitArgs a) {
    extern __shared__ double sm[];
    const ProblemView& pv = a.pv;
    const int d = pv.d, f = pv.f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* sd = sm + (size_t)warp * 2 * d * 32;
    double* se = sd + d * 32;
    const long long nbatch = (a.n + 31) / 32;
    for (long long batch = (long long)blockIdx.x * kWarpsA + warp; batch < nbatch; batch += (long long)gridDim.x * kWarpsA) {
        const long long p = batch * 32 + lane;
        const bool valid = p < a.n;
        const long long pc = valid ? p : a.n - 1;
        // ---- H build (HamiltonianLoss.py:131-134) + global-phase shift by the mean of the diagonal
        double mean = 0.0;
        for (int m = 0; m < d; ++m) {
            double acc = 0.0;
            for (int k = 0; k < f; ++k) {
                const double n = pv.occ[m * f + k];
                acc = acc + (pv.omegas[k] * n + (0.5 * a.chis[pc * f + k]) * (n * n));
            }
            sd[m * 32 + lane] = acc;
            se[m * 32 + lane] = (m < d - 1) ? pv.offd[m] : 0.0;
            mean += acc;
        }
        mean /= (double)d;
        for (int m = 0; m < d; ++m) sd[m * 32 + lane] -= mean;

        double2* rec = a.stream + pc * a.cap;
        unsigned* hdr = a.hdr + pc * kMaxSweeps;
        int pos = 0, ns = 0;
        bool overflow = false;
        // bit k of `small`: off[k] is negligible against its diagonal neighbours (bit d-1: sentinel).
        // Initialised by one scan, then kept current inside the sweeps (only swept entries change).
        unsigned small = 1u << (d - 1);
        for (int k = 0; k < d - 1; ++k) {
            const double dd = fabs(sd[k * 32 + lane]) + fabs(sd[(k + 1) * 32 + lane]);
            if (fabs(se[k * 32 + lane]) <= kEps * dd) small |= 1u << k;
        }
        int l = 0, iter = 0;
        while (l < d) {
            const int m = l + __ffs(small >> l) - 1;          // first negligible off-diagonal at or above l
            if (m == l) { ++l; iter = 0; continue; }
            if (++iter > 60 || ns >= kMaxSweeps - 1 || pos + (m - l) > a.cap) { overflow = true; break; }
            if (valid) hdr[ns] = (unsigned)l | ((unsigned)m << 8) | ((unsigned)pos << 16);   // pos < cap < 65536
            ++ns;
            const double dl = sd[l * 32 + lane], el = se[l * 32 + lane];
            double g = (sd[(l + 1) * 32 + lane] - dl) / (2.0 * el);
            double r = sqrt(fma(g, g, 1.0));
            g = sd[m * 32 + lane] - dl + el / (g + copysign(r, g));
            double s = 1.0, c = 1.0, pp = 0.0;
            double dnext = 0.0;                                // final d[i+2] while finishing plane i
            int i = m - 1;
            bool broke = false;
            for (; i >= l; --i) {
                const double ei = se[i * 32 + lane];
                const double fo = s * ei, bo = c * ei;
                const double r2 = fma(fo, fo, g * g);
                if (r2 == 0.0) {                               // underflow recovery: finish the sweep with identities
                    se[(i + 1) * 32 + lane] = 0.0;
                    sd[(i + 1) * 32 + lane] -= pp;
                    se[m * 32 + lane] = 0.0;
                    broke = true;
                    break;
                }
                const double rinv = rsqrt(r2);
                r = r2 * rinv;
                s = fo * rinv;
                c = g * rinv;
                const double di1 = sd[(i + 1) * 32 + lane];
                g = di1 - pp;
                const double rr = fma(sd[i * 32 + lane] - g, s, 2.0 * c * bo);
                pp = s * rr;
                const double dnew = g + pp;
                sd[(i + 1) * 32 + lane] = dnew;
                g = fma(c, rr, -bo);
                if (i + 1 < m) {                               // off[i+1] = r is final: refresh its flag
                    se[(i + 1) * 32 + lane] = r;
                    const unsigned bit = 1u << (i + 1);
                    if (r <= kEps * (fabs(dnew) + fabs(dnext))) small |= bit; else small &= ~bit;
                }
                dnext = dnew;
                if (valid) rec[pos] = make_double2(c, s);
                ++pos;
            }
            if (broke) {
                for (; i >= l; --i) { if (valid) rec[pos] = make_double2(1.0, 0.0); ++pos; }
                // rare path: rebuild the flags from scratch
                small = 1u << (d - 1);
                for (int k = 0; k < d - 1; ++k) {
                    const double dd = fabs(sd[k * 32 + lane]) + fabs(sd[(k + 1) * 32 + lane]);
                    if (fabs(se[k * 32 + lane]) <= kEps * dd) small |= 1u << k;
                }
                continue;
            }
            const double dlnew = sd[l * 32 + lane] - pp;
            sd[l * 32 + lane] = dlnew;
            se[l * 32 + lane] = g;
            se[m * 32 + lane] = 0.0;
            small |= 1u << m;
            {
                const unsigned bit = 1u << l;
                if (fabs(g) <= kEps * (fabs(dlnew) + fabs(dnext))) small |= bit; else small &= ~bit;
            }
        }
        if (valid) {
            hdr[ns] = kTerm;
            if (overflow) {
                hdr[0] = kTerm;
                const int slot = atomicAdd(a.overflow_count, 1);
                a.overflow_list[slot] = (int)p;
            }
            for (int m = 0; m < d; ++m) a.ev[p * d + m] = sd[m * 32 + lane];
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// K_B: warp-per-point.  D = compile-time padded dimension (rows/cols >= d stay identity).
// ------------------------------------------------------------------------------------------------
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

constexpr int kRing = 4;                   // sweeps of rotation records in flight per warp

template <int D>
__global__ void __launch_bounds__(kWarpsB * 32, 5) apply_kernel(SplitArgs a, int smem_per_warp_doubles) {
    extern __shared__ double sm[];
    const ProblemView& pv = a.pv;
    const int d = pv.d, f = pv.f, T = pv.T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* base = sm + (size_t)warp * smem_per_warp_doubles;
    double2* rbuf = reinterpret_cast<double2*>(base);          // [kRing][32]  staged rotations by plane index
    double2* zbuf = rbuf + kRing * 32;                          // [2][32]  z_i = c_i f_i^k                                  // [2][32]  z_i = c_i f_i^k
    double* lanev = reinterpret_cast<double*>(zbuf + 64);       // [8][32]  per-eigen-index vectors for the gradient
    unsigned* hbuf = reinterpret_cast<unsigned*>(lanev + 8 * 32);   // [kMaxSweeps + kRing] sweep headers
    double* U = lanev + 8 * 32 + (kMaxSweeps + 8) / 2;                                 // union: p[T][LDU] | V^T staging [D][LDU] | S packed
    constexpr int LDU = D | 1;                                  // odd row pitch: conflict-free column reads

    const double wsite = (lane < d) ? pv.occ[lane * f + a.site] : 0.0;
    const double dt = (T > 1) ? pv.tgrid[1] : 0.0;

    for (long long p = (long long)blockIdx.x * kWarpsB + warp; p < a.n; p += (long long)gridDim.x * kWarpsB) {
        const unsigned* hdr = a.hdr + p * kMaxSweeps;
        unsigned H0 = hdr[lane], H1 = hdr[32 + lane], H2 = hdr[64 + lane], H3 = hdr[96 + lane];
        const double2* rec = a.stream + p * a.cap;
        double v[D];
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] = (k == lane) ? 1.0 : 0.0;

        // ---- replay the rotation stream.  Records travel global -> shared with cp.async, kRing-1 sweeps
        // ahead of their use (each lane copies one record of a sweep straight to its plane slot), so the
        // HBM/L2 latency of the stream is off the critical path.  Sweep headers (l | m<<8 | pos<<16) are
        // staged in shared memory once per point.
        hbuf[lane] = H0; hbuf[32 + lane] = H1; hbuf[64 + lane] = H2; hbuf[96 + lane] = H3;
        if (lane < kRing) hbuf[kMaxSweeps + lane] = kTerm;
        __syncwarp();
        auto issue = [&](int s) {
            const unsigned hh = hbuf[s];
            if (hh != kTerm) {
                const int li = hh & 0xff, mi = (hh >> 8) & 0xff, pi = hh >> 16;
                if (lane < mi - li) cp_async16(rbuf + (s % kRing) * 32 + (mi - 1 - lane), rec + pi + lane);
            }
            cp_async_commit();
        };
#pragma unroll
        for (int s = 0; s < kRing - 1; ++s) issue(s);
        unsigned h = hbuf[0];
        for (int s = 0; h != kTerm; ++s) {
            const int l = h & 0xff, m = (h >> 8) & 0xff;
            __syncwarp();                                   // everyone is done with the slot about to be refilled
            issue(s + kRing - 1);
            h = hbuf[s + 1];
            cp_async_wait<kRing - 1>();                     // this lane's records of sweep s have landed
            double2* rb = rbuf + (s % kRing) * 32;
            // identity outside the window [l, m): above an interior split, and below l down to the group edge
            if (lane < D - 1 && (lane >= m || lane < l)) rb[lane] = make_double2(1.0, 0.0);
            __syncwarp();
            // Planes D-2 .. l, top-down, in groups of four with ONE warp-uniform exit test per group (QL
            // deflates from the top, so the window is almost always the suffix [l, d-2]).  Inside a group the
            // code is branch-free: the four (c, s) loads are issued ahead of the dependent chain and only one
            // fma per plane sits on that chain:  v[i] <- c x - s y  with  c x  computed off-chain.
#pragma unroll
            for (int i0 = D - 2; i0 >= 0; i0 -= 4) {
                if (i0 < l) break;
                double2 cs[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) cs[k] = rb[(i0 - k >= 0) ? (i0 - k) : 0];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i = i0 - k;
                    if (i >= 0) {
                        const double x = v[i], y = v[i + 1];
                        const double t1 = cs[k].x * x, t2 = cs[k].y * x;
                        v[i] = fma(-cs[k].y, y, t1);
                        v[i + 1] = fma(cs[k].x, y, t2);
                    }
                }
            }
        }
        cp_async_wait<0>();

        // ---- c = V[0,:], eigenvalues, unit phase step w = exp(-i e dt)
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < D; ++k) lanev[k] = v[k];
        }
        __syncwarp();
        const double ci = (lane < D) ? lanev[lane] : 0.0;
        const double ei = (lane < d) ? a.ev[p * d + lane] : 0.0;
        double wre, wim;
        {
            const double th = ei * dt;
            const double lo = fma(ei, dt, -th);             // exact product tail
            double sn, cs;
            sincos(th, &sn, &cs);
            // exp(-i (th + lo)) = (cs - i sn)(1 - i lo)
            wre = fma(-sn, lo, cs);
            wim = -fma(cs, lo, sn);
        }
        __syncwarp();

        // ---- time evolution: p_j(k) = n_site(j) |psi_j(t_k)|^2 staged as U[k][j]
        double zr = ci, zi = 0.0;
        for (int k = 0; k < T; ++k) {
            double2* zb = zbuf + (k & 1) * 32;
            zb[lane] = make_double2(zr, zi);
            __syncwarp();
            double ar4[4] = {0.0, 0.0, 0.0, 0.0}, ai4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const double2 z0 = zb[i];
                ar4[i & 3] = fma(v[i], z0.x, ar4[i & 3]);
                ai4[i & 3] = fma(v[i], z0.y, ai4[i & 3]);
            }
            const double pr = (ar4[0] + ar4[1]) + (ar4[2] + ar4[3]), pi = (ai4[0] + ai4[1]) + (ai4[2] + ai4[3]);
            if (lane < D) U[k * LDU + lane] = wsite * fma(pr, pr, pi * pi);
            const double nr = fma(zr, wre, -(zi * wim));
            zi = fma(zr, wim, zi * wre);
            zr = nr;
        }
        __syncwarp();
        // lane k sums step k over rows; arg-max / arg-min with the first index winning ties
        int kbest = 0;
        double vbest = 0.0;
        for (int k0 = 0; k0 < T; k0 += 32) {
            const int k = k0 + lane;
            double nk = 0.0;
            if (k < T) {
#pragma unroll
                for (int j = 0; j < D; ++j) nk += U[k * LDU + j];
                if (a.trace) a.trace[p * T + k] = nk;
            }
            double key = (k < T) ? (a.acceptor ? nk : -nk) : -1.0e300;
            int kk = k;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ok = __shfl_xor_sync(0xffffffffu, key, o);
                const int oi = __shfl_xor_sync(0xffffffffu, kk, o);
                if (ok > key || (ok == key && oi < kk)) { key = ok; kk = oi; }
            }
            if (k0 == 0 || key > vbest) { vbest = key; kbest = kk; }
        }
        if (lane == 0) {
            if (a.loss) a.loss[p] = a.acceptor ? ((double)pv.N - vbest) : -vbest;
            if (a.tstar) a.tstar[p] = kbest;
        }
        __syncwarp();
        if (a.grad == nullptr) continue;

        // ---- gradient at t* = t_kbest: f_i = w_i^kbest by binary powering
        double fr = 1.0, fi = 0.0;
        {
            double br = wre, bi = wim;
            for (int e = kbest; e > 0; e >>= 1) {
                if (e & 1) { const double t = fma(fr, br, -(fi * bi)); fi = fma(fr, bi, fi * br); fr = t; }
                const double t2 = fma(br, br, -(bi * bi)); bi = 2.0 * br * bi; br = t2;
            }
        }
        const double ts = pv.tgrid[kbest];
        {
            double2* zb = zbuf;
            zb[lane] = make_double2(ci * fr, ci * fi);
        }
        __syncwarp();
        double pr = 0.0, pi = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) { const double2 z = zbuf[i]; pr = fma(v[i], z.x, pr); pi = fma(v[i], z.y, pi); }
        // w conj(psi) per row, then a = V^T (w conj psi) through a transposed copy of V in shared memory
        double2* wc = zbuf + 32;
        wc[lane] = make_double2(wsite * pr, -wsite * pi);
        if (lane < D) {
#pragma unroll
            for (int i = 0; i < D; ++i) U[lane * LDU + i] = v[i];
        }
        __syncwarp();
        double ar = 0.0, ai = 0.0;
        const int lcol = (lane < D) ? lane : 0;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const double vt = U[j * LDU + lcol];          // V[j][lane]
            const double2 w = wc[j];
            ar = fma(vt, w.x, ar); ai = fma(vt, w.y, ai);
        }
        __syncwarp();
        // per-eigen-index vectors: e, c, f, a, b = a f
        const double bre = fma(ar, fr, -(ai * fi)), bim = fma(ar, fi, ai * fr);
        lanev[0 * 32 + lane] = ei; lanev[1 * 32 + lane] = ci;
        lanev[2 * 32 + lane] = fr; lanev[3 * 32 + lane] = fi;
        lanev[4 * 32 + lane] = ar; lanev[5 * 32 + lane] = ai;
        lanev[6 * 32 + lane] = bre; lanev[7 * 32 + lane] = bim;
        // S packed: diagonal K_ii at U[i], pairs (i<j) at U[32 + q]
        U[lane] = ci * ts * bim;
        __syncwarp();
        constexpr int NP = D * (D - 1) / 2;
        for (int q = lane; q < NP; q += 32) {
            const unsigned ij = a.pair_tab[q];
            const int i = ij & 0xff, j = ij >> 8;
            const double e_i = lanev[i], e_j = lanev[j];
            const double c_i = lanev[32 + i], c_j = lanev[32 + j];
            const double dl = e_i - e_j;
            double sij;
            if (fabs(dl * ts) < 1e-2) {
                // Phi_ij = ts * f_j * (exp(-i x) - 1)/x ; S = Re[(a_i c_j + a_j c_i) Phi_ij]
                double gr, gi;
                expm1_over_x(dl * ts, gr, gi);
                const double fjr = lanev[64 + j], fji = lanev[96 + j];
                const double phr = ts * (fjr * gr - fji * gi), phi = ts * (fjr * gi + fji * gr);
                const double mr = lanev[128 + i] * c_j + lanev[128 + j] * c_i;
                const double mi = lanev[160 + i] * c_j + lanev[160 + j] * c_i;
                sij = mr * phr - mi * phi;
            } else {
                const double rinv = fast_rcp(dl);
                // K_ij = c_j R (br_i - ar_i fr_j + ai_i fi_j) ; K_ji = -c_i R (br_j - ar_j fr_i + ai_j fi_i)
                const double kij = c_j * (lanev[192 + i] - lanev[128 + i] * lanev[64 + j] + lanev[160 + i] * lanev[96 + j]);
                const double kji = c_i * (lanev[192 + j] - lanev[128 + j] * lanev[64 + i] + lanev[160 + j] * lanev[96 + i]);
                sij = rinv * (kij - kji);
            }
            U[32 + q] = sij;
        }
        __syncwarp();
        double G = 0.0;
        {
            int q = 0;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                double acc = U[i] * v[i];
#pragma unroll
                for (int j = i + 1; j < D; ++j) { acc = fma(U[32 + q], v[j], acc); ++q; }
                G = fma(v[i], acc, G);
            }
        }
        G *= 2.0;
        for (int k = 0; k < f; ++k) {
            const double qk = (lane < d) ? pv.quad[lane * f + k] : 0.0;
            const double sres = warp_sum(qk * G);
            if (lane == 0) a.grad[p * f + k] = a.acceptor ? -sres : sres;
        }
        __syncwarp();
    }
}

template <int D>
int launch_apply(const SplitArgs& a, int sms, cudaStream_t st) {
    const int T = a.pv.T;
    const int urows = (T > D) ? T : D;
    int udoubles = urows * (D | 1);
    const int need_s = 32 + D * (D - 1) / 2;
    if (need_s > udoubles) udoubles = need_s;
    udoubles = (udoubles + 1) & ~1;
    const int per_warp = 2 * kRing * 32 + 128 + 8 * 32 + (kMaxSweeps + 8) / 2 + udoubles;
    const size_t smem = (size_t)per_warp * kWarpsB * sizeof(double);
    QTET_CUDA(cudaFuncSetAttribute(apply_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QTET_CUDA(cudaFuncSetAttribute(apply_kernel<D>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int per_sm = 0;
    QTET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, apply_kernel<D>, kWarpsB * 32, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sms * per_sm;
    const long long need = (a.n + kWarpsB - 1) / kWarpsB;
    if (grid > need) grid = need;
    apply_kernel<D><<<(unsigned)grid, kWarpsB * 32, smem, st>>>(a, per_warp);
    count_launch();
    QTET_CUDA(cudaGetLastError());
    return QTET_OK;
}

int padded_dim(int d) {
    const int sizes[] = {4, 6, 8, 12, 16, 21, 24, 32};
    for (int s : sizes) if (d <= s) return s;
    return -1;
}

}  // namespace

struct SplitWorkspace {
    int device = 0, sms = 0;
    int cap = 0, D = 0;
    long long chunk = 0;
    char* blob = nullptr;            // two chunk buffers (double-buffered between K_A and K_B)
    size_t blob_bytes = 0;
    unsigned short* pair_tab = nullptr;
    int* overflow_count = nullptr;   // [2]
    cudaStream_t aux = nullptr;      // K_A runs here, one chunk ahead of K_B on the caller's stream
    cudaEvent_t ev_fork = nullptr, ev_a[2] = {nullptr, nullptr}, ev_b[2] = {nullptr, nullptr};
};

bool split_eligible(const ProblemView& pv) {
    return pv.tridiag && pv.d >= 2 && pv.d <= 32 && pv.T <= 64 && pv.f <= 8;
}

void split_workspace_free(SplitWorkspace* ws) {
    if (!ws) return;
    if (ws->blob) cudaFree(ws->blob);
    if (ws->pair_tab) cudaFree(ws->pair_tab);
    if (ws->overflow_count) cudaFree(ws->overflow_count);
    if (ws->aux) cudaStreamDestroy(ws->aux);
    if (ws->ev_fork) cudaEventDestroy(ws->ev_fork);
    for (int i = 0; i < 2; ++i) {
        if (ws->ev_a[i]) cudaEventDestroy(ws->ev_a[i]);
        if (ws->ev_b[i]) cudaEventDestroy(ws->ev_b[i]);
    }
    delete ws;
}

// generic kernel restricted to an index list (qtet_generic.cu)
int launch_generic_list(const ProblemView& pv, const double* chis, const int* list, const int* count, long long max_count,
                        int site, double* loss, double* grad, int32_t* tstar, double* trace, cudaStream_t st);

template <int D>
static int apply_dispatch(const SplitArgs& a, int sms, cudaStream_t st) { return launch_apply<D>(a, sms, st); }

int launch_split(const ProblemView& pv, SplitWorkspace** wsp, const double* chis, int64_t B, int site,
                 double* loss, double* grad, int32_t* tstar, double* trace, cudaStream_t st) {
    if (B <= 0) return QTET_OK;
    if (!split_eligible(pv)) { set_error("split path not eligible for this problem"); return QTET_EUNSUPPORTED; }
    SplitWorkspace* ws = *wsp;
    const int d = pv.d;
    if (!ws) {
        ws = new SplitWorkspace();
        *wsp = ws;
        QTET_CUDA(cudaGetDevice(&ws->device));
        QTET_CUDA(cudaDeviceGetAttribute(&ws->sms, cudaDevAttrMultiProcessorCount, ws->device));
        ws->D = padded_dim(d);
        ws->cap = (3 * d * d) / 2 + 64;
        std::vector<unsigned short> tab;
        for (int i = 0; i < ws->D; ++i)
            for (int j = i + 1; j < ws->D; ++j) tab.push_back((unsigned short)(i | (j << 8)));
        if (tab.empty()) tab.push_back(0);
        QTET_CUDA(cudaMalloc(&ws->pair_tab, tab.size() * sizeof(unsigned short)));
        QTET_CUDA(cudaMemcpy(ws->pair_tab, tab.data(), tab.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
        QTET_CUDA(cudaMalloc(&ws->overflow_count, 2 * sizeof(int)));
        QTET_CUDA(cudaStreamCreateWithFlags(&ws->aux, cudaStreamNonBlocking));
        QTET_CUDA(cudaEventCreateWithFlags(&ws->ev_fork, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) {
            QTET_CUDA(cudaEventCreateWithFlags(&ws->ev_a[i], cudaEventDisableTiming));
            QTET_CUDA(cudaEventCreateWithFlags(&ws->ev_b[i], cudaEventDisableTiming));
        }
    }
    // per-point workspace: rotation stream + sweep headers + eigenvalues + overflow slot
    const size_t per_point = (size_t)ws->cap * sizeof(double2) + kMaxSweeps * sizeof(unsigned) + (size_t)d * 8 + 4;
    // chunking: two buffers in flight; aim for >= 8 chunks on big batches so that K_A(c+1) overlaps K_B(c),
    // but keep chunks large enough to fill the machine several times over
    long long chunk = (long long)((size_t)4 << 30) / (long long)per_point;      // <= 4 GiB per buffer
    long long want = (B + 3) / 4;
    if (want < 131072) want = 131072;
    if (chunk > want) chunk = want;
    chunk = ((chunk + 4095) / 4096) * 4096;
    if (chunk > B) chunk = ((B + 31) / 32) * 32;
    const size_t o_hdr = (size_t)chunk * ws->cap * sizeof(double2);
    const size_t o_ev = o_hdr + (size_t)chunk * kMaxSweeps * sizeof(unsigned);
    const size_t o_list = o_ev + (((size_t)chunk * d * 8 + 255) / 256) * 256;
    const size_t buf_bytes = ((o_list + (size_t)chunk * 4 + 255) / 256) * 256;
    const size_t total = 2 * buf_bytes;
    if (total > ws->blob_bytes) {
        if (ws->blob) { cudaDeviceSynchronize(); cudaFree(ws->blob); }
        ws->blob = nullptr; ws->blob_bytes = 0;
        cudaError_t e = cudaMalloc(&ws->blob, total);
        if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(split workspace) failed"); return QTET_ENOMEM; }
        ws->blob_bytes = total;
    }
    ws->chunk = chunk;

    SplitArgs a;
    a.pv = pv; a.site = site; a.acceptor = (site == pv.f - 1) ? 1 : 0; a.cap = ws->cap;
    a.pair_tab = ws->pair_tab;

    const size_t smemA = (size_t)kWarpsA * 2 * d * 32 * sizeof(double);
    QTET_CUDA(cudaFuncSetAttribute(ql_scalar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA));
    QTET_CUDA(cudaFuncSetAttribute(ql_scalar_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int per_sm_a = 0;
    QTET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_a, ql_scalar_kernel, kWarpsA * 32, smemA));
    if (per_sm_a < 1) per_sm_a = 1;

    // fork: the auxiliary stream starts after everything already queued on the caller's stream
    QTET_CUDA(cudaEventRecord(ws->ev_fork, st));
    QTET_CUDA(cudaStreamWaitEvent(ws->aux, ws->ev_fork, 0));
    int c = 0;
    for (long long lo = 0; lo < B; lo += chunk, ++c) {
        const int buf = c & 1;
        char* base = ws->blob + (size_t)buf * buf_bytes;
        const long long n = (B - lo < chunk) ? (B - lo) : chunk;
        a.stream = (double2*)base; a.hdr = (unsigned*)(base + o_hdr);
        a.ev = (double*)(base + o_ev); a.overflow_list = (int*)(base + o_list);
        a.overflow_count = ws->overflow_count + buf;
        a.chis = chis + lo * pv.f; a.n = n;
        a.loss = loss ? loss + lo : nullptr;
        a.grad = grad ? grad + lo * pv.f : nullptr;
        a.tstar = tstar ? tstar + lo : nullptr;
        a.trace = trace ? trace + lo * pv.T : nullptr;
        // K_A(c) on aux, after K_B(c-2) released this buffer
        if (c >= 2) QTET_CUDA(cudaStreamWaitEvent(ws->aux, ws->ev_b[buf], 0));
        QTET_CUDA(cudaMemsetAsync(a.overflow_count, 0, sizeof(int), ws->aux));
        long long gridA = (long long)ws->sms * per_sm_a;
        const long long needA = ((n + 31) / 32 + kWarpsA - 1) / kWarpsA;
        if (gridA > needA) gridA = needA;
        ql_scalar_kernel<<<(unsigned)gridA, kWarpsA * 32, smemA, ws->aux>>>(a);
        count_launch();
        QTET_CUDA(cudaGetLastError());
        QTET_CUDA(cudaEventRecord(ws->ev_a[buf], ws->aux));
        // K_B(c) on the caller's stream
        QTET_CUDA(cudaStreamWaitEvent(st, ws->ev_a[buf], 0));
        int rc;
        switch (ws->D) {
            case 4: rc = apply_dispatch<4>(a, ws->sms, st); break;
            case 6: rc = apply_dispatch<6>(a, ws->sms, st); break;
            case 8: rc = apply_dispatch<8>(a, ws->sms, st); break;
            case 12: rc = apply_dispatch<12>(a, ws->sms, st); break;
            case 16: rc = apply_dispatch<16>(a, ws->sms, st); break;
            case 21: rc = apply_dispatch<21>(a, ws->sms, st); break;
            case 24: rc = apply_dispatch<24>(a, ws->sms, st); break;
            default: rc = apply_dispatch<32>(a, ws->sms, st); break;
        }
        if (rc) return rc;
        // overflowed points (if any) are recomputed by the fused generic kernel
        rc = launch_generic_list(pv, a.chis, a.overflow_list, a.overflow_count, n, site, a.loss, a.grad, a.tstar, a.trace, st);
        if (rc) return rc;
        QTET_CUDA(cudaEventRecord(ws->ev_b[buf], st));
    }
    return QTET_OK;
}

}  // namespace qtet
