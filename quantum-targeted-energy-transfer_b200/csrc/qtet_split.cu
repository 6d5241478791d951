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
        for (int l = 0; l < d && !overflow; ++l) {
            int iter = 0;
            while (true) {
                int m = l;
                for (; m < d - 1; ++m) {
                    const double dd = fabs(sd[m * 32 + lane]) + fabs(sd[(m + 1) * 32 + lane]);
                    if (fabs(se[m * 32 + lane]) <= kEps * dd) break;
                }
                if (m == l) break;
                if (++iter > 60 || ns >= kMaxSweeps - 1 || pos + (m - l) > a.cap) { overflow = true; break; }
                if (valid) hdr[ns] = (unsigned)l | ((unsigned)m << 8);
                ++ns;
                const double dl = sd[l * 32 + lane], el = se[l * 32 + lane];
                double g = (sd[(l + 1) * 32 + lane] - dl) / (2.0 * el);
                double r = sqrt(fma(g, g, 1.0));
                g = sd[m * 32 + lane] - dl + el / (g + copysign(r, g));
                double s = 1.0, c = 1.0, pp = 0.0;
                int i = m - 1;
                bool broke = false;
                for (; i >= l; --i) {
                    const double ei = se[i * 32 + lane];
                    const double fo = s * ei, bo = c * ei;
                    const double r2 = fma(fo, fo, g * g);
                    if (r2 == 0.0) {                      // underflow recovery: finish the sweep with identities
                        se[(i + 1) * 32 + lane] = 0.0;
                        sd[(i + 1) * 32 + lane] -= pp;
                        se[m * 32 + lane] = 0.0;
                        broke = true;
                        break;
                    }
                    const double rinv = rsqrt(r2);
                    r = r2 * rinv;
                    se[(i + 1) * 32 + lane] = r;
                    s = fo * rinv;
                    c = g * rinv;
                    const double di1 = sd[(i + 1) * 32 + lane];
                    g = di1 - pp;
                    r = fma(sd[i * 32 + lane] - g, s, 2.0 * c * bo);
                    pp = s * r;
                    sd[(i + 1) * 32 + lane] = g + pp;
                    g = fma(c, r, -bo);
                    if (valid) rec[pos] = make_double2(c, s);
                    ++pos;
                }
                if (broke) {
                    for (; i >= l; --i) { if (valid) rec[pos] = make_double2(1.0, 0.0); ++pos; }
                    continue;
                }
                sd[l * 32 + lane] -= pp;
                se[l * 32 + lane] = g;
                se[m * 32 + lane] = 0.0;
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

template <int D>
__global__ void __launch_bounds__(kWarpsB * 32) apply_kernel(SplitArgs a, int smem_per_warp_doubles) {
    extern __shared__ double sm[];
    const ProblemView& pv = a.pv;
    const int d = pv.d, f = pv.f, T = pv.T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* base = sm + (size_t)warp * smem_per_warp_doubles;
    double2* rbuf = reinterpret_cast<double2*>(base);          // [2][32]  staged rotations by plane index
    double2* zbuf = rbuf + 64;                                  // [2][32]  z_i = c_i f_i^k
    double* lanev = reinterpret_cast<double*>(zbuf + 64);       // [8][32]  per-eigen-index vectors for the gradient
    double* U = lanev + 8 * 32;                                 // union: p[T][33] | V^T staging [32][33] | S packed
    constexpr int LDU = 33;

    const double wsite = (lane < d) ? pv.occ[lane * f + a.site] : 0.0;
    const double dt = (T > 1) ? pv.tgrid[1] : 0.0;

    for (long long p = (long long)blockIdx.x * kWarpsB + warp; p < a.n; p += (long long)gridDim.x * kWarpsB) {
        const unsigned* hdr = a.hdr + p * kMaxSweeps;
        unsigned H0 = hdr[lane], H1 = hdr[32 + lane], H2 = hdr[64 + lane], H3 = hdr[96 + lane];
        const double2* rec = a.stream + p * a.cap;
        double v[D];
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] = (k == lane) ? 1.0 : 0.0;

        // ---- replay the rotation stream
        auto header = [&](int s) -> unsigned {
            const unsigned h = (s < 32) ? H0 : (s < 64) ? H1 : (s < 96) ? H2 : H3;
            return __shfl_sync(0xffffffffu, h, s & 31);
        };
        int pos = 0;
        unsigned h = header(0);
        double2 cur = make_double2(1.0, 0.0);
        if (h != kTerm) {
            const int l0 = h & 0xff, m0 = (h >> 8) & 0xff;
            if (lane < m0 - l0) cur = rec[lane];
        }
        for (int s = 0; h != kTerm; ++s) {
            const int l = h & 0xff, m = (h >> 8) & 0xff;
            const int cnt = m - l;
            const unsigned hn = header(s + 1);
            double2 nxt = make_double2(1.0, 0.0);
            const int posn = pos + cnt;
            if (hn != kTerm) {
                const int ln = hn & 0xff, mn = (hn >> 8) & 0xff;
                if (lane < mn - ln) nxt = rec[posn + lane];
            }
            double2* rb = rbuf + (s & 1) * 32;
            if (lane < cnt) rb[m - 1 - lane] = cur;         // record j of the sweep acts on plane m-1-j
            __syncwarp();
            bool active = false;
#pragma unroll
            for (int i = D - 2; i >= 0; --i) {
                if (i == m - 1) active = true;
                if (active) {
                    const double2 cs = rb[i];
                    const double x = v[i], y = v[i + 1];
                    v[i + 1] = fma(cs.y, x, cs.x * y);
                    v[i] = fma(cs.x, x, -(cs.y * y));
                }
                if (i == l) active = false;
            }
            cur = nxt; pos = posn; h = hn;
        }

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
            double pr0 = 0.0, pi0 = 0.0, pr1 = 0.0, pi1 = 0.0;
#pragma unroll
            for (int i = 0; i + 1 < D; i += 2) {
                const double2 z0 = zb[i], z1 = zb[i + 1];
                pr0 = fma(v[i], z0.x, pr0); pi0 = fma(v[i], z0.y, pi0);
                pr1 = fma(v[i + 1], z1.x, pr1); pi1 = fma(v[i + 1], z1.y, pi1);
            }
            if (D & 1) {
                const double2 z0 = zb[D - 1];
                pr0 = fma(v[D - 1], z0.x, pr0); pi0 = fma(v[D - 1], z0.y, pi0);
            }
            const double pr = pr0 + pr1, pi = pi0 + pi1;
            U[k * LDU + lane] = wsite * fma(pr, pr, pi * pi);
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
#pragma unroll
        for (int i = 0; i < D; ++i) U[lane * LDU + i] = v[i];
        __syncwarp();
        double ar = 0.0, ai = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const double vt = U[j * LDU + lane];          // V[j][lane]
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
    const int urows = (T > 32) ? T : 32;
    int udoubles = urows * 33;
    const int need_s = 32 + D * (D - 1) / 2;
    if (need_s > udoubles) udoubles = need_s;
    udoubles = (udoubles + 1) & ~1;
    const int per_warp = 128 + 128 + 8 * 32 + udoubles;
    const size_t smem = (size_t)per_warp * kWarpsB * sizeof(double);
    QTET_CUDA(cudaFuncSetAttribute(apply_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
    char* blob = nullptr;
    size_t blob_bytes = 0;
    unsigned short* pair_tab = nullptr;
    int* overflow_count = nullptr;   // [1]
};

bool split_eligible(const ProblemView& pv) {
    return pv.tridiag && pv.d >= 2 && pv.d <= 32 && pv.T <= 64 && pv.f <= 8;
}

void split_workspace_free(SplitWorkspace* ws) {
    if (!ws) return;
    if (ws->blob) cudaFree(ws->blob);
    if (ws->pair_tab) cudaFree(ws->pair_tab);
    if (ws->overflow_count) cudaFree(ws->overflow_count);
    delete ws;
}

// generic kernel restricted to an index list (qtet_generic.cu)
int launch_generic_list(const ProblemView& pv, const double* chis, const int* list, const int* count, long long max_count,
                        int site, double* loss, double* grad, int32_t* tstar, double* trace, cudaStream_t st);

int launch_split(const ProblemView& pv, SplitWorkspace** wsp, const double* chis, int64_t B, int site,
                 double* loss, double* grad, int32_t* tstar, double* trace, cudaStream_t st) {
    if (B <= 0) return QTET_OK;
    if (!split_eligible(pv)) { set_error("split path not eligible for this problem"); return QTET_EUNSUPPORTED; }
    SplitWorkspace* ws = *wsp;
    const int d = pv.d;
    if (!ws) {
        ws = new SplitWorkspace();
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
        QTET_CUDA(cudaMalloc(&ws->overflow_count, sizeof(int)));
        *wsp = ws;
    }
    // per-point workspace: rotation stream + headers + eigenvalues + overflow slot
    const size_t per_point = (size_t)ws->cap * sizeof(double2) + kMaxSweeps * sizeof(unsigned) + (size_t)d * 8 + 4;
    long long chunk = (long long)((size_t)3 << 30) / (long long)per_point;       // <= 3 GiB of streams in flight
    chunk = (chunk / 4096) * 4096;
    if (chunk < 4096) chunk = 4096;
    if (chunk > B) chunk = ((B + 31) / 32) * 32;
    const size_t o_stream = 0;
    const size_t o_hdr = o_stream + (size_t)chunk * ws->cap * sizeof(double2);
    const size_t o_ev = o_hdr + (size_t)chunk * kMaxSweeps * sizeof(unsigned);
    const size_t o_list = o_ev + (((size_t)chunk * d * 8 + 255) / 256) * 256;
    const size_t total = o_list + (size_t)chunk * 4 + 256;
    if (total > ws->blob_bytes) {
        if (ws->blob) cudaFree(ws->blob);
        ws->blob = nullptr; ws->blob_bytes = 0;
        cudaError_t e = cudaMalloc(&ws->blob, total);
        if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(split workspace) failed"); return QTET_ENOMEM; }
        ws->blob_bytes = total;
    }
    ws->chunk = chunk;

    SplitArgs a;
    a.pv = pv; a.site = site; a.acceptor = (site == pv.f - 1) ? 1 : 0; a.cap = ws->cap;
    a.stream = (double2*)(ws->blob + o_stream); a.hdr = (unsigned*)(ws->blob + o_hdr);
    a.ev = (double*)(ws->blob + o_ev); a.overflow_list = (int*)(ws->blob + o_list);
    a.overflow_count = ws->overflow_count; a.pair_tab = ws->pair_tab;

    const size_t smemA = (size_t)kWarpsA * 2 * d * 32 * sizeof(double);
    QTET_CUDA(cudaFuncSetAttribute(ql_scalar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA));
    int per_sm_a = 0;
    QTET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_a, ql_scalar_kernel, kWarpsA * 32, smemA));
    if (per_sm_a < 1) per_sm_a = 1;

    for (long long lo = 0; lo < B; lo += chunk) {
        const long long n = (B - lo < chunk) ? (B - lo) : chunk;
        a.chis = chis + lo * pv.f; a.n = n;
        a.loss = loss ? loss + lo : nullptr;
        a.grad = grad ? grad + lo * pv.f : nullptr;
        a.tstar = tstar ? tstar + lo : nullptr;
        a.trace = trace ? trace + lo * pv.T : nullptr;
        QTET_CUDA(cudaMemsetAsync(ws->overflow_count, 0, sizeof(int), st));
        long long gridA = (long long)ws->sms * per_sm_a;
        const long long needA = ((n + 31) / 32 + kWarpsA - 1) / kWarpsA;
        if (gridA > needA) gridA = needA;
        ql_scalar_kernel<<<(unsigned)gridA, kWarpsA * 32, smemA, st>>>(a);
        count_launch();
        QTET_CUDA(cudaGetLastError());
        int rc;
        switch (ws->D) {
            case 4: rc = launch_apply<4>(a, ws->sms, st); break;
            case 6: rc = launch_apply<6>(a, ws->sms, st); break;
            case 8: rc = launch_apply<8>(a, ws->sms, st); break;
            case 12: rc = launch_apply<12>(a, ws->sms, st); break;
            case 16: rc = launch_apply<16>(a, ws->sms, st); break;
            case 21: rc = launch_apply<21>(a, ws->sms, st); break;
            case 24: rc = launch_apply<24>(a, ws->sms, st); break;
            default: rc = launch_apply<32>(a, ws->sms, st); break;
        }
        if (rc) return rc;
        // overflowed points (if any) are recomputed by the fused generic kernel
        rc = launch_generic_list(pv, a.chis, a.overflow_list, ws->overflow_count, n, site, a.loss, a.grad, a.tstar, a.trace, st);
        if (rc) return rc;
    }
    return QTET_OK;
}

}  // namespace qtet
