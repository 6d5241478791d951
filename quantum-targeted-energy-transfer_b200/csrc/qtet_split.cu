// Fast path for tridiagonal Hamiltonians with Fock dimension d <= 32 (every dimer up to N = 31;
// BASELINE's headline is the dimer N = 20, d = 21).
//
// The eigendecomposition is split by the kind of parallelism it offers:
//
//   ql_scalar_kernel (K_A)  one THREAD per parameter point, a warp = a BATCH of 32 points.
//        Implicit-shift QL on the tridiagonal (diag, off) only: O(d^2) scalar work with rsqrt/div
//        chains that would waste a whole warp if done per warp.  The warp advances in lock-step
//        ROUNDS (one QL sweep per still-active lane), so that the plane rotations (c, s) it emits
//        form a stream  rec[batch][iteration][lane]  written with fully coalesced 512 B rows, and
//        one header word per (round, lane) says which planes [l, m) the lane swept and where the
//        round starts in the stream.
//   apply_kernel (K_B)  G lanes per parameter point (32/G points of one batch per warp), every lane
//        owning R = ceil(D/G) rows of V in REGISTERS.  Replays the rounds on V = I -- four FP64
//        instructions per row and plane, R independent dependency chains per lane, (c, s) fetched
//        once per plane and point from a cp.async-fed shared-memory ring -- then the time evolution
//        psi = V (c . exp(-i e t)) with phases advanced by a complex rotation recurrence, <n(t)>,
//        the loss, and the analytic gradient through the symmetrised Daleckii-Krein kernel.
//
// Batches whose stream overflows its slot (pathological convergence) are appended to a list and
// recomputed by the fused generic kernel; no host synchronisation is involved.  K_A runs one chunk
// ahead of K_B on an auxiliary stream.
#include <cstdlib>

#include "qtet_split.cuh"

namespace qtet {

using namespace splitk;

namespace {

// Pass 1 only supplies shifts: an eigenvalue taken when its off-diagonal is below 2^-33 relative is still accurate to
// ~e^2/gap (second order), and the looser test saves a quarter of pass 1's rotations (tools/ql_count.py).
constexpr double kEpsPass1 = kEps * 1048576.0;
constexpr int kWarpsA = 4;                 // warps per CTA, scalar kernel

// ------------------------------------------------------------------------------------------------
// K_A: thread-per-point scalar QL.  diag/off live in shared memory as [2k + {0,1}][lane] (bank = lane:
// lanes may sit at different k without conflicts).
//
//   pass 1  plain implicit-shift QL (Wilkinson shifts), eigenvalues only, nothing recorded;
//   pass 2  the tridiagonal is rebuilt and swept again in warp-synchronous ROUNDS, this time recorded.
//           The first sweep at every index l uses the eigenvalue pass 1 left at that index as a PERFECT
//           shift (one sweep deflates it: ~d sweeps and ~d^2/2 rotations instead of ~1.7 d and ~0.9 d^2);
//           if rounding spoils a deflation the loop simply continues with Wilkinson shifts, so the result
//           never depends on pass 1 being right.
//
// A round covers the planes mtop-1 .. lbot (warp-uniform: the union of the lanes' windows); stream row
// it0 + (mtop-1-i) holds, for every lane, the rotation of plane i or the identity if the lane does not
// touch that plane in this round.  Rows are written by all 32 lanes: fully coalesced 512 B stores.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsA * 32) ql_scalar_kernel(SplitArgs a) {
    extern __shared__ double sm[];
    const ProblemView& pv = a.pv;
    const int d = pv.d, f = pv.f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    QlLane t;
    t.arr = sm + (size_t)warp * 2 * d * 32 + lane;
    t.d = d;
    const long long nbatch = (a.n + 31) / 32;
    const unsigned below_top = lowmask(d - 1);
    for (long long batch = (long long)blockIdx.x * kWarpsA + warp; batch < nbatch; batch += (long long)gridDim.x * kWarpsA) {
        const long long p = batch * 32 + lane;
        const bool valid = p < a.n;
        const long long pc = valid ? p : a.n - 1;
        double hchi[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) hchi[k] = (k < f) ? 0.5 * a.chis[pc * f + k] : 0.0;
        double lam[32];                                    // eigenvalue pass 1 left at each index (local memory)
        double mean = 0.0;
        for (int pass = 0; pass < 2; ++pass) {
            // ---- H build (HamiltonianLoss.py:131-134) + global-phase shift by the mean of the diagonal
            double sum = 0.0;
            for (int m = 0; m < d; ++m) {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (k < f) {
                        const double n = pv.occ[m * f + k];
                        acc = acc + (pv.omegas[k] * n + hchi[k] * (n * n));
                    }
                }
                t.dg(m) = acc;
                t.of(m) = (m < d - 1) ? pv.offd[m] : 0.0;
                sum += acc;
            }
            if (pass == 0) mean = sum / (double)d;
            for (int m = 0; m < d; ++m) t.dg(m) -= mean;
            if (pass == 1) break;
            // ---- pass 1: eigenvalues only
            unsigned small = t.scan_small(kEpsPass1);
            int l = 0, iter = 0;
            while (true) {
                const unsigned alive = (~small) & below_top & ~lowmask(l);
                if (alive == 0u) break;
                const int lnew = __ffs(alive) - 1;
                if (lnew != l) iter = 0;
                l = lnew;
                const int m = l + __ffs(small >> l) - 1;
                if (++iter > 60) break;                    // give up quietly: pass 2 does not rely on pass 1
                double g = t.wilkinson_g(l, m);
                double s = 1.0, c = 1.0, pp = 0.0, dnext = 0.0;
                bool broke = false;
                // dg(i+1), dg(i), of(i) of the coming plane travel in registers: each diagonal entry is loaded once (it
                // is read by two consecutive planes) and one plane ahead of its use, off the dependency chain
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
                    small = (r <= kEpsPass1 * (fabs(dnew) + fabs(dnext))) ? (small | fbit) : (small & ~fbit);
                    fbit >>= 1;
                    dnext = dnew; d_hi = d_i; d_i = d_n; e_i = e_n;
                }
                if (broke) { small = t.scan_small(kEpsPass1); continue; }
                const double dlnew = d_hi - pp;                // d_hi = dg(l) after the last plane
                t.dg(l) = dlnew; t.of(l) = g; t.of(m) = 0.0;
                small |= 1u << m;
                const unsigned bit = 1u << l;
                if (fabs(g) <= kEpsPass1 * (fabs(dlnew) + fabs(dnext))) small |= bit; else small &= ~bit;
            }
            for (int m = 0; m < d; ++m) lam[m] = t.dg(m);
        }

        // ---- pass 2: recorded rounds
        double2* rec = a.stream + (size_t)batch * a.capit * 32 + lane;
        unsigned* hdr = a.hdr + (size_t)batch * kMaxRounds * 32 + lane;
        unsigned small = t.scan_small();
        int it = 0, round = 0, l = 0, iter = 0, tried = -1;
        bool overflow = false;
        while (true) {
            const unsigned alive = (~small) & below_top & ~lowmask(l);
            const bool done = (alive == 0u);
            const int lnew = done ? d : (__ffs(alive) - 1);
            if (lnew != l) iter = 0;
            l = lnew;
            if (__all_sync(0xffffffffu, done)) break;
            const int m = done ? 0 : (l + __ffs(small >> l) - 1);       // first negligible off-diagonal above l
            ++iter;
            const int mtop = __reduce_max_sync(0xffffffffu, m);
            const int lbot = __reduce_min_sync(0xffffffffu, done ? 63 : l);
            const int nrows = mtop - lbot;
            if (__any_sync(0xffffffffu, !done && iter > 60) || round >= kMaxRounds - 1 || it + nrows + kTail > a.capit) {
                overflow = true;
                break;
            }
            hdr[round * 32] = (unsigned)(done ? 63 : l)   /* 63 = idle */ | ((unsigned)mtop << 6) | ((unsigned)lbot << 12) | ((unsigned)it << 18);
            double g = 0.0, s = 1.0, c = 1.0, pp = 0.0, dnext = 0.0;
            if (!done) {
                if (tried != l) { tried = l; g = t.dg(m) - lam[l]; }    // perfect shift, first sweep at this index
                else g = t.wilkinson_g(l, m);
            }
            bool broke = false;
            double2* row = rec + (size_t)it * 32;
            // as in pass 1: the coming plane's dg(i+1), dg(i), of(i) are carried in registers and loaded a plane ahead
            double* q = t.arr + 64 * (done ? 0 : (m - 1));
            double d_hi = 0.0, d_i = 0.0, e_i = 0.0;
            if (!done) { d_hi = q[64]; d_i = q[0]; e_i = q[32]; }
            unsigned fbit = done ? 0u : (1u << m);
            for (int i = mtop - 1; i >= lbot; --i, row += 32) {
                double2 out = make_double2(1.0, 0.0);
                if (!done && !broke && i < m && i >= l) {
                    const double* qn = (i > l) ? q - 64 : q;
                    const double d_n = qn[0], e_n = qn[32];
                    const double fo = s * e_i, bo = c * e_i;
                    const double r2 = fma(fo, fo, g * g);
                    if (r2 < kTinyR2) {                    // underflow recovery: the rest of the sweep is identity
                        q[96] = 0.0; q[64] = d_hi - pp; t.of(m) = 0.0;
                        broke = true;
                    } else {
                        const double rinv = rsqrt_fast(r2);
                        const double r = r2 * rinv;
                        s = fo * rinv; c = g * rinv;
                        g = d_hi - pp;
                        const double rr = fma(d_i - g, s, 2.0 * c * bo);
                        pp = s * rr;
                        const double dnew = g + pp;
                        q[64] = dnew;                      // dg(i+1)
                        g = fma(c, rr, -bo);
                        q[96] = r;                         // of(i+1) is final (of(m) is reset below): refresh its flag
                        small = (r <= kEps * (fabs(dnew) + fabs(dnext))) ? (small | fbit) : (small & ~fbit);
                        fbit >>= 1;
                        dnext = dnew; d_hi = d_i; d_i = d_n; e_i = e_n;
                        out = make_double2(c, s);
                        q -= 64;
                    }
                }
                *row = out;                                // all 32 lanes: one coalesced 512 B row
            }
#pragma unroll
            for (int k = 0; k < kTail; ++k, row += 32) *row = make_double2(1.0, 0.0);   // K_B reads whole groups of planes
            it += nrows + kTail;
            if (!done) {
                if (broke) small = t.scan_small();         // rare path: rebuild the flags from scratch
                else {
                    const double dlnew = d_hi - pp;            // d_hi = dg(l) after the last plane
                    t.dg(l) = dlnew; t.of(l) = g; t.of(m) = 0.0;
                    small |= 1u << m;
                    const unsigned bit = 1u << l;
                    if (fabs(g) <= kEps * (fabs(dlnew) + fabs(dnext))) small |= bit; else small &= ~bit;
                }
            }
            ++round;
        }
        if (lane == 0) a.nrounds[batch] = overflow ? 0 : round;
        if (valid) {
            if (overflow) {
                const int slot = atomicAdd(a.overflow_count, 1);
                a.overflow_list[slot] = (int)p;
                atomicAdd(pv.stats + 1, 1ULL);
            }
            for (int m = 0; m < d; ++m) a.ev[p * d + m] = t.dg(m);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// K_B helpers
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// K_B: G lanes per point, R rows per lane in registers.
// ------------------------------------------------------------------------------------------------
template <int D, int G, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) apply_kernel(SplitArgs a) {
    using L = BLayout<D, G>;
    constexpr int R = L::R, RZ = L::RZ, PPW = L::PPW, NPL = L::NPL, NP = L::NP;
    extern __shared__ double sm[];
    const ProblemView& pv = a.pv;
    const int d = pv.d, f = pv.f, T = pv.T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / G, q = lane % G;                       // point slot in the warp, lane within the point
    double* base = sm + (size_t)warp * L::PER_WARP;
    // rotation phase
    double2* ring = reinterpret_cast<double2*>(base);                        // [kRing][SLOT_ROWS][PPW]
    unsigned* hb = reinterpret_cast<unsigned*>(base + L::RING);              // [kMaxRounds][PPW]
    // gradient phase (aliases the rotation phase)
    double* lv = base;                                                       // [8][PPW][D]
    double2* part = reinterpret_cast<double2*>(base + L::LV);                // [PPW][G][DH]
    double* sbuf = base + L::LV;                                             // [PPW][D + NP]
    // always live
    double2* zbuf = reinterpret_cast<double2*>(base + L::PHASED);            // [2][PPW][D]
    double* cbuf = base + L::PHASED + L::ZB;                                 // [PPW][D]

    const double dt = (T > 1) ? pv.tgrid[1] : 0.0;
    double wsite[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int row = q * R + k;
        wsite[k] = (row < d) ? pv.occ[row * f + a.site] : 0.0;
    }
    // roles for moving stream rows into the ring: (row cj0 + ROWS_PER_COPY * t, point cg)
    const int cg = lane % PPW, cj0 = lane / PPW;

    const long long nbatch = (a.n + 31) / 32;
    const long long ntask = nbatch * G;                          // G warps' worth of work per batch
    for (long long task = (long long)blockIdx.x * WARPS + warp; task < ntask; task += (long long)gridDim.x * WARPS) {
        long long _t0 = 0;
        PHASE_START();
        const long long batch = task / G;
        const int sub = (int)(task - batch * G);
        const int blane = sub * PPW + g;                         // my point's lane in the K_A batch
        const long long p = batch * 32 + blane;
        const bool valid = p < a.n;
        const int nr = a.nrounds[batch];
        const double2* rec = a.stream + (size_t)batch * a.capit * 32 + sub * PPW;
        const unsigned* hdr = a.hdr + (size_t)batch * kMaxRounds * 32 + sub * PPW;

        double v[R][D];
#pragma unroll
        for (int k = 0; k < R; ++k)
#pragma unroll
            for (int i = 0; i < D; ++i) v[k][i] = (i == q * R + k) ? 1.0 : 0.0;

        // ---- identity guard rows of every ring slot (planes above a round's top), round headers of my points
        __syncwarp();
#pragma unroll
        for (int sl = 0; sl < kRing; ++sl)
            for (int idx = lane; idx < NPL * PPW; idx += 32) ring[sl * L::SLOT_ROWS * PPW + idx] = make_double2(1.0, 0.0);
        for (int idx = lane; idx < nr * PPW; idx += 32) hb[idx] = hdr[(idx / PPW) * 32 + (idx % PPW)];
        __syncwarp();

        // ---- replay the rounds.  A round's rows (one per plane, top-down from mtop-1) travel global -> shared
        // with cp.async kRing-1 rounds ahead of their use; the cells of my PPW points are 16 B x PPW contiguous
        // in every row and already hold the identity where a point does not touch a plane.
        auto issue = [&](int r) {
            if (r < nr) {
                const unsigned hh = hb[r * PPW];                       // uniform fields of the round
                const int mtop_i = (hh >> 6) & 63, lbot_i = (hh >> 12) & 63, it0 = hh >> 18;
                const int nrows = mtop_i - lbot_i + kTail;
                double2* slot = ring + ((r % kRing) * L::SLOT_ROWS + NPL) * PPW;
#pragma unroll
                for (int t = 0; t < L::NCOPY; ++t) {
                    const int j = cj0 + t * L::ROWS_PER_COPY;
                    if (j < nrows) cp_async16(slot + j * PPW + cg, rec + (size_t)(it0 + j) * 32 + cg);
                }
            }
            cp_async_commit();
        };
#pragma unroll
        for (int r = 0; r < kRing - 1; ++r) issue(r);
        for (int r = 0; r < nr; ++r) {
            __syncwarp();                                   // everyone is done with the slot about to be refilled
            issue(r + kRing - 1);
            const unsigned h = hb[r * PPW + g];
            const int mtop = (h >> 6) & 63;
            const int lcur = __reduce_min_sync(0xffffffffu, (int)(h & 63));   // idle points carry l = 63
            cp_async_wait<kRing - 1>();                     // this lane's copies of round r have landed
            __syncwarp();
            if (lcur >= mtop) continue;                     // none of my points sweeps in this round
            // plane i sits in data row mtop-1-i of the slot; rows above hold the identity (guard), and so do the
            // kTail rows below the round's lowest plane and every cell a point does not touch
            const double2* rb = ring + ((r % kRing) * L::SLOT_ROWS + NPL + mtop - 1) * PPW + g;
            // Planes D-2 .. lcur, top-down, in groups of four with ONE warp-uniform exit test per group.
            // Inside a group the code is branch-free: (c, s) of four planes are fetched ahead of the R
            // independent dependency chains, and only one fma per row and plane sits on a chain:
            //   v[i] <- c x - s y  with  c x  computed off-chain.
#pragma unroll
            for (int i0 = D - 2; i0 >= 0; i0 -= 4) {
                if (i0 < lcur) break;
                if (i0 - 3 >= mtop) continue;               // whole group above the swept range (padding / split)
                double2 cs[4];
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) cs[k4] = rb[-(((i0 - k4) >= 0) ? (i0 - k4) : 0) * PPW];
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const int i = i0 - k4;
                    if (i >= 0) {
#pragma unroll
                        for (int k = 0; k < R; ++k) {
                            const double x = v[k][i], y = v[k][i + 1];
                            const double t1 = cs[k4].x * x, t2 = cs[k4].y * x;
                            v[k][i] = fma(-cs[k4].y, y, t1);
                            v[k][i + 1] = fma(cs[k4].x, y, t2);
                        }
                    }
                }
            }
        }
        cp_async_wait<0>();
        PHASE_MARK(0);
#ifdef QTET_KB_ROT_ONLY      // profiling hook (-DQTET_KB_ROT_ONLY=<CTAs per SM>): the rotation replay alone; results are garbage
        {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < R; ++k)
#pragma unroll
                for (int i = 0; i < D; ++i) acc += v[k][i];
            if (valid && a.loss) a.loss[p] = acc;
            __syncwarp();
            continue;
        }
#endif

        evolve_and_grad<D, G>(a, v, wsite, lv, part, sbuf, zbuf, cbuf, p, valid, g, q, lane, dt, _t0);
    }
}

template <int D, int G, int WARPS, int MINB>
int launch_apply(const SplitArgs& a, int sms, cudaStream_t st) {
    using L = BLayout<D, G>;
    const size_t smem = (size_t)L::PER_WARP * WARPS * sizeof(double);
    auto kern = apply_kernel<D, G, WARPS, MINB>;
    QTET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QTET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int per_sm = 0;
    QTET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    if (const char* env = std::getenv("QTET_KB_CTAS_PER_SM")) { const int v = std::atoi(env); if (v > 0 && v < per_sm) per_sm = v; }
    long long grid = (long long)sms * per_sm;
    const long long ntask = ((a.n + 31) / 32) * G;
    const long long need = (ntask + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(a);
    count_launch();
    QTET_CUDA(cudaGetLastError());
    return QTET_OK;
}

int padded_dim(int d) {
    const int sizes[] = {4, 8, 16, 21, 24, 32};
    for (int s : sizes) if (d <= s) return s;
    return -1;
}

}  // namespace

struct SplitWorkspace {
    int device = 0, sms = 0;
    int capit = 0, D = 0;
    long long chunk = 0;
    char* blob = nullptr;            // two chunk buffers (double-buffered between K_A and K_B)
    size_t blob_bytes = 0;
    int* overflow_count = nullptr;   // [2]
    cudaStream_t aux = nullptr;      // K_A runs here, one chunk ahead of K_B on the caller's stream
    cudaEvent_t ev_fork = nullptr, ev_a[2] = {nullptr, nullptr}, ev_b[2] = {nullptr, nullptr};
};

#ifdef QTET_PHASE_TIMING
extern "C" void qtet_debug_phase_cycles(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)); }
}
#endif

bool split_eligible(const ProblemView& pv) {
    return pv.tridiag && pv.d >= 2 && pv.d <= 32 && pv.f <= 8;
}

void split_workspace_free(SplitWorkspace* ws) {
    if (!ws) return;
    if (ws->blob) cudaFree(ws->blob);
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

int launch_split(const ProblemView& pv, SplitWorkspace** wsp, const double* chis, int64_t B, int site,
                 double* loss, double* grad, int32_t* tstar, double* trace, cudaStream_t st, Profiler* prof,
                 const ChunkHook* hook) {
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
        ws->capit = 2 * d * d + 64;                       // stream rows per batch (warp iterations)
        if (const char* env = std::getenv("QTET_SPLIT_CAPIT")) {   // test hook: force the overflow -> generic fallback
            const int v = std::atoi(env);
            if (v > 0 && v < ws->capit) ws->capit = v;
        }
        QTET_CUDA(cudaMalloc(&ws->overflow_count, 2 * sizeof(int)));
        QTET_CUDA(cudaStreamCreateWithFlags(&ws->aux, cudaStreamNonBlocking));
        QTET_CUDA(cudaEventCreateWithFlags(&ws->ev_fork, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) {
            QTET_CUDA(cudaEventCreateWithFlags(&ws->ev_a[i], cudaEventDisableTiming));
            QTET_CUDA(cudaEventCreateWithFlags(&ws->ev_b[i], cudaEventDisableTiming));
        }
    }
    // per-batch workspace: rotation stream + round headers + round count; per point: eigenvalues + overflow slot
    const size_t per_batch = (size_t)ws->capit * 32 * sizeof(double2) + (size_t)kMaxRounds * 32 * sizeof(unsigned) + 4 +
                             32 * ((size_t)d * 8 + 4);
    // chunking: two buffers in flight so that K_A(c+1) overlaps K_B(c); chunks large enough to fill the machine
    long long chunk_b = (long long)((size_t)2 << 30) / (long long)per_batch;    // <= 2 GiB per buffer
    long long want_b = ((B + 31) / 32 + 3) / 4;
    if (want_b < 2048) want_b = 2048;
    if (chunk_b > want_b) chunk_b = want_b;
    const long long all_b = (B + 31) / 32;
    if (chunk_b > all_b) chunk_b = all_b;
    const long long chunk = chunk_b * 32;
    const size_t o_hdr = (size_t)chunk_b * ws->capit * 32 * sizeof(double2);
    const size_t o_nr = o_hdr + (size_t)chunk_b * kMaxRounds * 32 * sizeof(unsigned);
    const size_t o_ev = o_nr + (((size_t)chunk_b * 4 + 255) / 256) * 256;
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
    a.pv = pv; a.site = site; a.acceptor = (site == pv.f - 1) ? 1 : 0; a.capit = ws->capit;

    const size_t smemA = (size_t)kWarpsA * 2 * d * 32 * sizeof(double);
    QTET_CUDA(cudaFuncSetAttribute(ql_scalar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA));
    QTET_CUDA(cudaFuncSetAttribute(ql_scalar_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int per_sm_a = 0;
    QTET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_a, ql_scalar_kernel, kWarpsA * 32, smemA));
    if (per_sm_a < 1) per_sm_a = 1;
    if (const char* env = std::getenv("QTET_KA_CTAS_PER_SM")) { const int v = std::atoi(env); if (v > 0 && v < per_sm_a) per_sm_a = v; }

    // fork: the auxiliary stream starts after everything already queued on the caller's stream
    QTET_CUDA(cudaEventRecord(ws->ev_fork, st));
    QTET_CUDA(cudaStreamWaitEvent(ws->aux, ws->ev_fork, 0));
    int c = 0;
    for (long long lo = 0; lo < B; lo += chunk, ++c) {
        const int buf = c & 1;
        char* base = ws->blob + (size_t)buf * buf_bytes;
        const long long n = (B - lo < chunk) ? (B - lo) : chunk;
        a.stream = (double2*)base; a.hdr = (unsigned*)(base + o_hdr); a.nrounds = (int*)(base + o_nr);
        a.ev = (double*)(base + o_ev); a.overflow_list = (int*)(base + o_list);
        a.overflow_count = ws->overflow_count + buf;
        a.chis = chis + lo * pv.f; a.n = n;
        a.loss = loss ? loss + lo : nullptr;
        a.grad = grad ? grad + lo * pv.f : nullptr;
        a.tstar = tstar ? tstar + lo : nullptr;
        a.trace = trace ? trace + lo * pv.T : nullptr;
        // K_A(c) on aux, after K_B(c-2) released this buffer
        if (c >= 2) QTET_CUDA(cudaStreamWaitEvent(ws->aux, ws->ev_b[buf], 0));
        if (hook && hook->before) { const int hrc = hook->before(hook->ctx, lo, n, ws->aux); if (hrc) return hrc; }
        QTET_CUDA(cudaMemsetAsync(a.overflow_count, 0, sizeof(int), ws->aux));
        long long gridA = (long long)ws->sms * per_sm_a;
        const long long needA = ((n + 31) / 32 + kWarpsA - 1) / kWarpsA;
        if (gridA > needA) gridA = needA;
        if (prof) prof->begin(0, n, ws->aux);
        ql_scalar_kernel<<<(unsigned)gridA, kWarpsA * 32, smemA, ws->aux>>>(a);
        count_launch();
        if (prof) prof->finish(0, ws->aux);
        QTET_CUDA(cudaGetLastError());
        QTET_CUDA(cudaEventRecord(ws->ev_a[buf], ws->aux));
        // K_B(c) on the caller's stream
        QTET_CUDA(cudaStreamWaitEvent(st, ws->ev_a[buf], 0));
        int rc;
        if (prof) prof->begin(1, n, st);
        switch (ws->D) {
            case 4: rc = launch_apply<4, 4, 4, 4>(a, ws->sms, st); break;
            case 8: rc = launch_apply<8, 8, 4, 4>(a, ws->sms, st); break;
            case 16: rc = launch_apply<16, 8, 2, 6>(a, ws->sms, st); break;
#ifdef QTET_KB_ROT_ONLY
            case 21: rc = launch_apply<21, 8, 2, QTET_KB_ROT_ONLY>(a, ws->sms, st); break;
#else
            case 21: rc = launch_apply<21, 8, 2, 4>(a, ws->sms, st); break;
#endif
            case 24: rc = launch_apply<24, 8, 2, 4>(a, ws->sms, st); break;
            default: rc = launch_apply<32, 16, 2, 4>(a, ws->sms, st); break;
        }
        if (prof) prof->finish(1, st);
        if (rc) return rc;
        // overflowed points (if any) are recomputed by the fused generic kernel
        rc = launch_generic_list(pv, a.chis, a.overflow_list, a.overflow_count, n, site, a.loss, a.grad, a.tstar, a.trace, st);
        if (rc) return rc;
        QTET_CUDA(cudaEventRecord(ws->ev_b[buf], st));
        if (hook && hook->after) { const int hrc = hook->after(hook->ctx, lo, n, st); if (hrc) return hrc; }
    }
    return QTET_OK;
}

}  // namespace qtet
