// C ABI of libqtet_b200.so (see include/qtet.h).  Host-side problem setup + dispatch.
#include <atomic>
#include <cmath>
#include <cstring>
#include <map>
#include <new>

#include "qtet_common.cuh"

namespace qtet {

static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};

void set_error(const std::string& msg) { g_last_error = msg; }
int cuda_fail(cudaError_t e, const char* what) {
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return QTET_ECUDA;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Descending-lexicographic enumeration = the order Loss.derive produces (HamiltonianLoss.py:62-110).
static void enumerate_states(int N, int f, std::vector<int>& cur, int k, int remaining,
                             std::vector<std::vector<int>>& out) {
    if (k == f - 1) {
        cur[k] = remaining;
        out.push_back(cur);
        return;
    }
    for (int n = remaining; n >= 0; --n) {
        cur[k] = n;
        enumerate_states(N, f, cur, k + 1, remaining - n, out);
    }
}

int ensure_scratch(qtet_problem* p, size_t bytes) {
    if (bytes <= p->scratch_bytes) return QTET_OK;
    if (p->scratch) cudaFree(p->scratch);
    if (p->sort_ws) cudaFree(p->sort_ws);
    p->scratch = nullptr;
    p->scratch_bytes = 0;
    size_t want = bytes + bytes / 4 + 4096;
    cudaError_t e = cudaMalloc(&p->scratch, want);
    if (e != cudaSuccess) { set_error("cudaMalloc(scratch) failed"); cudaGetLastError(); return QTET_ENOMEM; }
    p->scratch_bytes = want;
    return QTET_OK;
}

static int auto_path(const ProblemView& v) {
    if (direct_eligible(v) || big_direct_eligible(v)) return QTET_PATH_DIRECT;
    return (split_eligible(v) || big_eligible(v)) ? QTET_PATH_SPLIT : QTET_PATH_GENERIC;
}

int loss_grad_dispatch(qtet_problem* p, const double* chis, int64_t B, int site, double* loss,
                       double* grad, int32_t* tstar, double* trace, cudaStream_t st, const ChunkHook* hook,
                       const int* n_dev) {
    if (p->path == QTET_PATH_DIRECT && !direct_eligible(p->view)) {
        if (n_dev) { set_error("a device-side batch size is only supported on the d <= 32 direct path"); return QTET_EINVAL; }
        return launch_big_direct(p->view, &p->big_ws, chis, B, site, loss, grad, tstar, trace, st, p->prof.on ? &p->prof : nullptr, hook);
    }
    if (p->path == QTET_PATH_DIRECT)
        return launch_direct(p->view, &p->direct_ws, chis, B, n_dev, site, loss, grad, tstar, trace, st, p->prof.on ? &p->prof : nullptr, hook);
    if (n_dev) { set_error("a device-side batch size is only supported on the direct path"); return QTET_EINVAL; }
    if (p->path == QTET_PATH_SPLIT && !split_eligible(p->view))
        return launch_big(p->view, &p->big_ws, chis, B, site, loss, grad, tstar, trace, st, p->prof.on ? &p->prof : nullptr);
    if (p->path == QTET_PATH_SPLIT)
        return launch_split(p->view, &p->split_ws, chis, B, site, loss, grad, tstar, trace, st, p->prof.on ? &p->prof : nullptr, hook);
    return launch_generic(p->view, chis, B, site, loss, grad, tstar, trace, st, p->prof.on ? &p->prof : nullptr);
}

}  // namespace qtet

using namespace qtet;

extern "C" {

int qtet_problem_create(int max_N, int sites, const double* omegas, double coupling, double max_t,
                        int timesteps, int device, qtet_problem** out) {
    if (!out) { set_error("out is NULL"); return QTET_EINVAL; }
    *out = nullptr;
    if (max_N < 1 || sites < 2 || !omegas || timesteps < 1) {
        set_error("need max_N >= 1, sites >= 2, omegas != NULL, timesteps >= 1");
        return QTET_EINVAL;
    }
    if (sites > 8) { set_error("at most 8 sites are supported"); return QTET_EUNSUPPORTED; }
    std::vector<std::vector<int>> st;
    {
        // guard the enumeration size before doing it
        double dim = 1.0;
        for (int i = 1; i <= sites - 1; ++i) dim = dim * (max_N + i) / i;
        // d <= 115: shared-memory / register kernels; beyond that the fused generic kernel keeps V and M in global memory, its
        // vectors (17 d doubles per CTA) must still fit the 227 KB of shared memory
        if (dim > 1536) { set_error("Fock dimension " + std::to_string((long long)dim) + " is too large (supported: up to 1536)"); return QTET_EUNSUPPORTED; }
        std::vector<int> cur(sites, 0);
        enumerate_states(max_N, sites, cur, 0, max_N, st);
    }
    const int d = (int)st.size();
    QTET_ON_DEVICE(device);

    qtet_problem* p = new (std::nothrow) qtet_problem();
    if (!p) { set_error("out of host memory"); return QTET_ENOMEM; }
    p->N = max_N; p->f = sites; p->d = d; p->T = timesteps; p->device = device;
    p->coupling = coupling; p->max_t = max_t;
    p->omegas.assign(omegas, omegas + sites);
    p->states.resize((size_t)d * sites);
    std::map<std::vector<int>, int> index;
    for (int m = 0; m < d; ++m) {
        index[st[m]] = m;
        for (int k = 0; k < sites; ++k) p->states[(size_t)m * sites + k] = st[m][k];
    }
    // hopping part (HamiltonianLoss.py:136-163); the reference locates the neighbour state through
    // a float hash + searchsorted (:112-121), an exact map lookup gives the same index.
    std::vector<double> dense((size_t)d * d, 0.0);
    int bandwidth = 0;
    for (int m = 0; m < d; ++m) {
        for (int k = 0; k < sites - 1; ++k) {
            const int nk = st[m][k], nk1 = st[m][k + 1];
            if (nk != 0) {
                std::vector<int> t = st[m]; t[k] -= 1; t[k + 1] += 1;
                const int n = index[t];
                dense[(size_t)n * d + m] -= coupling * std::sqrt((double)(nk1 + 1) * nk);
                bandwidth = std::max(bandwidth, std::abs(n - m));
            }
            if (nk1 != 0) {
                std::vector<int> t = st[m]; t[k] += 1; t[k + 1] -= 1;
                const int n = index[t];
                dense[(size_t)n * d + m] -= coupling * std::sqrt((double)nk1 * (nk + 1));
                bandwidth = std::max(bandwidth, std::abs(n - m));
            }
        }
    }
    p->tridiag = bandwidth <= 1;
    bool offd_nonzero = p->tridiag && d >= 2;
    for (int m = 0; m + 1 < d && offd_nonzero; ++m) offd_nonzero = dense[(size_t)(m + 1) * d + m] != 0.0;
    // time grid = numpy.linspace(0, max_t, T): k*step, last sample exactly max_t
    p->tgrid.resize(timesteps);
    if (timesteps == 1) p->tgrid[0] = 0.0;
    else {
        const double step = max_t / (double)(timesteps - 1);
        for (int k = 0; k < timesteps; ++k) p->tgrid[k] = (double)k * step;
        p->tgrid[timesteps - 1] = max_t;
    }
    // pack constants
    std::vector<double> blob;
    // every section starts on a 16-byte boundary (the tiled Householder kernel reads the hopping rows with 16-byte loads)
    auto push = [&](const std::vector<double>& v) { if (blob.size() & 1) blob.push_back(0.0); size_t o = blob.size(); blob.insert(blob.end(), v.begin(), v.end()); return o; };
    std::vector<double> lin(d), quad((size_t)d * sites);
    for (int m = 0; m < d; ++m) {
        double acc = 0.0;
        for (int k = 0; k < sites; ++k) {
            const double n = st[m][k];
            acc += omegas[k] * n;
            quad[(size_t)m * sites + k] = 0.5 * n * n;
        }
        lin[m] = acc;
    }
    std::vector<double> offd;
    if (p->tridiag) {
        offd.resize(d > 1 ? d - 1 : 1, 0.0);
        for (int m = 0; m + 1 < d; ++m) offd[m] = dense[(size_t)(m + 1) * d + m];
    } else {
        offd = dense;
    }
    const size_t o_lin = push(lin), o_quad = push(quad), o_occ = push(p->states), o_off = push(offd),
                 o_t = push(p->tgrid), o_om = push(p->omegas);
    if (blob.size() & 1) blob.push_back(0.0);
    const size_t o_stats = blob.size();
    blob.resize(blob.size() + 4, 0.0);            // 4 x uint64 counters (all-zero bit pattern)
    cudaError_t e = cudaMalloc(&p->dev_blob, blob.size() * sizeof(double));
    if (e != cudaSuccess) { delete p; return cuda_fail(e, "cudaMalloc(constants)"); }
    e = cudaMemcpy(p->dev_blob, blob.data(), blob.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(p->dev_blob); delete p; return cuda_fail(e, "cudaMemcpy(constants)"); }
    ProblemView& v = p->view;
    v.N = max_N; v.f = sites; v.d = d; v.T = timesteps; v.tridiag = p->tridiag ? 1 : 0; v.offd_nonzero = offd_nonzero ? 1 : 0; v.max_t = max_t;
    v.lin = p->dev_blob + o_lin; v.quad = p->dev_blob + o_quad; v.occ = p->dev_blob + o_occ;
    v.offd = p->dev_blob + o_off; v.tgrid = p->dev_blob + o_t; v.omegas = p->dev_blob + o_om;
    v.stats = reinterpret_cast<unsigned long long*>(p->dev_blob + o_stats);
    p->path = auto_path(v);
    *out = p;
    return QTET_OK;
}

void qtet_problem_destroy(qtet_problem* p) {
    if (!p) return;
    DeviceGuard _guard(p->device);
    if (p->dev_blob) cudaFree(p->dev_blob);
    if (p->scratch) cudaFree(p->scratch);
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_run) cudaStreamDestroy(p->s_run);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    if (p->s_poll) cudaStreamDestroy(p->s_poll);
    if (p->h_poll) cudaFreeHost(p->h_poll);
    for (int i = 0; i < qtet_problem::kPollRing; ++i) {
        if (p->ev_cnt[i]) cudaEventDestroy(p->ev_cnt[i]);
        if (p->ev_poll[i]) cudaEventDestroy(p->ev_poll[i]);
    }
    for (int i = 0; i < qtet_problem::kMaxSlices; ++i) {
        if (p->ev_in[i]) cudaEventDestroy(p->ev_in[i]);
        if (p->ev_run[i]) cudaEventDestroy(p->ev_run[i]);
    }
    split_workspace_free(p->split_ws);
    direct_workspace_free(p->direct_ws);
    big_workspace_free(p->big_ws);
    p->prof.clear();
    delete p;
}

int qtet_problem_dim(const qtet_problem* p) { return p ? p->d : QTET_EINVAL; }

int qtet_problem_states(const qtet_problem* p, double* states_host) {
    if (!p || !states_host) { set_error("null argument"); return QTET_EINVAL; }
    std::memcpy(states_host, p->states.data(), p->states.size() * sizeof(double));
    return QTET_OK;
}

int qtet_problem_time_grid(const qtet_problem* p, double* t_host) {
    if (!p || !t_host) { set_error("null argument"); return QTET_EINVAL; }
    std::memcpy(t_host, p->tgrid.data(), p->tgrid.size() * sizeof(double));
    return QTET_OK;
}

int qtet_set_path(qtet_problem* p, int path) {
    if (!p) { set_error("null handle"); return QTET_EINVAL; }
    if (path == QTET_PATH_AUTO) { p->path = auto_path(p->view); return QTET_OK; }
    if (path == QTET_PATH_DIRECT) {
        if (!direct_eligible(p->view) && !big_direct_eligible(p->view)) { set_error("direct path needs non-zero hopping and dim <= 101 (tridiagonal H) / 84 (dense H)"); return QTET_EUNSUPPORTED; }
        p->path = path;
        return QTET_OK;
    }
    if (path == QTET_PATH_GENERIC) { p->path = path; return QTET_OK; }
    if (path == QTET_PATH_SPLIT) {
        if (!split_eligible(p->view) && !big_eligible(p->view)) { set_error("the split path needs Fock dimension <= 115 (larger systems run on the generic path)"); return QTET_EUNSUPPORTED; }
        p->path = path;
        return QTET_OK;
    }
    set_error("unknown path");
    return QTET_EINVAL;
}

int qtet_get_path(const qtet_problem* p) { return p ? p->path : QTET_EINVAL; }

// buffers of the locality-ordered qtet_loss_grad (order, keys, histogram, gathered inputs, ordered outputs)
static size_t sort_ws_bytes(int64_t B, size_t f) {
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    return 2 * al((size_t)B * 4) + al(65536 * 4) + 256 + 2 * al((size_t)B * f * 8) + al((size_t)B * 8) + al((size_t)B * 4);
}
static int ensure_sort_ws(qtet_problem* p, int64_t B) {
    const size_t total = sort_ws_bytes(B, (size_t)p->f);
    if (total <= p->sort_bytes) return QTET_OK;
    if (p->sort_ws) { cudaDeviceSynchronize(); cudaFree(p->sort_ws); }
    p->sort_ws = nullptr; p->sort_bytes = 0;
    if (cudaMalloc(&p->sort_ws, total) != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(locality-sort buffers) failed"); return QTET_ENOMEM; }
    p->sort_bytes = total;
    return QTET_OK;
}

int qtet_problem_reserve(qtet_problem* p, int64_t B) {
    if (!p || B < 0) { set_error("bad argument"); return QTET_EINVAL; }
    QTET_ON_DEVICE(p->device);
    const size_t f = p->f;
    // scratch of the *_host entry point and of the optimiser (the larger of the two layouts)
    const size_t host_call = (size_t)B * (2 * f + 1) * 8 + (size_t)B * 4 + 64;
    const size_t adam = (size_t)B * (f * 52 + 24 + 28) + 65536 * 4 + 32 * 256;        // qtet_adam_run's layout incl. the re-ordering buffers
    int rc = ensure_scratch(p, host_call > adam ? host_call : adam);
    if (rc) return rc;
    if (p->locality_sort && B >= 4096) { rc = ensure_sort_ws(p, B); if (rc) return rc; }
    if (p->path == QTET_PATH_DIRECT && direct_eligible(p->view)) return direct_reserve(p->view, &p->direct_ws, B, B <= (1LL << 22));
    return QTET_OK;
}

static int check_call(qtet_problem* p, const void* chis, int64_t B, int site) {
    if (!p) { set_error("null handle"); return QTET_EINVAL; }
    if (B < 0) { set_error("negative batch"); return QTET_EINVAL; }
    if (B > 0 && !chis) { set_error("chis is NULL"); return QTET_EINVAL; }
    if (site < 0 || site >= p->f) { set_error("target_site out of range"); return QTET_EINVAL; }
    return QTET_OK;
}

int qtet_loss_grad(qtet_problem* p, const double* chis_dev, int64_t B, int target_site,
                   double* loss_dev, double* grad_dev, int32_t* tstar_dev, void* stream) {
    int rc = check_call(p, chis_dev, B, target_site);
    if (rc) return rc;
    if (B == 0) return QTET_OK;
    if (!loss_dev) { set_error("loss is NULL"); return QTET_EINVAL; }
    QTET_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (!p->locality_sort || B < 4096)
        return loss_grad_dispatch(p, chis_dev, B, target_site, loss_dev, grad_dev, tstar_dev, nullptr, st);
    // Locality order: warps of the kernels hold consecutive points of the batch and cost what their most expensive point costs, so
    // unordered (random) batches run ~20 % below grid-ordered ones.  Order by a coarse binning, evaluate, scatter back -- every
    // result is bit-identical to the unordered call (batch invariance).
    const size_t f = p->f;
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t o_order = 0, o_keys = o_order + al(B * 4), o_hist = o_keys + al(B * 4), o_mm = o_hist + al(65536 * 4),
                 o_xs = o_mm + 256, o_loss = o_xs + al(B * f * 8), o_grad = o_loss + al(B * 8), o_ts = o_grad + al(B * f * 8);
    rc = ensure_sort_ws(p, B);          // (sized by qtet_problem_reserve if the option was set before it)
    if (rc) return rc;
    char* w = (char*)p->sort_ws;
    int* order = (int*)(w + o_order);
    double* xs = (double*)(w + o_xs);
    double* loss_s = (double*)(w + o_loss);
    double* grad_s = grad_dev ? (double*)(w + o_grad) : nullptr;
    int* ts_s = tstar_dev ? (int*)(w + o_ts) : nullptr;
    rc = launch_locality_order(chis_dev, B, (int)f, order, (int*)(w + o_keys), (int*)(w + o_hist), (unsigned long long*)(w + o_mm), st);
    if (rc) return rc;
    rc = launch_sort_gather(order, B, (int)f, chis_dev, xs, st);
    if (rc) return rc;
    rc = loss_grad_dispatch(p, xs, B, target_site, loss_s, grad_s, ts_s, nullptr, st);
    if (rc) return rc;
    return launch_sort_scatter(order, B, (int)f, loss_s, grad_s, ts_s, loss_dev, grad_dev, tstar_dev, st);
}

int qtet_set_locality_sort(qtet_problem* p, int on) {
    if (!p) { set_error("null handle"); return QTET_EINVAL; }
    p->locality_sort = on ? 1 : 0;
    return QTET_OK;
}

int qtet_trace(qtet_problem* p, const double* chis_dev, int64_t B, int target_site, double* trace_dev, void* stream) {
    int rc = check_call(p, chis_dev, B, target_site);
    if (rc) return rc;
    if (B == 0) return QTET_OK;
    if (!trace_dev) { set_error("trace is NULL"); return QTET_EINVAL; }
    QTET_ON_DEVICE(p->device);
    return loss_grad_dispatch(p, chis_dev, B, target_site, nullptr, nullptr, nullptr, trace_dev, (cudaStream_t)stream);
}

int qtet_loss_grad_host(qtet_problem* p, const double* chis_host, int64_t B, int target_site,
                        double* loss_host, double* grad_host, int32_t* tstar_host) {
    int rc = check_call(p, chis_host, B, target_site);
    if (rc) return rc;
    if (B == 0) return QTET_OK;
    if (!loss_host) { set_error("loss is NULL"); return QTET_EINVAL; }
    QTET_ON_DEVICE(p->device);
    const size_t f = p->f;
    const size_t nb_chi = (size_t)B * f * 8, nb_loss = (size_t)B * 8, nb_grad = (size_t)B * f * 8, nb_t = (size_t)B * 4;
    rc = ensure_scratch(p, nb_chi + nb_loss + nb_grad + nb_t + 64);
    if (rc) return rc;
    char* base = (char*)p->scratch;
    double* dchi = (double*)base;
    double* dloss = (double*)(base + nb_chi);
    double* dgrad = (double*)(base + nb_chi + nb_loss);
    int32_t* dts = (int32_t*)(base + nb_chi + nb_loss + nb_grad);
    // H2D, kernels and D2H run on three streams.  The register split path processes the batch in chunks (two workspace
    // buffers, qtet_split.cu); the copies hang on its per-chunk hooks, so the inputs of chunk c+1 arrive and the results of
    // chunk c-1 leave while chunk c computes.  The other paths copy the whole batch before / after their single launch
    // sequence (the large-d path wants every batch it can get in one launch: scalar QL latency, qtet_big.cu).
    if (!p->s_in) {
        QTET_CUDA(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
        QTET_CUDA(cudaStreamCreateWithFlags(&p->s_run, cudaStreamNonBlocking));
        QTET_CUDA(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < qtet_problem::kMaxSlices; ++i) {
            QTET_CUDA(cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming));
            QTET_CUDA(cudaEventCreateWithFlags(&p->ev_run[i], cudaEventDisableTiming));
        }
    }
    struct Ctx {
        qtet_problem* p; size_t f;
        const double* chis_host; double* dchi;
        double* loss_host; double* dloss; double* grad_host; double* dgrad; int32_t* tstar_host; int32_t* dts;
        int nin, nout;
    } cx{p, f, chis_host, dchi, loss_host, dloss, grad_host, dgrad, tstar_host, dts, 0, 0};
    auto copy_in = [](void* c_, long long lo, long long n, cudaStream_t consumer) -> int {
        Ctx& c = *(Ctx*)c_;
        cudaEvent_t ev = c.p->ev_in[c.nin++ % qtet_problem::kMaxSlices];
        QTET_CUDA(cudaMemcpyAsync(c.dchi + lo * c.f, c.chis_host + lo * c.f, (size_t)n * c.f * 8, cudaMemcpyHostToDevice, c.p->s_in));
        QTET_CUDA(cudaEventRecord(ev, c.p->s_in));
        QTET_CUDA(cudaStreamWaitEvent(consumer, ev, 0));
        return QTET_OK;
    };
    auto copy_out = [](void* c_, long long lo, long long n, cudaStream_t producer) -> int {
        Ctx& c = *(Ctx*)c_;
        cudaEvent_t ev = c.p->ev_run[c.nout++ % qtet_problem::kMaxSlices];
        QTET_CUDA(cudaEventRecord(ev, producer));
        QTET_CUDA(cudaStreamWaitEvent(c.p->s_out, ev, 0));
        QTET_CUDA(cudaMemcpyAsync(c.loss_host + lo, c.dloss + lo, (size_t)n * 8, cudaMemcpyDeviceToHost, c.p->s_out));
        if (c.grad_host) QTET_CUDA(cudaMemcpyAsync(c.grad_host + lo * c.f, c.dgrad + lo * c.f, (size_t)n * c.f * 8, cudaMemcpyDeviceToHost, c.p->s_out));
        if (c.tstar_host) QTET_CUDA(cudaMemcpyAsync(c.tstar_host + lo, c.dts + lo, (size_t)n * 4, cudaMemcpyDeviceToHost, c.p->s_out));
        return QTET_OK;
    };
    const bool chunked = (p->path == QTET_PATH_DIRECT) || (p->path == QTET_PATH_SPLIT && split_eligible(p->view));
    if (chunked) {
        ChunkHook hook;
        hook.ctx = &cx; hook.before = copy_in; hook.after = copy_out;
        rc = loss_grad_dispatch(p, dchi, B, target_site, dloss, grad_host ? dgrad : nullptr, tstar_host ? dts : nullptr, nullptr, p->s_run, &hook);
        if (rc) return rc;
    } else {
        rc = copy_in(&cx, 0, B, p->s_run);
        if (rc) return rc;
        rc = loss_grad_dispatch(p, dchi, B, target_site, dloss, grad_host ? dgrad : nullptr, tstar_host ? dts : nullptr, nullptr, p->s_run);
        if (rc) return rc;
        rc = copy_out(&cx, 0, B, p->s_run);
        if (rc) return rc;
    }
    QTET_CUDA(cudaStreamSynchronize(p->s_out));
    QTET_CUDA(cudaStreamSynchronize(p->s_run));
    return QTET_OK;
}

int qtet_hamiltonian_host(qtet_problem* p, const double* chis_host, double* H_host) {
    if (!p || !chis_host || !H_host) { set_error("null argument"); return QTET_EINVAL; }
    QTET_ON_DEVICE(p->device);
    const size_t nH = (size_t)p->d * p->d * 8, nchi = (size_t)p->f * 8;
    int rc = ensure_scratch(p, nH + nchi + 64);
    if (rc) return rc;
    double* dchi = (double*)p->scratch;
    double* dH = (double*)((char*)p->scratch + ((nchi + 63) / 64) * 64);
    QTET_CUDA(cudaMemcpy(dchi, chis_host, nchi, cudaMemcpyHostToDevice));
    rc = launch_build_h(p->view, dchi, dH, nullptr);
    if (rc) return rc;
    QTET_CUDA(cudaMemcpy(H_host, dH, nH, cudaMemcpyDeviceToHost));
    return QTET_OK;
}

int qtet_argmin(const double* values_dev, int64_t B, double* min_dev, int64_t* idx_dev, void* stream) {
    if (B <= 0 || !values_dev || !min_dev || !idx_dev) { set_error("bad argument"); return QTET_EINVAL; }
    cudaPointerAttributes attr{};
    QTET_CUDA(cudaPointerGetAttributes(&attr, values_dev));
    if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged) { set_error("values_dev is not device memory"); return QTET_EINVAL; }
    QTET_ON_DEVICE(attr.device);                 // launch where the data lives, whatever the caller's current device is
    return launch_argmin(values_dev, B, min_dev, idx_dev, (cudaStream_t)stream);
}

int qtet_profile_enable(qtet_problem* p, int on) {
    if (!p) { set_error("null handle"); return QTET_EINVAL; }
    QTET_ON_DEVICE(p->device);
    QTET_CUDA(cudaDeviceSynchronize());
    p->prof.clear();
    p->prof.on = on != 0;
    return QTET_OK;
}

int qtet_profile_read(qtet_problem* p, double* ms_sum, int64_t* launches, int64_t* points) {
    if (!p || !ms_sum || !launches || !points) { set_error("null argument"); return QTET_EINVAL; }
    QTET_ON_DEVICE(p->device);
    QTET_CUDA(cudaDeviceSynchronize());
    for (int k = 0; k < Profiler::kKinds; ++k) {
        ms_sum[k] = 0.0; launches[k] = 0; points[k] = 0;
        for (size_t i = 0; i < p->prof.beg[k].size(); ++i) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, p->prof.beg[k][i], p->prof.end[k][i]) == cudaSuccess) {
                ms_sum[k] += ms; launches[k] += 1; points[k] += p->prof.points[k][i];
            } else cudaGetLastError();
        }
    }
    return QTET_OK;
}

int qtet_stats(qtet_problem* p, int64_t* out4, int reset) {
    if (!p || !out4) { set_error("null argument"); return QTET_EINVAL; }
    QTET_ON_DEVICE(p->device);
    QTET_CUDA(cudaDeviceSynchronize());
    unsigned long long h[4] = {0, 0, 0, 0};
    QTET_CUDA(cudaMemcpy(h, p->view.stats, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; ++i) out4[i] = (int64_t)h[i];
    if (reset) QTET_CUDA(cudaMemset(p->view.stats, 0, sizeof(h)));
    return QTET_OK;
}

int64_t qtet_launch_count(void) { return g_launches.load(); }
const char* qtet_last_error(void) { return g_last_error.c_str(); }
const char* qtet_version(void) { return "qtet-b200 0.1 (sm_100a)"; }

}  // extern "C"
