"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C ABI
(tet._native -> libqtet_b200.so), against the CPU oracle on the same seeded inputs, against the golden
vectors produced by the unmodified reference, and -- at BASELINE sizes -- through size-independent
properties (boson-number conservation, initial condition, bounds).

Tolerances (FP64 path vs the FP64/complex128 oracle; the reference itself evolves in complex64 and is
only good to ~1e-5, see tests/test_oracle.py):
    loss   : |dL|  <= LOSS_ATOL * max(1, N)      absolute     (phase errors scale with eps*||H||*max_t)
    grad   : |dg|  <= GRAD_RTOL * max(1, max|g|) per point
    t_star : identical, except where the two best samples are closer than the loss tolerance
"""
import os

import numpy as np
import pytest

from conftest import make_const, record_ties

pytestmark = pytest.mark.gpu

LOSS_ATOL = 1e-10
GRAD_RTOL = 1e-8


@pytest.fixture(scope="module")
def native():
    from tet import _native
    return _native


@pytest.fixture(scope="module")
def O():
    from oracle import qtet_oracle
    return qtet_oracle


def _compare(native, O, const, X, site, path=None, loss_atol=None, grad_rtol=None):
    prob = native.Problem.from_const(const)
    if path is not None:
        prob.set_path(path)
    loss, grad, ts = prob.loss_grad(X, site)
    loss, grad, ts = loss.cpu().numpy(), grad.cpu().numpy(), ts.cpu().numpy()
    Lr, gr, kr = O.loss_grad_batch(const, X, site)
    N = const["max_N"]
    la = (loss_atol or LOSS_ATOL) * max(1, N)
    assert np.isfinite(loss).all() and np.isfinite(grad).all()
    assert np.abs(loss - Lr).max() <= la, np.abs(loss - Lr).max()
    same_t = ts == kr
    if not same_t.all():
        # a different arg-max sample is acceptable only on a numerical tie of the two samples
        tr = O.trace(const, X[~same_t], site)
        idx = np.arange(tr.shape[0])
        assert np.abs(tr[idx, ts[~same_t]] - tr[idx, kr[~same_t]]).max() <= la
    record_ties(f"N={N} f={const['sites']} site={site}", int((~same_t).sum()), len(X))
    gscale = np.maximum(1.0, np.abs(gr).max(axis=1, keepdims=True))
    err = (np.abs(grad - gr) / gscale)[same_t]
    emax = err.max() if err.size else 0.0          # every point a numerical tie (e.g. no hopping): nothing to compare
    assert emax <= (grad_rtol or GRAD_RTOL), emax
    prob.close()
    return np.abs(loss - Lr).max(), emax


@pytest.mark.parametrize("N", [1, 2, 3, 4, 7, 20, 31])
@pytest.mark.parametrize("site", [1, 0])
def test_dimer_loss_grad_vs_oracle(native, O, N, site):
    rng = np.random.default_rng(100 + N)
    X = rng.uniform(-10, 10, (257, 2))
    _compare(native, O, make_const(N), X, site)


@pytest.mark.parametrize("N", [3, 20, 31])
def test_dimer_generic_path_vs_oracle(native, O, N):
    rng = np.random.default_rng(7 + N)
    X = rng.uniform(-10, 10, (130, 2))
    _compare(native, O, make_const(N), X, 1, path=native.QTET_PATH_GENERIC)


def test_dimer_large_fock_space(native, O):
    # BASELINE config 3 shape (dimer N=100, d=101); phases reach |e t| ~ 1e6 so the tolerance is the
    # eps*||H||*max_t scale of SURVEY.md §7 ("hard parts")
    rng = np.random.default_rng(0)
    X = rng.uniform(-10, 10, (48, 2))
    _compare(native, O, make_const(100), X, 1, loss_atol=2e-9, grad_rtol=2e-6)


@pytest.mark.parametrize("N,omegas", [(2, (-3, 3, 3)), (4, (-3, 3, 3)), (10, (-3, 3, 3)), (5, (-3, 1, 3)), (3, (-3, 0, 1, 3))])
def test_chain_loss_grad_vs_oracle(native, O, N, omegas):
    rng = np.random.default_rng(N)
    f = len(omegas)
    X = rng.uniform(-10, 10, (96, f))
    for site in (f - 1, 0, 1):
        _compare(native, O, make_const(N, omegas), X, site, loss_atol=2e-10, grad_rtol=5e-8)


@pytest.mark.parametrize("N,omegas", [
    (33, (-3, 3)),          # d = 34  -> register kernels padded to 48
    (50, (-3, 3)),          # d = 51  -> 66
    (70, (-3, 3)),          # d = 71  -> 84
    (90, (-3, 3)),          # d = 91  -> 101
    (3, (-3, 3, 3)),        # d = 10  -> 16, dense (Householder in registers)
    (6, (-3, 3, 3)),        # d = 28  -> 32, dense
    (8, (-3, 1, 3)),        # d = 45  -> 48, dense
    (5, (-3, 0, 1, 3)),     # d = 56  -> 66, dense 4-site chain
    (6, (-3, 0, 1, 3)),     # d = 84  -> 84, dense
    (12, (-3, 3, 3)),       # d = 91  -> 101, dense
])
def test_register_resident_large_d_kernels(native, O, N, omegas):
    """Every padded size of the register-resident large-d kernels (qtet_bigreg.cu), tridiagonal and dense."""
    rng = np.random.default_rng(1000 + N)
    f = len(omegas)
    X = rng.uniform(-10, 10, (41, f))
    X[0] = 0.0
    c = make_const(N, omegas)
    scale = 1.0 + 0.02 * N * N                      # phases reach |e t| ~ N^2 * 10 * 25: eps-level errors scale with it
    for site in (f - 1, 0):
        _compare(native, O, c, X, site, loss_atol=2e-10 * scale, grad_rtol=5e-8 * scale)
    prob = native.Problem.from_const(c)
    tr = prob.trace(X, f - 1).cpu().numpy()
    assert np.abs(tr - O.trace(c, X, f - 1)).max() < 3e-10 * scale * max(1, N)
    prob.close()


def test_register_and_shared_memory_large_d_paths_agree():
    """QTET_BIG_SMEM=1 keeps the older shared-memory kernels of the large-d path: both families must agree."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = f"""
import sys, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'quantum-targeted-energy-transfer_b200')!r})
from tet import _native
out = []
for N, om in [(10, [-3, 3, 3]), (60, [-3, 3])]:
    c = {{"max_N": N, "sites": len(om), "max_t": 25, "omegas": om, "chis": [0] * len(om), "coupling": 1, "timesteps": 30}}
    X = np.random.default_rng(5).uniform(-10, 10, (300, len(om)))
    p = _native.Problem.from_const(c)
    L, g, k = p.loss_grad_host(X, len(om) - 1)
    out.append((L, g, k))
np.savez(sys.argv[1], *[a for t in out for a in t])
"""
    import tempfile
    res = []
    for env in ({}, {"QTET_BIG_SMEM": "1"}):
        with tempfile.NamedTemporaryFile(suffix=".npz") as fh:
            out = subprocess.run([sys.executable, "-c", code, fh.name], env={**os.environ, **env}, capture_output=True, text=True, timeout=900)
            assert out.returncode == 0, out.stderr[-2000:]
            z = np.load(fh.name)
            res.append([z[k] for k in z.files])
    for a, b in zip(res[0], res[1]):
        if a.dtype.kind == "i":
            assert (a == b).mean() > 0.995
        else:
            assert np.abs(a - b).max() < 1e-6 * max(1.0, np.abs(b).max())


def test_householder_variants_agree():
    """Dense large-d systems: the staged tiled Householder reduction (default), the same in one kernel (QTET_HH_STAGED=0), the
    one-barrier-per-column variant (QTET_HH_FRAG=2) and round 1's one-row-per-thread kernel (QTET_HH_FRAG=0) are four
    implementations of the same reduction -- losses, gradients and arg-max samples must agree to rounding."""
    import subprocess
    import sys
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = f"""
import sys, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'quantum-targeted-energy-transfer_b200')!r})
from tet import _native
out = []
for N, om in [(10, [-3, 3, 3]), (6, [-3, 0, 1, 3]), (9, [-3, 3, 3])]:        # d = 66, 84 (both staged), 55 (staged, ragged tiles)
    c = {{"max_N": N, "sites": len(om), "max_t": 25, "omegas": om, "chis": [0] * len(om), "coupling": 1, "timesteps": 30}}
    X = np.random.default_rng(7).uniform(-10, 10, (200, len(om)))
    p = _native.Problem.from_const(c)
    L, g, k = p.loss_grad_host(X, len(om) - 1)
    assert p.stats()["nonconverged"] == 0
    out.append((L, g, k))
np.savez(sys.argv[1], *[a for t in out for a in t])
"""
    res = []
    for env in ({}, {"QTET_HH_STAGED": "0"}, {"QTET_HH_FRAG": "2"}, {"QTET_HH_FRAG": "0"}):
        with tempfile.NamedTemporaryFile(suffix=".npz") as fh:
            out = subprocess.run([sys.executable, "-c", code, fh.name], env={**os.environ, **env}, capture_output=True, text=True, timeout=900)
            assert out.returncode == 0, (env, out.stderr[-2000:])
            z = np.load(fh.name)
            res.append([z[k] for k in z.files])
    for other in res[1:]:
        for a, b in zip(res[0], other):
            if a.dtype.kind == "i":
                assert (a == b).mean() > 0.99
            else:
                assert np.abs(a - b).max() < 1e-8 * max(1.0, np.abs(b).max())


def test_large_d_chunked_workspace_matches_single_chunk():
    """The large-d path sizes its workspace from the free HBM; QTET_BIG_WS_GIB=0.25 forces several chunks per call.
    Chunked and unchunked evaluation must agree bit for bit (same kernels, same points)."""
    import subprocess
    import sys
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = f"""
import sys, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'quantum-targeted-energy-transfer_b200')!r})
from tet import _native
c = {{"max_N": 10, "sites": 3, "max_t": 25, "omegas": [-3, 3, 3], "chis": [0, 0, 0], "coupling": 1, "timesteps": 30}}
X = np.random.default_rng(21).uniform(-10, 10, (9000, 3))
p = _native.Problem.from_const(c)
L, g, k = p.loss_grad_host(X, 2)
tr = p.trace(X[:700], 0).cpu().numpy()
np.savez(sys.argv[1], L=L, g=g, k=k, tr=tr)
"""
    res = []
    for env in ({}, {"QTET_BIG_WS_GIB": "0.25"}):          # 0.25 GiB / 3.6 MB per batch -> ~70 batches = 2240 points per chunk
        with tempfile.NamedTemporaryFile(suffix=".npz") as fh:
            out = subprocess.run([sys.executable, "-c", code, fh.name], env={**os.environ, **env}, capture_output=True, text=True, timeout=900)
            assert out.returncode == 0, out.stderr[-2000:]
            z = np.load(fh.name)
            res.append({k: z[k] for k in z.files})
    for key in ("L", "g", "k", "tr"):
        assert np.array_equal(res[0][key], res[1][key]), key
    N = 10
    assert np.all(res[0]["L"] > -1e-9) and np.all(res[0]["L"] < N + 1e-9)


def test_large_d_conservation_at_scale(native):
    """BASELINE configs 3/4 shapes (reduced batch): boson-number conservation sum_k <n_k(t)> = N and <n_0(0)> = N."""
    for N, om, B in ((10, (-3, 3, 3), 1 << 15), (100, (-3, 3), 1 << 13)):
        c = make_const(N, om)
        X = np.random.default_rng(N).uniform(-10, 10, (B, len(om)))
        prob = native.Problem.from_const(c)
        tot = None
        for site in range(len(om)):
            tr = prob.trace(X, site)
            tot = tr if tot is None else tot + tr
            if site == 0:
                assert float((tr[:, 0] - N).abs().max()) < 1e-9 * N
        assert float((tot - N).abs().max()) < 2e-8 * max(1, N / 10)
        loss, grad, ts = prob.loss_grad(X)
        assert bool((loss > -1e-9).all()) and bool((loss < N + 1e-9).all()) and bool(grad.isfinite().all())
        prob.close()


def test_golden_reference_vectors(native, golden):
    # values produced by the reference's own HamiltonianLoss.py (c128 stand-in arithmetic)
    for r in golden["loss"]:
        prob = native.Problem.from_const(r["const"])
        loss, _, _ = prob.loss_grad(np.array([r["chis"]]), r["site"])
        tr = prob.trace(np.array([r["chis"]]), r["site"])[0].cpu().numpy()
        assert abs(float(loss[0]) - r["loss_c128"]) < 1e-10
        assert np.abs(tr - np.array(r["trace_c128"])).max() < 1e-10
        # and within float32 noise of the reference's real (complex64) arithmetic
        assert abs(float(loss[0]) - r["loss_c64"]) < 1e-4
        prob.close()


def test_basis_and_hamiltonian(native, O, golden):
    for b in golden["basis"]:
        prob = native.Problem(b["max_N"], b["sites"], [-3, 3] + [3] * (b["sites"] - 2), 1.0, 25, 30)
        assert prob.dim == b["dim"]
        assert prob.states().astype(int).tolist() == b["states"]
        prob.close()
    for h in golden["hamiltonian"]:
        prob = native.Problem.from_const(h["const"])
        H = prob.hamiltonian(h["chis"])
        Href = np.array([float.fromhex(x) for x in h["H_hex"]]).reshape(H.shape)
        assert np.abs(H - Href).max() <= 4e-16 * max(1.0, np.abs(Href).max())      # fma contraction only
        prob.close()


def test_degenerate_and_resonant_points(native, O):
    # exact TET resonance (degenerate diagonal), chi = 0, symmetric points, huge nonlinearity
    c = make_const(20)
    X = np.array([[0.3, -0.3], [0.0, 0.0], [10.0, 10.0], [-10.0, -10.0], [10.0, -10.0], [1e-9, -1e-9],
                  [6.0 / 20, -6.0 / 20], [0.6, 0.0], [0.0, -0.6], [5.0, 5.0]])
    _compare(native, O, c, X, 1)
    _compare(native, O, c, X, 0)
    c3 = make_const(4, (-3, 3, 3))
    X3 = np.array([[0.0, 0.0, 0.0], [1.5, 0.0, -1.5], [2.0, 2.0, 2.0], [10.0, -10.0, 10.0]])
    _compare(native, O, c3, X3, 2, loss_atol=2e-10, grad_rtol=5e-8)


def test_trace_and_conservation_full_size(native):
    # BASELINE config 2 size sample: sum over sites of <n_k(t)> = N for every t, <n_0(0)> = N, bounds
    c = make_const(20)
    ax = np.linspace(-10, 10, 1024)
    rng = np.random.default_rng(5)
    X = np.stack([ax[rng.integers(0, 1024, 20000)], ax[rng.integers(0, 1024, 20000)]], axis=1)
    prob = native.Problem.from_const(c)
    t0 = prob.trace(X, 0).cpu().numpy()
    t1 = prob.trace(X, 1).cpu().numpy()
    assert np.abs(t0 + t1 - 20).max() < 1e-9
    assert np.abs(t0[:, 0] - 20).max() < 1e-11 and np.abs(t1[:, 0]).max() < 1e-11
    loss, grad, ts = prob.loss_grad(X, 1)
    loss = loss.cpu().numpy()
    assert (loss > -1e-9).all() and (loss < 20 + 1e-9).all()
    assert np.abs((20 - t1.max(axis=1)) - loss).max() < 1e-12
    assert (t1.argmax(axis=1) == ts.cpu().numpy()).all()
    prob.close()


def test_host_entry_point_and_edge_batches(native, O):
    c = make_const(5)
    prob = native.Problem.from_const(c)
    rng = np.random.default_rng(2)
    for B in (1, 2, 31, 33, 1000):
        X = rng.uniform(-4, 4, (B, 2))
        L, g, k = prob.loss_grad_host(X, 1)
        Lr, gr, kr = O.loss_grad_batch(c, X, 1)
        assert np.abs(L - Lr).max() < 1e-10 and np.abs(g - gr).max() < 1e-8 and (k == kr).all()
    # empty batch is a no-op
    L, g, k = prob.loss_grad_host(np.zeros((0, 2)), 1)
    assert L.shape == (0,)
    with pytest.raises(ValueError):
        prob.loss_grad(np.zeros((3, 3)))
    with pytest.raises(ValueError):
        prob.loss_grad(np.zeros((3, 2)), site=2)
    with pytest.raises(ValueError):
        prob.loss_grad(np.zeros((3, 2)), site=1.0)
    assert float(prob.loss_grad(np.array([[1.0, -1.0]]), "x1")[0][0]) == float(prob.loss_grad(np.array([[1.0, -1.0]]), 1)[0][0])
    prob.close()


def test_split_and_generic_paths_agree(native):
    c = make_const(20)
    rng = np.random.default_rng(11)
    X = rng.uniform(-10, 10, (4096, 2))
    p = native.Problem.from_const(c)
    assert p.get_path() == native.QTET_PATH_DIRECT          # what AUTO picks for a dimer with hopping
    p.set_path(native.QTET_PATH_GENERIC)
    b = [t.cpu().numpy() for t in p.loss_grad(X, 1)]
    for path in (native.QTET_PATH_DIRECT, native.QTET_PATH_SPLIT):
        p.set_path(path)
        a = [t.cpu().numpy() for t in p.loss_grad(X, 1)]
        assert np.abs(a[0] - b[0]).max() < 2e-9
        same = a[2] == b[2]
        assert same.mean() > 0.999
        assert (np.abs(a[1] - b[1]) / np.maximum(1, np.abs(b[1]).max(axis=1, keepdims=True)))[same].max() < 1e-7
    p.close()


def test_stream_overflow_falls_back_to_generic_kernel(tmp_path):
    """Batches whose rotation stream does not fit its slot are recomputed by the fused generic kernel without a
    host round trip.  QTET_SPLIT_CAPIT (test hook) shrinks the slot so that EVERY batch overflows."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = f"""
import sys, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'quantum-targeted-energy-transfer_b200')!r})
from tet import _native
from oracle import qtet_oracle as O
c = {{"max_N": 20, "sites": 2, "max_t": 25, "omegas": [-3, 3], "chis": [0, 0], "coupling": 1, "timesteps": 30}}
X = np.random.default_rng(9).uniform(-10, 10, (333, 2))
p = _native.Problem.from_const(c)
p.set_path(_native.QTET_PATH_SPLIT)
L, g, k = p.loss_grad_host(X, 1)
Lr, gr, kr = O.loss_grad_batch(c, X, 1)
assert np.abs(L - Lr).max() < 2e-9 and (k == kr).all() and np.abs(g - gr).max() < 1e-7 * max(1, np.abs(gr).max())
print('ok')
"""
    out = subprocess.run([sys.executable, "-c", code], env={**os.environ, "QTET_SPLIT_CAPIT": "7"},
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_large_dimension_limits(native, O):
    """d = 115 (dimer N = 114) is the largest Fock space whose V and K fit shared memory.  Beyond it the fused generic
    kernel keeps the two matrices in a global-memory slab (the reference has no size limit, HamiltonianLoss.py:39-47):
    trimer N = 14 (d = 120, dense H) and dimer N = 128 (d = 129) must return the oracle's numbers, not an error."""
    c = make_const(114)
    X = np.array([[0.05, -0.05], [1.0, 2.0]])
    p = native.Problem.from_const(c)
    assert p.dim == 115 and p.get_path() == native.QTET_PATH_SPLIT
    L, g, k = p.loss_grad_host(X, 1)
    Lr, gr, kr = O.loss_grad_batch(c, X, 1)
    assert np.abs(L - Lr).max() < 5e-9 and (k == kr).all()
    p.close()
    rng = np.random.default_rng(11)
    for N, omegas in [(128, (-3, 3)), (14, (-3, 3, 3))]:
        c = make_const(N, omegas)
        f = len(omegas)
        X = np.concatenate([rng.uniform(-2.0, 2.0, (7, f)), np.array([[6.0 / N, -6.0 / N] + [0.0] * (f - 2)])])
        p = native.Problem.from_const(c)
        assert p.dim > 115 and p.get_path() == native.QTET_PATH_GENERIC
        p.close()
        _compare(native, O, c, X, f - 1, loss_atol=2e-9, grad_rtol=2e-6)      # tolerances of test_dimer_large_fock_space
    with pytest.raises(native.QtetError, match="too large"):
        native.Problem.from_const(make_const(60, (-3, 3, 3)))     # d = 1891


@pytest.mark.parametrize("N,omegas,coupling,max_t,timesteps", [
    (6, (-3, 3), 0.0, 25, 30),        # no hopping: H diagonal, nothing ever moves
    (6, (-3, 3), -2.5, 25, 30),       # negative coupling
    (12, (-1.5, 0.5), 4.0, 10, 7),    # strong coupling, short grid
    (5, (-3, 3), 1.0, 25, 1),         # a single time sample (t = 0)
    (5, (-3, 3), 1.0, 25, 100),       # long grid
    (5, (-3, 3), 1.0, 12.9, 30),      # max_t is truncated to int32 like HamiltonianLoss.py:29
    (1, (-3, 3), 1.0, 25, 30),        # one boson: d = 2
    (3, (-3, 3, 3), 0.0, 25, 30),     # dense path without hopping
    (40, (-3, 3), 1.0, 25, 30),       # large-d split path, tridiagonal
    (40, (-3, 3), 1.0, 25, 1),        # large-d, DMMA evolution: a single sample (chunks of 8 samples, 7 padded)
    (36, (-3, 3), 1.0, 25, 13),       # large-d, DMMA evolution: ragged last chunk
    (5, (-3, 3, 3), 1.0, 25, 67),     # dense large-d (d = 21 -> 32), nine chunks, the last one ragged
    (7, (-3, 3, 3), 1.5, 20, 8),      # dense large-d (d = 36 -> 48), exactly one chunk
])
def test_unusual_constants(native, O, N, omegas, coupling, max_t, timesteps):
    c = make_const(N, omegas, coupling=coupling, max_t=max_t, timesteps=timesteps)
    rng = np.random.default_rng(17)
    X = rng.uniform(-6, 6, (70, len(omegas)))
    X[0] = 0.0
    for site in (len(omegas) - 1, 0):
        _compare(native, O, c, X, site, loss_atol=3e-10, grad_rtol=1e-7)
    prob = native.Problem.from_const(c)
    tr = prob.trace(X, 0).cpu().numpy()
    assert tr.shape == (70, timesteps) and np.abs(tr - O.trace(c, X, 0)).max() < 3e-10 * max(1, N)
    prob.close()
