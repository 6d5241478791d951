"""CPU oracle for the QTET loss+gradient hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a plain numpy (FP64 / complex128) restatement of the reference's algorithm
for the hot path named in BASELINE.json.  It is *not* part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the CPU baseline.
The product path (``quantum-targeted-energy-transfer_b200/``) never imports it and fails
loudly when its CUDA library is missing.

Parity pinning: the reference ships no tests, golden vectors or fixtures.
This oracle is pinned instead against outputs of the reference's own, unmodified
``src/tet/HamiltonianLoss.py`` executed in the build container on a numpy stand-in for the
~20 ``tf.*`` symbols it uses (``tests/golden/make_golden.py`` -> ``tests/golden/*.json``):
the Fock basis and the Hamiltonian matrices agree bit-for-bit, the loss / trace agree to
the precision of the reference's complex64 evolution (~1e-5).  TensorFlow itself is absent
from the image, so the gradient is pinned against central differences, torch-CPU autograd
through ``eigh`` and mpmath at 50 digits -- not against TF's tape: "parity pinned on
forward values, gradient pinned on independent derivations".

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from itertools import product

import numpy as np

__all__ = [
    "fock_dim", "fock_basis", "offdiag_matrix", "hamiltonian", "time_grid",
    "trace", "loss", "loss_grad", "loss_grad_batch", "adam_train", "get_combinations",
    "solve", "algorithmic_flops",
]


# ----------------------------------------------------------------------------------------
# Fock basis                                         src/tet/HamiltonianLoss.py:39-40, 62-110
# ----------------------------------------------------------------------------------------
def fock_dim(max_N: int, sites: int) -> int:
    """dim = C(N+f-1, f-1)  (HamiltonianLoss.py:39-40 computes it by float factorials)."""
    return math.comb(max_N + sites - 1, sites - 1)


def fock_basis(max_N: int, sites: int) -> np.ndarray:
    """All occupation vectors with sum N in the order ``Loss.derive`` produces
    (HamiltonianLoss.py:62-110): row 0 is (N,0,..,0) and rows follow in descending
    lexicographic order.  Returned as float64 [dim, f] like ``Loss.states``."""
    out = []

    def rec(prefix, remaining, k):
        if k == sites - 1:
            out.append(prefix + [remaining])
            return
        for n in range(remaining, -1, -1):
            rec(prefix + [n], remaining - n, k + 1)

    rec([], max_N, 0)
    states = np.array(out, dtype=np.float64)
    assert states.shape == (fock_dim(max_N, sites), sites)
    return states


# ----------------------------------------------------------------------------------------
# Hamiltonian                                           src/tet/HamiltonianLoss.py:124-176
# ----------------------------------------------------------------------------------------
def offdiag_matrix(states: np.ndarray, coupling: float) -> np.ndarray:
    """The chi-independent part of H (HamiltonianLoss.py:136-163): for every basis state m
    and bond k, moving one boson k->k+1 gives state n with H[n,m] -= lam*sqrt((n_{k+1}+1) n_k)
    and moving one k+1->k gives H[n,m] -= lam*sqrt(n_{k+1} (n_k+1)).  The reference finds n
    through a float hash + searchsorted (:112-121, :148, :160); a dict lookup is identical."""
    d, f = states.shape
    index = {tuple(int(x) for x in s): i for i, s in enumerate(states)}
    H = np.zeros((d, d))
    for m in range(d):
        s = [int(x) for x in states[m]]
        for k in range(f - 1):
            nk, nk1 = s[k], s[k + 1]
            if nk != 0:                       # term 2a (:143-151)
                t = list(s); t[k] -= 1; t[k + 1] += 1
                H[index[tuple(t)], m] -= coupling * np.sqrt((nk1 + 1) * nk)
            if nk1 != 0:                      # term 2b (:154-163)
                t = list(s); t[k] += 1; t[k + 1] -= 1
                H[index[tuple(t)], m] -= coupling * np.sqrt(nk1 * (nk + 1))
    return H


def diag_energies(states: np.ndarray, omegas, chis) -> np.ndarray:
    """H_mm = sum_k omega_k n_k + 0.5 chi_k n_k^2   (HamiltonianLoss.py:131-134).
    ``chis`` may be [f] or [B, f]; returns [d] or [B, d]."""
    omegas = np.asarray(omegas, dtype=np.float64)
    chis = np.asarray(chis, dtype=np.float64)
    acc = np.zeros(chis.shape[:-1] + (states.shape[0],))
    for k in range(states.shape[1]):      # same summation order as the reference loop (:132-134)
        acc = acc + (omegas[k] * states[:, k] + (0.5 * chis[..., k:k + 1]) * states[:, k] ** 2)
    return acc


def hamiltonian(const: dict, chis) -> np.ndarray:
    """Dense H [d,d] (or [B,d,d]) as ``Loss.createHamiltonian`` stacks it (:168-176)."""
    states = fock_basis(const["max_N"], const["sites"])
    off = offdiag_matrix(states, const["coupling"])
    dg = diag_energies(states, const["omegas"], chis)
    d = states.shape[0]
    if dg.ndim == 1:
        return off + np.diag(dg)
    H = np.broadcast_to(off, (dg.shape[0], d, d)).copy()
    idx = np.arange(d)
    H[:, idx, idx] += dg
    return H


def time_grid(const: dict) -> np.ndarray:
    """np.linspace(0, max_t, timesteps) with max_t cast to int32 (HamiltonianLoss.py:29, 221)."""
    return np.linspace(0, np.int32(const["max_t"]), const["timesteps"])


def _target(site, sites: int) -> int:
    """HamiltonianLoss.py:49-60: int, or a string whose last character is the site."""
    if type(site) is int or isinstance(site, (np.integer,)):
        return int(site)
    if type(site) is str:
        return int(site[-1])
    raise ValueError("Invalid type for site variable. Must be int.")


# ----------------------------------------------------------------------------------------
# Evolution, loss, gradient        src/tet/HamiltonianLoss.py:179-233, src/tet/Optimizer.py:85-91
# ----------------------------------------------------------------------------------------
class _Problem:
    """Everything that does not depend on chi (cached per const)."""
    _cache: dict = {}

    def __init__(self, const):
        self.N = int(const["max_N"]); self.f = int(const["sites"])
        self.states = fock_basis(self.N, self.f)
        self.d = self.states.shape[0]
        self.off = offdiag_matrix(self.states, const["coupling"])
        self.omegas = np.asarray(const["omegas"], dtype=np.float64)
        self.lin = self.states @ self.omegas
        self.const = dict(const)
        self.quad = 0.5 * self.states ** 2            # dH_mm/dchi_k  [d, f]
        self.t = time_grid(const)

    @classmethod
    def get(cls, const):
        key = (const["max_N"], const["sites"], float(const["coupling"]),
               tuple(float(x) for x in const["omegas"]), float(const["max_t"]),
               int(const["timesteps"]))
        p = cls._cache.get(key)
        if p is None:
            p = cls._cache[key] = cls(const)
        return p


def _eigh_batch(P: _Problem, chis: np.ndarray):
    chis = np.atleast_2d(np.asarray(chis, dtype=np.float64))
    B = chis.shape[0]
    H = np.broadcast_to(P.off, (B, P.d, P.d)).copy()
    idx = np.arange(P.d)
    H[:, idx, idx] += diag_energies(P.states, P.omegas, chis)
    return np.linalg.eigh(H)                           # HamiltonianLoss.py:182


def trace(const: dict, chis, site=None, c64: bool = False) -> np.ndarray:
    """<n_site(t)> on the time grid, [T] or [B,T]  (HamiltonianLoss.py:204-226,
    ``single_value=False``).  psi0 = basis state 0 so c = V[0,:] (:186-198).
    ``c64=True`` emulates the reference's complex64 casts (:205-210)."""
    P = _Problem.get(const)
    site = P.f - 1 if site is None else _target(site, P.f)
    single = np.ndim(chis) == 1
    e, V = _eigh_batch(P, chis)
    c = V[:, 0, :]                                               # [B,d]
    if c64:
        e32 = e.astype(np.complex64); t32 = P.t.astype(np.complex64)
        ph = np.exp(np.complex64(-1j) * e32[:, None, :] * t32[None, :, None])
        z = c.astype(np.complex64)[:, None, :] * ph
        psi = np.einsum("bji,bti->btj", V.astype(np.complex64), z)
        n = (np.abs(psi) ** 2 * P.states[:, site].astype(np.float32)).sum(-1).astype(np.float64)
    else:
        ph = np.exp(-1j * e[:, None, :] * P.t[None, :, None])      # [B,T,d]
        psi = np.einsum("bji,bti->btj", V, c[:, None, :] * ph)     # [B,T,d]
        n = (np.abs(psi) ** 2 * P.states[:, site]).sum(-1)
    return n[0] if single else n


def loss(const: dict, chis, site=None, c64: bool = False):
    """Scalar loss (HamiltonianLoss.py:227-231): N - max_t<n> for the acceptor (site == f-1),
    min_t<n> for any other site."""
    P = _Problem.get(const)
    s = P.f - 1 if site is None else _target(site, P.f)
    n = trace(const, chis, s, c64=c64)
    return (P.N - n.max(-1)) if s == P.f - 1 else n.min(-1)


def _dk_phi(e: np.ndarray, t: float) -> np.ndarray:
    """Daleckii-Krein divided differences of f(x)=exp(-i x t):
    Phi_ij = (f_i-f_j)/(e_i-e_j), Phi_ii = -i t f_i, written degeneracy-safe as
    -i t exp(-i (e_i+e_j) t/2) sinc((e_i-e_j) t/2)."""
    s = (e[:, None] + e[None, :]) * (0.5 * t)
    x = (e[:, None] - e[None, :]) * (0.5 * t)
    return -1j * t * np.exp(-1j * s) * np.sinc(x / np.pi)


def loss_grad(const: dict, chis, site=None):
    """(loss, grad[f], t_star) for one point: d loss / d chi_k with the arg-max (acceptor) /
    arg-min sample held fixed -- the sub-gradient TF's reduce_max/min back-propagates
    (Optimizer.py:85-91).  Only H's diagonal depends on chi: dH_mm/dchi_k = n_k(m)^2/2
    (HamiltonianLoss.py:133-134)."""
    P = _Problem.get(const)
    s = P.f - 1 if site is None else _target(site, P.f)
    chis = np.asarray(chis, dtype=np.float64)
    H = P.off + np.diag(diag_energies(P.states, P.omegas, chis))
    e, V = np.linalg.eigh(H)
    c = V[0, :]
    ph = np.exp(-1j * e[None, :] * P.t[:, None])
    psi = (c[None, :] * ph) @ V.T                                  # [T,d]
    w = P.states[:, s]
    n = (np.abs(psi) ** 2 * w).sum(-1)
    acceptor = s == P.f - 1
    k = int(np.argmax(n)) if acceptor else int(np.argmin(n))
    t = P.t[k]
    a = V.T @ (w * np.conj(psi[k]))                                # [d] complex
    M = (a[:, None] * c[None, :]) * _dk_phi(e, t)                  # [d,d]
    K = np.real(M)
    G = 2.0 * np.einsum("mi,ij,mj->m", V, K, V)                    # d<n(t*)>/dH_mm
    dn = P.quad.T @ G
    if acceptor:
        return P.N - n[k], -dn, k
    return n[k], dn, k


def loss_grad_batch(const: dict, chis, site=None):
    """Vectorised version of :func:`loss_grad` over chis [B,f] -> (loss[B], grad[B,f], t_star[B])."""
    P = _Problem.get(const)
    s = P.f - 1 if site is None else _target(site, P.f)
    chis = np.atleast_2d(np.asarray(chis, dtype=np.float64))
    B = chis.shape[0]
    e, V = _eigh_batch(P, chis)
    c = V[:, 0, :]
    ph = np.exp(-1j * e[:, None, :] * P.t[None, :, None])          # [B,T,d]
    psi = np.einsum("bji,bti->btj", V, c[:, None, :] * ph)         # [B,T,d]
    w = P.states[:, s]
    n = (np.abs(psi) ** 2 * w).sum(-1)                             # [B,T]
    acceptor = s == P.f - 1
    k = n.argmax(-1) if acceptor else n.argmin(-1)
    ar = np.arange(B)
    t = P.t[k]                                                     # [B]
    psik = psi[ar, k]                                              # [B,d]
    a = np.einsum("bji,bj->bi", V, w * np.conj(psik))              # [B,d]
    ssum = (e[:, :, None] + e[:, None, :]) * (0.5 * t)[:, None, None]
    x = (e[:, :, None] - e[:, None, :]) * (0.5 * t)[:, None, None]
    phi = (-1j * t)[:, None, None] * np.exp(-1j * ssum) * np.sinc(x / np.pi)
    K = np.real(a[:, :, None] * c[:, None, :] * phi)
    G = 2.0 * np.einsum("bmi,bij,bmj->bm", V, K, V)
    dn = G @ P.quad                                                # [B,f]
    nk = n[ar, k]
    if acceptor:
        return P.N - nk, -dn, k.astype(np.int32)
    return nk, dn, k.astype(np.int32)


def algorithmic_flops(d: int, T: int, tridiagonal: bool) -> float:
    """SURVEY.md §8(d) convention: F = F_eigh + 4 T d^2 + 4 d^3, F_eigh = 6 d^3 (tridiagonal)
    or 9 d^3 (needs reduction)."""
    return (6.0 if tridiagonal else 9.0) * d ** 3 + 4.0 * T * d ** 2 + 4.0 * d ** 3


# ----------------------------------------------------------------------------------------
# Per-guess optimiser loop                                     src/tet/Optimizer.py:99-242
# ----------------------------------------------------------------------------------------
def adam_train(const: dict, init_chis, target_site=None, train_sites=(0, 1), iterations=1000,
               lr=0.1, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad=False, tol=1e-8,
               loss_grad_fn=None):
    """Restates ``Optimizer.train`` for ONE initial guess, with keras Adam
    (m += (g-m)(1-b1); v += (g^2-v)(1-b2); a = lr sqrt(1-b2^t)/(1-b1^t); x -= a m/(sqrt(v)+eps)
    -- 3rd-party, TF/Keras `Adam._resource_apply_dense`).  Returns the dict
    ``Optimizer._train`` builds (Optimizer.py:259-263) plus 'epochs'.

    Order of operations per epoch follows Optimizer.py:146-224 exactly:
    loss+grad -> lr:=1e-4 if loss<=0.1 (:165) -> record loss (:169) -> best tracking (:172-175)
    -> Adam step (:178) -> |dx| (:181) -> revert if loss jumped >=0.5 and loss>=0.5 (:184-186)
    -> trajectory sample every 10 steps (:189-192) -> break if |loss|<0.1 (:195)
    -> stop when a trained variable moved < tol more than twice (:202-206)."""
    f = len(const["chis"])
    if target_site is None:
        target_site = f - 1
    lg = loss_grad_fn or (lambda x: loss_grad(const, x, target_site)[:2])
    x = np.array([float(v) for v in init_chis], dtype=np.float64)
    train = [j for j in range(f) if j in train_sites]
    m = np.zeros(f); v = np.zeros(f); vhat = np.zeros(f)
    losses = [float(const["max_N"])]
    best = np.zeros(f); best_loss = float(const["max_N"])
    var_data = [[] for _ in range(f)]
    err_count = [0] * f
    counter = 0
    step = 0
    cur_lr = lr
    epoch = -1
    for epoch in range(iterations):
        prev = x.copy()
        L, g = lg(x)
        L = float(L)
        if L <= 0.1:
            cur_lr = 0.0001
        losses.append(L)
        if losses[epoch + 1] < min(losses[:epoch + 1]):
            best = x.copy(); best_loss = L
        step += 1
        a = cur_lr * math.sqrt(1.0 - beta_2 ** step) / (1.0 - beta_1 ** step)
        for j in train:
            m[j] += (g[j] - m[j]) * (1.0 - beta_1)
            v[j] += (g[j] * g[j] - v[j]) * (1.0 - beta_2)
            if amsgrad:
                vhat[j] = max(vhat[j], v[j])
                x[j] -= a * m[j] / (math.sqrt(vhat[j]) + epsilon)
            else:
                x[j] -= a * m[j] / (math.sqrt(v[j]) + epsilon)
        var_error = np.abs(x - prev)
        if losses[epoch + 1] - losses[epoch] >= 0.5 and L >= 0.5:
            x = prev.copy()
        counter += 1
        if counter % 10 == 0:
            for k in range(f):
                var_data[k].append(float(x[k]))
        if abs(L) < 0.1:
            break
        stop = False
        for j in range(f):
            if var_error[j] < tol and j in train_sites:
                err_count[j] += 1
                if err_count[j] > 2:
                    stop = True
                    break
        if stop:
            break
    return {"loss": losses[1:], "var_data": var_data, "best_vars": [float(b) for b in best],
            "best_loss": best_loss, "epochs": epoch + 1, "final_vars": x.copy()}


# ----------------------------------------------------------------------------------------
# Initial guesses and arg-min over guesses             src/tet/solver_mp.py:17-84, 182-184
# ----------------------------------------------------------------------------------------
def get_combinations(trainable_vars_limits: dict, train_sites=(0, 1), method="bins", grid=2, rng=None):
    """``getCombinations`` (solver_mp.py:17-84).  'grid': cartesian product of linspace per
    trained site (:75-84).  'bins': same points histogram-binned, one random pick per non-empty
    bin (:40-73; the reference draws from the unseeded global numpy RNG, here ``rng``)."""
    if method not in ("grid", "bins"):
        raise ValueError("Provided method not in list of supported methods ['grid', 'bins']")
    spans = [np.linspace(trainable_vars_limits[f"x{i}lims"][0], trainable_vars_limits[f"x{i}lims"][1], grid)
             for i in train_sites]
    if method == "grid":
        return list(product(*spans))
    data = np.array(list(product(*spans)))
    data_list = [data[:, i] for i in range(data.shape[1])]
    extents = [(trainable_vars_limits[f"x{i}lims"][0] - 0.1, trainable_vars_limits[f"x{i}lims"][1] + 0.1)
               for i in train_sites]
    _, edges = np.histogramdd(data_list, bins=grid, range=extents)
    hit = [np.digitize(data_list[i], edges[i]) for i in range(len(data_list))]
    hitbins = list(zip(*hit))
    rng = rng or np.random
    combos = []
    for _bin in product(*[range(1, grid + 1) for _ in train_sites]):
        items = [d for d, hb in zip(data, hitbins) if tuple(hb) == _bin]
        if items:
            combos.append(items[rng.choice(list(range(len(items))))])
    return combos


def solve(const: dict, combinations, target_site=None, train_sites=(0, 1), iterations=500,
          lr=0.1, beta_1=0.9, amsgrad=False):
    """mp_opt over every guess + the arg-min of solver_mp.py:182-184.
    Returns (all_losses [G, f+1], optimal_vars [f], min_loss)."""
    f = len(const["chis"])
    rows = []
    for comb in combinations:
        x0 = [0.0] * f                                  # Optimizer.py:307-312
        for idx, val in zip(train_sites, comb):
            x0[idx] = float(val)
        r = adam_train(const, x0, target_site, train_sites, iterations, lr, beta_1, amsgrad=amsgrad)
        rows.append([*r["best_vars"], min(r["loss"])])  # Optimizer.py:321
    all_losses = np.array(rows)
    j = int(np.argmin(all_losses[:, f]))
    return all_losses, [float(all_losses[j, i]) for i in range(f)], float(all_losses[j, f])
