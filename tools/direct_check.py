#!/usr/bin/env python3
"""First-contact check of the direct path (qtet_direct.cu) on a GPU box: parity against the oracle on random, grid and
(nearly) degenerate points, agreement with the split path, and a quick timing of the config-2 grid on both paths."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "quantum-targeted-energy-transfer_b200"))
import torch
from tet import _native as nat
from oracle import qtet_oracle as O


def const(N, om=(-3, 3), coupling=1, max_t=25, T=30):
    return {"max_N": N, "sites": len(om), "max_t": max_t, "omegas": list(om), "chis": [0] * len(om), "coupling": coupling, "timesteps": T}


def cmp(c, X, site, path, tag):
    p = nat.Problem.from_const(c)
    p.set_path(path)
    L, g, k = [t.cpu().numpy() for t in p.loss_grad(X, site)]
    Lr, gr, kr = O.loss_grad_batch(c, X, site)
    same = k == kr
    gs = np.maximum(1.0, np.abs(gr).max(axis=1, keepdims=True))
    ge = (np.abs(g - gr) / gs)[same].max() if same.any() else 0.0
    print(f"{tag:42s} path={nat.PATH_NAMES[p.get_path()]:7s} n={len(X):6d} max|dL|={np.abs(L - Lr).max():.2e} rel|dg|={ge:.2e} "
          f"t* mismatches={int((~same).sum())} finite={bool(np.isfinite(L).all() and np.isfinite(g).all())}", flush=True)
    p.close()
    return L, g, k


rng = np.random.default_rng(0)
for N in (1, 2, 3, 4, 7, 15, 20, 23, 31):
    X = rng.uniform(-10, 10, (513, 2))
    for site in (1, 0):
        cmp(const(N), X, site, nat.QTET_PATH_DIRECT, f"dimer N={N} site={site} random")
# config-2 grid: a 64-row slab + the special lines
ax = np.linspace(-10, 10, 1024)
rows = np.r_[0:8, 508:516, 1016:1024]
xd, xa = np.meshgrid(ax[rows], ax, indexing="ij")
X = np.stack([xd.ravel(), xa.ravel()], 1)
cmp(const(20), X, 1, nat.QTET_PATH_DIRECT, "config-2 grid, 24 rows")
d = ax
lines = np.concatenate([np.stack([d, d], 1), np.stack([d, -d], 1), np.stack([d, 0 * d], 1), np.stack([0 * d, d], 1),
                        np.stack([d, 0 * d + 0.3], 1), np.stack([0 * d + 0.3, d], 1), np.stack([0 * d - 0.3, d], 1)])
cmp(const(20), lines, 1, nat.QTET_PATH_DIRECT, "config-2 special lines")
# exactly / nearly degenerate spectra (symmetric dimer): fallback + Gram-Schmidt
hard = np.array([[0.05, 0.05], [0.2, 0.2], [0.5, 0.5], [1, 1], [2, 2], [5, 5], [10, 10], [3, 3.0000001], [1.0, 1.000000001], [0.3, -0.3], [0, 0]])
for om in ((3, 3), (-3, 3), (0, 0)):
    cmp(const(20, om), hard, 1, nat.QTET_PATH_DIRECT, f"hard points omegas={om}")
    cmp(const(20, om), hard, 1, nat.QTET_PATH_SPLIT, f"hard points omegas={om}")
# unusual constants
cmp(const(20, (-3, 3), coupling=0.01), rng.uniform(-10, 10, (257, 2)), 1, nat.QTET_PATH_DIRECT, "coupling 0.01")
cmp(const(20, (-3, 3), coupling=7.5, max_t=3, T=67), rng.uniform(-10, 10, (257, 2)), 1, nat.QTET_PATH_DIRECT, "coupling 7.5, T=67")
cmp(const(12, (-1, 2), coupling=1, max_t=25, T=1), rng.uniform(-10, 10, (100, 2)), 0, nat.QTET_PATH_DIRECT, "T=1")

# timing
c = const(20)
xd, xa = np.meshgrid(ax, ax, indexing="ij")
X = torch.from_numpy(np.stack([xd.ravel(), xa.ravel()], 1)).cuda()
B = X.shape[0]
out = (torch.empty(B, dtype=torch.float64, device="cuda"), torch.empty(B, 2, dtype=torch.float64, device="cuda"), torch.empty(B, dtype=torch.int32, device="cuda"))
res = {}
for path in (nat.QTET_PATH_SPLIT, nat.QTET_PATH_DIRECT):
    p = nat.Problem.from_const(c); p.set_path(path)
    for _ in range(3): p.loss_grad(X, out=out)
    torch.cuda.synchronize()
    p.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): p.loss_grad(X, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    kern = p.profile_read(); p.profile(False)
    res[path] = [t.clone() for t in out]
    print(f"grid 1024^2 path={nat.PATH_NAMES[path]}: {ms:.3f} ms/step  {B / ms * 1e3:.3e} evals/s  kernels(ms/step)={ {k: round(v[0] / 10, 3) for k, v in kern.items()} }", flush=True)
    p.close()
a, b = res[nat.QTET_PATH_SPLIT], res[nat.QTET_PATH_DIRECT]
print("direct vs split on the full grid: max|dL| = %.3e, t* differ at %d points, max rel|dg| = %.3e" % (
    (a[0] - b[0]).abs().max().item(), int((a[2] != b[2]).sum()),
    (((a[1] - b[1]).abs() / a[1].abs().amax(1, keepdim=True).clamp(min=1.0))[a[2] == b[2]]).max().item()))
