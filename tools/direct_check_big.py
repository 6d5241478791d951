#!/usr/bin/env python3
"""Large-d direct path (qtet_big.cu / qtet_bigreg.cu DIRECT variant) on a GPU box: parity against the oracle for every padded
size (tridiagonal and dense), agreement with the rotation-stream path, and timings of BASELINE's dimer N=100 / trimer N=10."""
import os, sys
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
    try:
        p.set_path(path)
    except nat.QtetError as e:
        print(f"{tag:40s} path {path} not eligible: {e}"); p.close(); return
    L, g, k = [t.cpu().numpy() for t in p.loss_grad(X, site)]
    Lr, gr, kr = O.loss_grad_batch(c, X, site)
    same = k == kr
    gs = np.maximum(1.0, np.abs(gr).max(axis=1, keepdims=True))
    ge = (np.abs(g - gr) / gs)[same].max() if same.any() else 0.0
    print(f"{tag:40s} d={p.dim:3d} path={nat.PATH_NAMES[p.get_path()]:7s} n={len(X):5d} max|dL|={np.abs(L - Lr).max():.2e} rel|dg|={ge:.2e} "
          f"t* mismatches={int((~same).sum())} finite={bool(np.isfinite(L).all() and np.isfinite(g).all())} stats={p.stats()}", flush=True)
    p.close()


rng = np.random.default_rng(0)
for N, om in [(33, (-3, 3)), (50, (-3, 3)), (70, (-3, 3)), (90, (-3, 3)), (100, (-3, 3)), (3, (-3, 3, 3)), (6, (-3, 3, 3)), (8, (-3, 1, 3)),
              (10, (-3, 3, 3)), (5, (-3, 0, 1, 3)), (6, (-3, 0, 1, 3)), (12, (-3, 3, 3)), (4, (-3, 3, 3)), (2, (-3, 3, 3))]:
    f = len(om)
    X = rng.uniform(-10, 10, (65, f)); X[0] = 0.0
    for site in (f - 1, 0):
        cmp(const(N, om), X, site, nat.QTET_PATH_DIRECT, f"N={N} om={om} site={site}")
# clustered spectra: symmetric trimer / dimer
hard = np.array([[0.5, 0.5], [2, 2], [5, 5], [1.0, 1.000000001], [0.06, -0.06], [0, 0]])
cmp(const(60, (3, 3)), hard, 1, nat.QTET_PATH_DIRECT, "dimer N=60 symmetric, hard points")
hard3 = np.array([[1, 1, 1], [2, 0, 2], [0, 0, 0], [5, 5, 5], [0.6, 0, -0.6], [3, 3.0000001, 3]])
cmp(const(10, (3, 3, 3)), hard3, 2, nat.QTET_PATH_DIRECT, "trimer N=10 symmetric, hard points")
cmp(const(10, (-3, 3, 3)), hard3, 2, nat.QTET_PATH_DIRECT, "trimer N=10, hard points")

for name, c, B in [("dimer100", const(100), 1 << 16), ("trimer10", const(10, (-3, 3, 3)), 1 << 16)]:
    f = c["sites"]
    X = torch.from_numpy(np.random.default_rng(0).uniform(-10, 10, (B, f))).cuda()
    out = (torch.empty(B, dtype=torch.float64, device="cuda"), torch.empty(B, f, dtype=torch.float64, device="cuda"), torch.empty(B, dtype=torch.int32, device="cuda"))
    res = {}
    for path in (nat.QTET_PATH_SPLIT, nat.QTET_PATH_DIRECT):
        p = nat.Problem.from_const(c); p.set_path(path)
        for _ in range(2): p.loss_grad(X, out=out)
        torch.cuda.synchronize()
        p.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): p.loss_grad(X, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        kern = p.profile_read(); p.profile(False)
        res[path] = [t.clone() for t in out]
        print(f"{name} B={B} path={nat.PATH_NAMES[path]}: {ms:.3f} ms/step  {B / ms * 1e3:.3e} evals/s  kernels(ms/step)={ {k: round(v[0] / 3, 3) for k, v in kern.items()} } stats={p.stats()}", flush=True)
        p.close()
    a, b = res[nat.QTET_PATH_SPLIT], res[nat.QTET_PATH_DIRECT]
    print(f"{name}: direct vs split: max|dL| = %.3e, t* differ at %d points, max rel|dg| = %.3e" % (
        (a[0] - b[0]).abs().max().item(), int((a[2] != b[2]).sum()),
        (((a[1] - b[1]).abs() / a[1].abs().amax(1, keepdim=True).clamp(min=1.0))[a[2] == b[2]]).max().item()))
