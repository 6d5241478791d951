#!/usr/bin/env python3
"""A/B timing of one library build (QTET_B200_LIB=...): config-2 grid (dimer N=20), dimer N=100 and trimer N=10, direct path."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantum-targeted-energy-transfer_b200"))
import torch
from tet import _native as nat

tag = os.path.basename(os.environ.get("QTET_B200_LIB", "default"))
which = sys.argv[1:] or ["dimer20", "dimer100", "trimer10"]


def const(N, om):
    return {"max_N": N, "sites": len(om), "max_t": 25, "omegas": list(om), "chis": [0] * len(om), "coupling": 1, "timesteps": 30}


def run(name, c, X, steps):
    X = torch.from_numpy(X).cuda()
    B, f = X.shape
    out = (torch.empty(B, dtype=torch.float64, device="cuda"), torch.empty(B, f, dtype=torch.float64, device="cuda"), torch.empty(B, dtype=torch.int32, device="cuda"))
    p = nat.Problem.from_const(c)
    for _ in range(3): p.loss_grad(X, out=out)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps): p.loss_grad(X, out=out)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    p.profile(True)
    for _ in range(steps): p.loss_grad(X, out=out)
    kern = p.profile_read(); p.profile(False)
    print(f"[{tag}] {name}: path={nat.PATH_NAMES[p.get_path()]} {best:.3f} ms/step {B / best * 1e3:.4e} evals/s  kernels(ms/step)={ {k: round(v[0] / steps, 3) for k, v in kern.items()} } checksum={float(out[0].sum()):.9f}", flush=True)
    p.close()


if "dimer20" in which:
    ax = np.linspace(-10, 10, 1024)
    xd, xa = np.meshgrid(ax, ax, indexing="ij")
    run("dimer20 grid 1024^2", const(20, (-3, 3)), np.stack([xd.ravel(), xa.ravel()], 1), 10)
if "dimer100" in which:
    run("dimer100 2^16", const(100, (-3, 3)), np.random.default_rng(0).uniform(-10, 10, (1 << 16, 2)), 3)
if "trimer10" in which:
    run("trimer10 2^16", const(10, (-3, 3, 3)), np.random.default_rng(0).uniform(-10, 10, (1 << 16, 3)), 3)
