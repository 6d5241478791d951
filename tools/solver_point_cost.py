#!/usr/bin/env python3
"""Why config 5 runs below the grid rate: loss+grad rate on the points the optimiser actually visits (the positions of the 10^5
guesses after 50 / 300 / 1000 epochs) against the rate on the uniform initial guesses and on the config-2 grid."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantum-targeted-energy-transfer_b200"))
import torch
from tet import _native as nat
from tet.solver_mp import randomCombinations
c = {"max_N": 20, "sites": 2, "max_t": 25, "omegas": [-3, 3], "chis": [0, 0], "coupling": 1, "timesteps": 30}
p = nat.Problem.from_const(c)
G = 100000
x0 = randomCombinations({"x0lims": [-10, 10], "x1lims": [-10, 10]}, [0, 1], G, seed=0)


def rate(X, label):
    X = torch.from_numpy(np.ascontiguousarray(X)).cuda()
    out = (torch.empty(len(X), dtype=torch.float64, device="cuda"), torch.empty(len(X), 2, dtype=torch.float64, device="cuda"),
           torch.empty(len(X), dtype=torch.int32, device="cuda"))
    for _ in range(3): p.loss_grad(X, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): p.loss_grad(X, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{label:46s} {len(X):7d} points  {ms:7.3f} ms  {len(X) / ms * 1e3:.4e} evals/s", flush=True)


rate(x0, "uniform initial guesses")


def binsort(X, nb=256):
    b0 = np.clip(((X[:, 0] - X[:, 0].min()) / np.ptp(X[:, 0]) * (nb - 1e-3)).astype(int), 0, nb - 1)
    b1 = np.clip(((X[:, 1] - X[:, 1].min()) / np.ptp(X[:, 1]) * (nb - 1e-3)).astype(int), 0, nb - 1)
    b1 = np.where(b0 & 1, nb - 1 - b1, b1)
    return X[np.argsort(b0 * nb + b1, kind="stable")]


rate(binsort(x0), "uniform initial guesses, bin-sorted 256x256")
rate(binsort(x0, 64), "uniform initial guesses, bin-sorted 64x64")
rate(x0[np.lexsort((x0[:, 1], x0[:, 0]))], "uniform initial guesses, lexsorted")
for it in (50, 300, 1000):
    out = p.adam_run(x0, train_sites=(0, 1), iterations=it, lr=0.1)
    bv = out["best_vars"].cpu().numpy()
    rate(bv, f"best positions after {it} epochs")
    rate(binsort(bv), f"best positions after {it} epochs, bin-sorted")
ax = np.linspace(-10, 10, 1024)
xd, xa = np.meshgrid(ax, ax, indexing="ij")
rate(np.stack([xd.ravel(), xa.ravel()], 1)[:G], "first 10^5 points of the config-2 grid")
rate(np.stack([xd.ravel(), xa.ravel()], 1), "whole config-2 grid")
