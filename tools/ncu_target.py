#!/usr/bin/env python3
"""Small driver for ncu captures: one workload, two passes (the second one is the one to capture with --launch-skip).
    python tools/ncu_target.py dimer20|dimer100|trimer10 [points]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantum-targeted-energy-transfer_b200"))
import torch
from tet import _native as nat

name = sys.argv[1] if len(sys.argv) > 1 else "dimer20"
if name == "dimer20":
    c = {"max_N": 20, "sites": 2, "max_t": 25, "omegas": [-3, 3], "chis": [0, 0], "coupling": 1, "timesteps": 30}
    ax = np.linspace(-10, 10, 1024)
    xd, xa = np.meshgrid(ax, ax, indexing="ij")
    X = np.stack([xd.ravel(), xa.ravel()], 1)
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
    X = X[256 * 1024:256 * 1024 + n]                 # rows 256.. of the config-2 grid: a representative quarter
elif name == "dimer100":
    c = {"max_N": 100, "sites": 2, "max_t": 25, "omegas": [-3, 3], "chis": [0, 0], "coupling": 1, "timesteps": 30}
    X = np.random.default_rng(0).uniform(-10, 10, (int(sys.argv[2]) if len(sys.argv) > 2 else 16384, 2))
else:
    c = {"max_N": 10, "sites": 3, "max_t": 25, "omegas": [-3, 3, 3], "chis": [0, 0, 0], "coupling": 1, "timesteps": 30}
    X = np.random.default_rng(0).uniform(-10, 10, (int(sys.argv[2]) if len(sys.argv) > 2 else 16384, 3))
p = nat.Problem.from_const(c)
x = torch.from_numpy(np.ascontiguousarray(X)).cuda()
for _ in range(2):
    out = p.loss_grad(x)
torch.cuda.synchronize()
print(name, len(X), "points, path", nat.PATH_NAMES[p.get_path()], "loss sum", float(out[0].sum()))
