"""Headline system on RANDOM points (no locality between the 32 points of a batch / 4 points of a warp)."""
import sys, numpy as np
sys.path.insert(0, "quantum-targeted-energy-transfer_b200")
import torch
from tet import _native
const = {"max_N": 20, "sites": 2, "max_t": 25, "omegas": [-3, 3], "chis": [0, 0], "coupling": 1, "timesteps": 30}
prob = _native.Problem.from_const(const)
B = 1 << 20
for name, pts in [("grid", None), ("random", np.random.default_rng(0).uniform(-10, 10, (B, 2)))]:
    if pts is None:
        ax = np.linspace(-10, 10, 1024); xd, xa = np.meshgrid(ax, ax, indexing="ij"); pts = np.stack([xd.ravel(), xa.ravel()], 1)
    x = torch.from_numpy(np.ascontiguousarray(pts)).cuda()
    for _ in range(2): prob.loss_grad(x)
    torch.cuda.synchronize()
    prob.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): prob.loss_grad(x)
    e1.record(); torch.cuda.synchronize()
    k = prob.profile_read(); prob.profile(False)
    ms = e0.elapsed_time(e1) / 3
    print(f"{name:7s} {ms:8.3f} ms  {B/ms*1e3:.3e} evals/s  apply {k['apply'][0]/k['apply'][1]:.3f} ms/launch  ql {k['ql_scalar'][0]/k['ql_scalar'][1]:.3f} ms/launch")
