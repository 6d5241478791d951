"""Small invocations of every kernel family for compute-sanitizer."""
import os, sys, numpy as np
sys.path.insert(0, "quantum-targeted-energy-transfer_b200")
import torch
from tet import _native
rng = np.random.default_rng(0)
for N, om, B in [(20, [-3, 3], 77), (4, [-3, 3], 33), (31, [-3, 3], 40), (40, [-3, 3], 9), (3, [-3, 3, 3], 21), (5, [-3, 3, 3], 13)]:
    const = {"max_N": N, "sites": len(om), "max_t": 25, "omegas": om, "chis": [0]*len(om), "coupling": 1, "timesteps": 30}
    prob = _native.Problem.from_const(const)
    x = rng.uniform(-10, 10, (B, len(om)))
    for path in (None, _native.QTET_PATH_GENERIC):
        if path is not None: prob.set_path(path)
        l, g, t = prob.loss_grad(x); tr = prob.trace(x)
        torch.cuda.synchronize()
    out = prob.adam_run(x[:8], iterations=5, history=True)
    mn, ix = prob.argmin(out["best_loss"])
    torch.cuda.synchronize()
    prob.close()
print("sanitize probe ok")
