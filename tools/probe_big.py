import os, sys, numpy as np
sys.path.insert(0, "quantum-targeted-energy-transfer_b200")
import torch
from tet import _native
for name, N, om, B in [("dimer N=100", 100, [-3,3], 1<<16), ("trimer N=10", 10, [-3,3,3], 1<<17)]:
    const = {"max_N": N, "sites": len(om), "max_t": 25, "omegas": om, "chis": [0]*len(om), "coupling": 1, "timesteps": 30}
    prob = _native.Problem.from_const(const)
    x = torch.from_numpy(np.random.default_rng(0).uniform(-10, 10, (B, len(om)))).cuda()
    prob.loss_grad(x); torch.cuda.synchronize()
    prob.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); prob.loss_grad(x); e1.record(); torch.cuda.synchronize()
    k = prob.profile_read(); prob.profile(False)
    ms = e0.elapsed_time(e1)
    print(name, "B", B, "ms", round(ms,2), "evals/s %.3e" % (B/ms*1e3), {n: (round(v[0],2), v[1]) for n, v in k.items()})
    prob.close()
