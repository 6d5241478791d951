import os, sys, time, numpy as np
sys.path.insert(0, "quantum-targeted-energy-transfer_b200")
import torch
from tet import _native
N, om = 10, [-3, 3, 3]
const = {"max_N": N, "sites": len(om), "max_t": 25, "omegas": om, "chis": [0]*len(om), "coupling": 1, "timesteps": 30}
prob = _native.Problem.from_const(const)
for B in (32768, 65536, 131072, 131072, 65536):
    x = torch.from_numpy(np.random.default_rng(0).uniform(-10, 10, (B, len(om)))).cuda()
    for rep in range(3):
        torch.cuda.synchronize()
        prob.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(); prob.loss_grad(x); e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        k = prob.profile_read(); prob.profile(False)
        ms = e0.elapsed_time(e1)
        ks = sum(v[0] for v in k.values())
        print(f"B={B} rep={rep} events {ms:.2f} ms kernels {ks:.2f} ms gap {ms-ks:.2f} enqueue {1e3*(t1-t0):.2f} ms", {n: round(v[0],2) for n, v in k.items()})
