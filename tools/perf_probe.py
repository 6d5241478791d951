#!/usr/bin/env python3
"""Throughput of the hot path on the non-headline configs (dimer N=4, N=100, trimer N=10 ...)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantum-targeted-energy-transfer_b200"))
import torch
from tet import _native
cases = [("dimer N=4", 4, [-3, 3], 1 << 20), ("dimer N=10", 10, [-3, 3], 1 << 20), ("dimer N=31", 31, [-3, 3], 1 << 18),
         ("dimer N=100", 100, [-3, 3], 1 << 16), ("trimer N=4", 4, [-3, 3, 3], 1 << 18), ("trimer N=10", 10, [-3, 3, 3], 1 << 17),
         ("4-site chain N=3", 3, [-3, 0, 1, 3], 1 << 18)]
for name, N, om, B in cases:
    const = {"max_N": N, "sites": len(om), "max_t": 25, "omegas": om, "chis": [0] * len(om), "coupling": 1, "timesteps": 30}
    prob = _native.Problem.from_const(const)
    x = torch.from_numpy(np.random.default_rng(0).uniform(-10, 10, (B, len(om)))).cuda()
    prob.loss_grad(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): prob.loss_grad(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    d = prob.dim
    F = (6.0 if len(om) == 2 else 9.0) * d ** 3 + 4 * 30 * d ** 2 + 4 * d ** 3
    print(f"{name:12s} d={d:4d} path={prob.get_path()} B={B:8d} {ms:9.3f} ms  {B/ms*1e3:12.3e} evals/s  {B/ms*1e3*F/1e12:7.3f} TFLOP/s(alg)")
    prob.close()
