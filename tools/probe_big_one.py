import os, sys, numpy as np
sys.path.insert(0, "quantum-targeted-energy-transfer_b200")
import torch
from tet import _native
N, om, B = (int(sys.argv[1]), [-3,3] if sys.argv[2]=="2" else [-3,3,3], int(sys.argv[3]))
const = {"max_N": N, "sites": len(om), "max_t": 25, "omegas": om, "chis": [0]*len(om), "coupling": 1, "timesteps": 30}
prob = _native.Problem.from_const(const)
x = torch.from_numpy(np.random.default_rng(0).uniform(-10, 10, (B, len(om)))).cuda()
for _ in range(2): prob.loss_grad(x)
torch.cuda.synchronize()
print("done")
