import sys, time, numpy as np
sys.path.insert(0, "quantum-targeted-energy-transfer_b200")
import torch
from tet import _native
N, om, B = (int(sys.argv[1]), [-3, 3] if sys.argv[2] == "2" else [-3, 3, 3], int(sys.argv[3]))
const = {"max_N": N, "sites": len(om), "max_t": 25, "omegas": om, "chis": [0]*len(om), "coupling": 1, "timesteps": 30}
prob = _native.Problem.from_const(const)
f = len(om)
pts = np.random.default_rng(0).uniform(-10, 10, (B, f))
x = torch.from_numpy(pts).cuda()
h_chis = torch.from_numpy(pts).pin_memory(); h_loss = torch.empty(B, dtype=torch.float64).pin_memory()
h_grad = torch.empty(B, f, dtype=torch.float64).pin_memory(); h_ts = torch.empty(B, dtype=torch.int32).pin_memory()
def host():
    rc = _native.lib.qtet_loss_grad_host(prob._h, h_chis.data_ptr(), B, f - 1, h_loss.data_ptr(), h_grad.data_ptr(), h_ts.data_ptr())
    assert rc == 0
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter(); prob.loss_grad(x); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"device path {1e3*(t1-t0):.2f} ms")
for rep in range(4):
    prob.profile(True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); host(); t1 = time.perf_counter()
    k = prob.profile_read(); prob.profile(False)
    print(f"host path   {1e3*(t1-t0):.2f} ms  kernels {sum(v[0] for v in k.values()):.2f}", {n: round(v[0], 2) for n, v in k.items()})
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); prob.loss_grad(x); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"device path {1e3*(t1-t0):.2f} ms")
