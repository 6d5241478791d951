#!/usr/bin/env python3
"""Debug: the tiled Householder stage alone, staged against single-kernel (needs a -DQTET_DEBUG_EXPORTS build)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantum-targeted-energy-transfer_b200"))
import torch
from tet import _native as nat
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
c = {"max_N": N, "sites": 3, "max_t": 25, "omegas": [-3, 3, 3], "chis": [0, 0, 0], "coupling": 1, "timesteps": 30}
p = nat.Problem.from_const(c)
d = p.dim
n = 5
X = torch.from_numpy(np.random.default_rng(0).uniform(-10, 10, (n, 3))).cuda()
f = nat.lib.qtet_debug_hh
f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong] + [ctypes.c_void_p] * 4
f.restype = ctypes.c_int
out = {}
for staged in (0, 1):
    td = torch.zeros(n, d, dtype=torch.float64, device="cuda"); te = torch.zeros_like(td)
    q = torch.zeros(n, d, d, dtype=torch.float64, device="cuda")
    rfl = torch.zeros(n, 64 * 110, dtype=torch.float64, device="cuda")
    rc = f(p._h, X.data_ptr(), n, td.data_ptr(), te.data_ptr(), q.data_ptr(), rfl.data_ptr() if staged else None)
    assert rc == 0, nat.lib.qtet_last_error()
    out[staged] = (td.cpu().numpy(), te.cpu().numpy(), q.cpu().numpy())
for name, i in (("tri_d", 0), ("tri_e", 1), ("Q", 2)):
    a, b = out[0][i], out[1][i]
    print(name, "max |staged - single| =", np.abs(a - b).max())
    if i < 2:
        bad = np.argwhere(np.abs(a - b) > 1e-9)
        print("   first mismatches (point, index):", bad[:8].tolist())
    else:
        bad = np.argwhere(np.abs(a[0] - b[0]) > 1e-9)
        print("   point 0: mismatching (row of X = col of Q, col) count", len(bad), "first", bad[:6].tolist(), "rows", sorted(set(bad[:, 0].tolist()))[:12], "cols", sorted(set(bad[:, 1].tolist()))[:12])
Q = out[1][2][0]
print("staged: |Q^T Q - I| =", np.abs(Q @ Q.T - np.eye(d)).max(), " single:", np.abs(out[0][2][0] @ out[0][2][0].T - np.eye(d)).max())
