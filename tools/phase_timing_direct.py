#!/usr/bin/env python3
"""Per-phase cycle breakdown of the direct kernel (debug build with -DQTET_PHASE_TIMING).

    QTET_NVCC_DEFS=-DQTET_PHASE_TIMING QTET_LIB_SUFFIX=_dbg python quantum-targeted-energy-transfer_b200/build_native.py
    QTET_B200_LIB=quantum-targeted-energy-transfer_b200/lib/libqtet_b200_dbg.so python tools/phase_timing_direct.py
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantum-targeted-energy-transfer_b200"))
import torch  # noqa: E402
from tet import _native  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
const = {"max_N": N, "sites": 2, "max_t": 25, "omegas": [-3, 3], "chis": [0, 0], "coupling": 1, "timesteps": 30}
ax = np.linspace(-10, 10, 1024)
xd, xa = np.meshgrid(ax, ax, indexing="ij")
pts = np.stack([xd.ravel(), xa.ravel()], axis=1)
B = len(pts)
prob = _native.Problem.from_const(const)
assert prob.get_path() == _native.QTET_PATH_DIRECT
x = torch.from_numpy(pts).cuda()
for _ in range(2):
    prob.loss_grad(x)
torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 16)()
_native.lib.qtet_debug_phase_cycles_direct(out, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); prob.loss_grad(x); e1.record(); torch.cuda.synchronize()
_native.lib.qtet_debug_phase_cycles_direct(out, 0)
names = ["0 recurrences pass 2 (join, scale)", "1 sort/readback + c/phase setup", "2 evolution+loss", "3 S pairs", "4 G = row S row", "5 grad reduce+store",
         "6 f=w^k powering", "7 psi(t*)", "8 a reduction", "9 records", "10 chi/ev loads + diagonal", "11 recurrence from the bottom", "12 recurrence from the top (join row)"]
tot = sum(out[i] for i in range(len(names)))
print(f"B={B} N={N} total step {e0.elapsed_time(e1):.3f} ms ; warp-cycles per point (4 points per warp):")
for i, n in enumerate(names):
    print(f"  {n:34s} {out[i]/B*4:10.1f} cycles/warp-task  {100*out[i]/tot:5.1f}%")
print(f"  {'sum':34s} {tot/B*4:10.1f}")

import ctypes as C
raw = (C.c_int64 * 4)()
_native.lib.qtet_stats(prob._h, raw, 0)
print("sum of ncol over warp-tasks:", raw[3], " (3 passes over the grid) -> mean ncol =", raw[3] / (3 * B / 4))
