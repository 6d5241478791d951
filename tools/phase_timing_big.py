#!/usr/bin/env python3
"""Per-phase cycle breakdown of the register-resident large-d kernels (debug build with -DQTET_PHASE_TIMING).

    QTET_NVCC_DEFS=-DQTET_PHASE_TIMING QTET_LIB_SUFFIX=_dbg python quantum-targeted-energy-transfer_b200/build_native.py
    QTET_B200_LIB=quantum-targeted-energy-transfer_b200/lib/libqtet_b200_dbg.so python tools/phase_timing_big.py 10 3 32768
"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantum-targeted-energy-transfer_b200"))
import torch  # noqa: E402
from tet import _native  # noqa: E402
N = int(sys.argv[1]); om = [-3, 3] if sys.argv[2] == "2" else [-3, 3, 3]; B = int(sys.argv[3])
const = {"max_N": N, "sites": len(om), "max_t": 25, "omegas": om, "chis": [0] * len(om), "coupling": 1, "timesteps": 30}
prob = _native.Problem.from_const(const)
x = torch.from_numpy(np.random.default_rng(0).uniform(-10, 10, (B, len(om)))).cuda()
prob.loss_grad(x); torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 16)()
_native.lib.qtet_debug_big_phase_cycles(out, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); prob.loss_grad(x); e1.record(); torch.cuda.synchronize()
_native.lib.qtet_debug_big_phase_cycles(out, 0)
names = ["hh 0 build H row", "hh 1 reduction", "hh 2 Q accumulation", "hh 3 store", "rp 4 load V + headers", "rp 5 replay",
         "rp 6 evolution+loss", "rp 7 psi*, V->smem, a", "rp 8 S pairs", "rp 9 G + grad"]
print(f"N={N} f={len(om)} d={prob.dim} B={B} step {e0.elapsed_time(e1):.3f} ms ; CTA-cycles per point:")
for i, n in enumerate(names):
    print(f"  {n:26s} {out[i]/B:10.1f}")
if hasattr(_native.lib, "qtet_debug_hf_phase_cycles") or True:
    try:
        hf = (ctypes.c_ulonglong * 8)()
        _native.lib.qtet_debug_hf_phase_cycles(hf, 0)
        for i, n in enumerate(["hf 0 build H tiles", "hf 1 reduction: update (+ loop)", "hf 2 Q^T accumulation", "hf 3 store", "hf 4 red: owner (thread 0 view)", "hf 5 red: barrier 1 wait", "hf 6 red: matvec", "hf 7 red: barrier 2 wait"]):
            print(f"  {n:26s} {hf[i]/(2*B):10.1f}   (tiled Householder kernel, both calls)")
    except AttributeError:
        pass
if any(out[i] for i in range(10, 15)):       # replay statistics (only in builds that count them)
    print(f"  per point: active rounds {out[10]/B:.1f}  own-window planes {out[11]/B:.0f}  batch-window planes {out[12]/B:.0f}  trip steps {out[13]/B:.0f}  trips {out[14]/B:.1f}")
