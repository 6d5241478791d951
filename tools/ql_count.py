#!/usr/bin/env python3
"""CPU emulation of the two-pass perfect-shift QL of ql_scalar*_kernel: counts sweeps and plane rotations per point
(how much work the rotation replay has to do).   python tools/ql_count.py 10 3 40"""
import sys, os
import numpy as np
import scipy.linalg as sla
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import qtet_oracle as O

EPS = 2.0 ** -53


def ql(dg, of, lam=None):
    """returns (#sweeps, #rotations, windows); lam = eigenvalues by index for perfect shifts (pass 2) or None."""
    d = len(dg); dg = dg.copy(); of = np.append(of.copy(), 0.0)
    sweeps = rots = 0; wins = []
    tried = -1; it = 0; lprev = -1
    l = 0
    while l < d - 1:
        # first negligible off-diagonal at or above l
        m = l
        while m < d - 1 and abs(of[m]) > EPS * (abs(dg[m]) + abs(dg[m + 1])):
            m += 1
        if m == l:
            l += 1; continue
        if l != lprev: it = 0; lprev = l
        it += 1
        if it > 60: break
        if lam is not None and tried != l:
            tried = l; g = dg[m] - lam[l]
        else:
            gg = (dg[l + 1] - dg[l]) / (2.0 * of[l]); rr = np.hypot(gg, 1.0)
            g = dg[m] - dg[l] + of[l] / (gg + np.copysign(rr, gg))
        s = c = 1.0; p = 0.0
        sweeps += 1; wins.append((l, m))
        broke = False
        for i in range(m - 1, l - 1, -1):
            f = s * of[i]; b = c * of[i]
            r = np.hypot(f, g)
            of[i + 1] = r
            if r == 0.0:
                dg[i + 1] -= p; of[m] = 0.0; broke = True; break
            s = f / r; c = g / r
            g = dg[i + 1] - p
            r = (dg[i] - g) * s + 2.0 * c * b
            p = s * r
            dg[i + 1] = g + p
            g = c * r - b
            rots += 1
        if broke: continue
        dg[l] -= p; of[l] = g; of[m] = 0.0
    return sweeps, rots, wins, dg


N = int(sys.argv[1]); f = int(sys.argv[2]); npts = int(sys.argv[3])
om = [-3, 3] if f == 2 else [-3, 3, 3]
const = {"max_N": N, "sites": f, "max_t": 25, "omegas": om, "chis": [0] * f, "coupling": 1, "timesteps": 30}
rng = np.random.default_rng(0)
X = rng.uniform(-10, 10, (npts, f))
tot = []
for x in X:
    H = O.hamiltonian(const, x)
    d = H.shape[0]
    if f > 2:
        T = sla.hessenberg(H)
        dg, of = np.diag(T).copy(), np.diag(T, -1).copy()
    else:
        dg, of = np.diag(H).copy(), np.diag(H, -1).copy()
    dg -= dg.mean()
    s1, r1, _, lam = ql(dg, of)
    s2, r2, wins, _ = ql(dg, of, lam)
    tot.append((s1, r1, s2, r2))
tot = np.array(tot)
print(f"d={d}: pass1 sweeps {tot[:,0].mean():.1f} rots {tot[:,1].mean():.0f} | pass2 sweeps {tot[:,2].mean():.1f} (max {tot[:,2].max()}) rots {tot[:,3].mean():.0f} (max {tot[:,3].max()})  d^2/2={d*d/2:.0f}")
# how tight are the windows against the top of the matrix?  (identity planes a top-anchored sweep would execute)
waste = used = 0
for x in X[:min(len(X), 20)]:
    H = O.hamiltonian(const, x)
    if f > 2:
        T = sla.hessenberg(H); dg, of = np.diag(T).copy(), np.diag(T, -1).copy()
    else:
        dg, of = np.diag(H).copy(), np.diag(H, -1).copy()
    dg -= dg.mean()
    _, _, _, lam = ql(dg, of)
    _, _, wins, _ = ql(dg, of, lam)
    for (l, m) in wins:
        used += m - l; waste += (len(dg) - 1 - m)
print(f"planes in windows {used}, identity planes between window top and d-1: {waste} ({100*waste/used:.1f}%)")
