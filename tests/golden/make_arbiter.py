#!/usr/bin/env python3
"""High-precision arbiter for the large Fock spaces (dimer N=100, trimer N=10) and a few dimer N=20 points.

At |e t| ~ 1e6 the FP64 oracle (numpy eigh + complex128 phases) carries a phase error of eps*||H||*t ~ 1e-10...1e-9 itself,
so a GPU-vs-oracle difference of that size does not say which side is off.  This script evaluates the SAME formulas
(HamiltonianLoss.py:124-233: H, eigh, c = V[0,:], psi(t), <n(t)> on linspace(0, max_t, T), N - max) with mpmath at 40
digits, and the gradient as a central difference (h = 1e-12, 40 digits) of the trace sample t* held fixed, which is what
TF's reduce_max sub-gradient differentiates (Optimizer.py:85-91).  Output: tests/golden/arbiter_mp.json.

    python tests/golden/make_arbiter.py        (~15 min on 8 cores; mpmath only, no reference import needed)
"""
import json
import os
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import qtet_oracle as O  # noqa: E402  (only for the integer basis / hopping structure, cross-checked bit-exactly
                                      #               against the reference in reference_golden.json)

DPS = 40


def const(N, om):
    return {"max_N": N, "sites": len(om), "max_t": 25, "omegas": list(om), "chis": [0] * len(om), "coupling": 1, "timesteps": 30}


def mp_trace(c, chi, site, only=None):
    import mpmath as mp
    mp.mp.dps = DPS
    states = O.fock_basis(c["max_N"], c["sites"])
    off = O.offdiag_matrix(states, c["coupling"])          # entries -lam*sqrt(integer): recomputed exactly below
    d = len(states)
    f = c["sites"]
    M = mp.matrix(d, d)
    for i in range(d):
        acc = mp.mpf(0)
        for k in range(f):
            n = mp.mpf(int(states[i, k]))
            acc += mp.mpf(c["omegas"][k]) * n + chi[k] * n * n / 2
        M[i, i] = acc
        for j in range(d):
            if i != j and off[i, j] != 0.0:
                M[i, j] = -mp.mpf(c["coupling"]) * mp.sqrt(mp.mpf(int(round(off[i, j] ** 2 / c["coupling"] ** 2))))
    E, Q = mp.eigsy(M)
    ts = [mp.mpf(25) * k / 29 for k in range(30)]
    out = []
    for k, t in enumerate(ts):
        if only is not None and k != only:
            out.append(None)
            continue
        ph = [Q[0, i] * mp.expj(-E[i] * t) for i in range(d)]
        tot = mp.mpf(0)
        for j in range(d):
            nj = int(states[j, site])
            if nj == 0:
                continue
            psi = mp.fsum(Q[j, i] * ph[i] for i in range(d))
            tot += nj * (psi.real ** 2 + psi.imag ** 2)
        out.append(tot)
    return out


def job(args):
    import mpmath as mp
    mp.mp.dps = DPS
    name, N, om, chi = args
    c = const(N, om)
    f = len(om)
    site = f - 1
    x = [mp.mpf(float(v)) for v in chi]
    tr = mp_trace(c, x, site)
    kstar = max(range(30), key=lambda k: (tr[k], -k))            # first maximum
    loss = mp.mpf(N) - tr[kstar]
    h = mp.mpf(10) ** -12
    grad = []
    for k in range(f):
        xp = list(x); xm = list(x)
        xp[k] += h; xm[k] -= h
        gp = mp_trace(c, xp, site, only=kstar)[kstar]
        gm = mp_trace(c, xm, site, only=kstar)[kstar]
        grad.append(-(gp - gm) / (2 * h))
    srt = sorted(tr, reverse=True)
    return {"system": name, "max_N": N, "omegas": list(om), "chi": [float(v) for v in chi], "site": site,
            "trace": [float(v) for v in tr], "loss": float(loss), "tstar": int(kstar), "grad": [float(g) for g in grad],
            "runner_up_gap": float(srt[0] - srt[1])}


def main():
    jobs = []
    X = np.random.default_rng(0).uniform(-10, 10, (48, 2))            # the points of test_dimer_large_fock_space
    for x in list(X[:4]) + [np.array([0.06, -0.06]), np.array([10.0, -10.0])]:
        jobs.append(("dimer100", 100, (-3, 3), x))
    X = np.random.default_rng(10).uniform(-10, 10, (96, 3))           # the points of test_chain_loss_grad_vs_oracle (N=10)
    for x in list(X[:4]) + [np.array([0.6, 0.0, -0.6]), np.array([10.0, 10.0, 10.0])]:
        jobs.append(("trimer10", 10, (-3, 3, 3), x))
    X = np.random.default_rng(120).uniform(-10, 10, (257, 2))         # test_dimer_loss_grad_vs_oracle (N=20)
    for x in list(X[:4]) + [np.array([0.3, -0.3]), np.array([2.0, 2.0]), np.array([1.0, 1.0]), np.array([0.6, 0.0])]:
        jobs.append(("dimer20", 20, (-3, 3), x))
    with Pool(min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(job, jobs, chunksize=1)
    out = {"dps": DPS, "fd_step": 1e-12, "points": res}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "arbiter_mp.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", len(res), "points")


if __name__ == "__main__":
    main()
