"""CPU tests: the numpy oracle against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py), against independent derivations of the gradient, and against the
known answers / conservation laws of SURVEY.md §4."""
import numpy as np
import pytest

from oracle import qtet_oracle as O
from conftest import make_const


def test_basis_matches_reference_derive(golden):
    for b in golden["basis"]:
        s = O.fock_basis(b["max_N"], b["sites"])
        assert s.shape == (b["dim"], b["sites"])
        assert s.astype(int).tolist() == b["states"]


def test_hamiltonian_bit_exact(golden):
    for h in golden["hamiltonian"]:
        H = O.hamiltonian(h["const"], h["chis"])
        Href = np.array([float.fromhex(x) for x in h["H_hex"]]).reshape(H.shape)
        assert np.array_equal(H, Href)
        assert np.array_equal(H, H.T)


def test_loss_and_trace_vs_reference_formulas_c128(golden):
    # reference control flow evaluated in complex128: pins the formulas; tolerance 1e-12 absolute
    for r in golden["loss"]:
        l = O.loss(r["const"], r["chis"], r["site"])
        tr = O.trace(r["const"], r["chis"], r["site"])
        assert abs(l - r["loss_c128"]) < 1e-12
        assert np.abs(tr - np.array(r["trace_c128"])).max() < 1e-12


def test_loss_and_trace_vs_reference_precision_c64(golden):
    # the reference's real arithmetic (complex64 evolution, HamiltonianLoss.py:205-212):
    # FP64 results sit within the float32 noise of it (measured <= 3e-5 on these cases)
    for r in golden["loss"]:
        l = O.loss(r["const"], r["chis"], r["site"])
        tr = O.trace(r["const"], r["chis"], r["site"])
        assert abs(l - r["loss_c64"]) < 1e-4
        assert np.abs(tr - np.array(r["trace_c64"])).max() < 1e-4


def test_train_loop_vs_reference_optimizer(golden):
    # Optimizer.train run unmodified (FD gradient, c128): same number of epochs (stop rules,
    # revert rule, break rule), same losses / best_vars up to FD-noise amplified by Adam.
    for t in golden["train"]:
        res = O.adam_train(t["const"], t["init"], None, t["train_sites"], t["iterations"], t["lr"])
        assert len(res["loss"]) == len(t["loss"]), t["init"]
        assert np.abs(np.array(res["loss"]) - np.array(t["loss"])).max() < 1e-3
        assert np.abs(np.array(res["best_vars"]) - np.array(t["best_vars"])).max() < 1e-4
        assert [len(v) for v in res["var_data"]] == [len(v) for v in t["var_data"]]
        for a, b in zip(res["var_data"], t["var_data"]):
            if a:
                assert np.abs(np.array(a) - np.array(b)).max() < 1e-3


@pytest.mark.parametrize("N,omegas,chis", [
    (4, (-3, 3), (1.1, -0.7)), (20, (-3, 3), (0.31, -0.28)), (20, (-3, 3), (-6.5, 4.25)),
    (5, (-3, 3, 3), (0.6, 0.1, -0.6)), (10, (-3, 3, 3), (1.5, -0.5, 0.75)),
])
def test_gradient_vs_central_differences(N, omegas, chis):
    c = make_const(N, omegas)
    for site in (len(omegas) - 1, 0):
        L, g, k = O.loss_grad(c, chis, site)
        assert abs(L - O.loss(c, chis, site)) < 1e-13
        for j in range(len(chis)):
            h = 1e-6
            xp = list(chis); xm = list(chis)
            xp[j] += h; xm[j] -= h
            fd = (O.loss(c, xp, site) - O.loss(c, xm, site)) / (2 * h)
            assert abs(fd - g[j]) < 1e-6 * max(1.0, abs(g[j]))


def test_gradient_vs_torch_autograd_through_eigh():
    torch = pytest.importorskip("torch")
    for (N, omegas, chis) in [(4, (-3, 3), (1.1, -0.7)), (20, (-3, 3), (2.3, -1.9)),
                              (10, (-3, 3, 3), (0.6, 0.1, -0.6))]:
        c = make_const(N, omegas)
        f = len(omegas)
        states = torch.tensor(O.fock_basis(N, f))
        off = torch.tensor(O.offdiag_matrix(states.numpy(), 1.0))
        x = torch.tensor(chis, dtype=torch.float64, requires_grad=True)
        diag = states @ torch.tensor(omegas, dtype=torch.float64) + (0.5 * states ** 2) @ x
        e, V = torch.linalg.eigh(off + torch.diag(diag))
        t = torch.tensor(O.time_grid(c))
        z = V[0, :][None, :] * torch.exp(-1j * e[None, :] * t[:, None])
        psi = z @ V.T.to(torch.complex128)
        n = (psi.abs() ** 2 * states[:, f - 1]).sum(-1)
        loss = N - n.max()
        loss.backward()
        L, g, _ = O.loss_grad(c, chis, f - 1)
        assert abs(L - loss.item()) < 1e-12
        assert np.abs(g - x.grad.numpy()).max() < 1e-9 * max(1.0, np.abs(g).max())


def test_batch_matches_single():
    c = make_const(6)
    rng = np.random.default_rng(0)
    X = rng.uniform(-10, 10, (17, 2))
    for site in (1, 0):
        Lb, gb, kb = O.loss_grad_batch(c, X, site)
        for i in range(len(X)):
            L, g, k = O.loss_grad(c, X[i], site)
            assert abs(L - Lb[i]) < 1e-12 and np.abs(g - gb[i]).max() < 1e-10 and k == kb[i]


def test_known_answers():
    # SURVEY.md §4 [probe]: structure
    c = make_const(3)
    H = O.hamiltonian(c, [1.5, -2.0])
    assert np.allclose(np.diag(H), [-2.25, -1.0, -0.25, 0.0])
    assert np.allclose(np.diag(H, 1), [-np.sqrt(3), -2, -np.sqrt(3)])
    assert O.fock_basis(2, 3).astype(int).tolist() == [[2, 0, 0], [1, 1, 0], [1, 0, 1], [0, 2, 0], [0, 1, 1], [0, 0, 2]]
    # physics: default constants -> 2.70049; TET resonance chi_D = -chi_A = 6/N
    assert abs(O.loss(make_const(3), [0, 0]) - 2.7004884) < 2e-6
    assert abs(O.loss(make_const(3), [2, -2]) - 0.004190) < 1e-6
    assert abs(O.loss(make_const(4), [1.5, -1.5]) - 0.005586) < 1e-6
    assert abs(O.loss(make_const(20), [0.3, -0.3]) - 0.027931) < 5e-6   # probe value was c64 arithmetic


def test_conservation_and_initial_condition():
    rng = np.random.default_rng(3)
    for (N, omegas) in [(7, (-3, 3)), (4, (-3, 3, 3))]:
        c = make_const(N, omegas)
        x = rng.uniform(-5, 5, len(omegas))
        tot = sum(O.trace(c, x, s) for s in range(len(omegas)))
        assert np.abs(tot - N).max() < 1e-11
        assert abs(O.trace(c, x, 0)[0] - N) < 1e-12
        assert abs(O.trace(c, x, len(omegas) - 1)[0]) < 1e-12


def test_truth_mpmath_small():
    """50-digit evaluation of the same formulas: the FP64 oracle's own error (dimer N=4)."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    c = make_const(4)
    chis = [1.37, -0.91]
    H = mp.matrix(O.hamiltonian(c, chis).tolist())
    E, Q = mp.eigsy(H)
    n_site = O.fock_basis(4, 2)[:, 1]
    tr = []
    for t in O.time_grid(c):
        tt = mp.mpf(float(t))
        psi = [sum(Q[j, i] * Q[0, i] * mp.e ** (-1j * E[i] * tt) for i in range(5)) for j in range(5)]
        tr.append(float(sum(n_site[j] * abs(psi[j]) ** 2 for j in range(5))))
    assert np.abs(np.array(tr) - O.trace(c, chis, 1)).max() < 1e-12


def test_site_argument_forms():
    c = make_const(3)
    assert O.loss(c, [1.0, -1.0], "x1") == O.loss(c, [1.0, -1.0], 1)
    with pytest.raises(ValueError):
        O.loss(c, [1.0, -1.0], 1.0)


def test_get_combinations_grid_and_bins():
    lims = {"x0lims": [0.5, 2.5], "x1lims": [-2.5, -0.5]}
    g = O.get_combinations(lims, [0, 1], "grid", 2)
    assert g == [(0.5, -2.5), (0.5, -0.5), (2.5, -2.5), (2.5, -0.5)]
    b = O.get_combinations(lims, [0, 1], "bins", 2, rng=np.random.default_rng(0))
    assert sorted(map(tuple, b)) == sorted(g)
    with pytest.raises(ValueError):
        O.get_combinations(lims, [0, 1], "nope", 2)
