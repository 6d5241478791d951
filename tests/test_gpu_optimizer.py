"""GPU tests of the batched optimiser (Optimizer.train semantics), the arg-min and the tet-compatible
Python API, against the oracle and the golden trajectories of the unmodified reference."""
import json
import os

import numpy as np
import pytest

from conftest import make_const

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def native():
    from tet import _native
    return _native


@pytest.fixture(scope="module")
def O():
    from oracle import qtet_oracle
    return qtet_oracle


def test_adam_golden_trajectories(native, golden):
    for t in golden["train"]:
        prob = native.Problem.from_const(t["const"])
        out = prob.adam_run(np.array([t["init"]]), train_sites=t["train_sites"], iterations=t["iterations"],
                            lr=t["lr"], history=True)
        n = int(out["epochs"][0])
        assert n == len(t["loss"]), (t["init"], n, len(t["loss"]))
        lh = out["loss_hist"][:n, 0].cpu().numpy()
        assert np.abs(lh - np.array(t["loss"])).max() < 1e-3          # FD-gradient noise of the golden run
        assert np.abs(out["best_vars"][0].cpu().numpy() - np.array(t["best_vars"])).max() < 1e-4
        assert abs(float(out["best_loss"][0]) - min(t["loss"])) < 1e-4     # same FD-noise amplification
        vh = out["var_hist"][:n // 10, 0].cpu().numpy()
        for k, row in enumerate(t["var_data"]):
            assert len(row) == n // 10
            if row:
                assert np.abs(vh[:, k] - np.array(row)).max() < 1e-3
        prob.close()


def test_adam_batch_vs_oracle(native, O):
    """The device optimiser against the oracle's restatement of Optimizer.train.  Adam in this rugged
    landscape amplifies 1e-13 loss/grad differences chaotically over tens of epochs, so the
    control flow is checked with BOTH loops fed by the same loss/grad source (the CUDA kernel), and the
    pure-oracle loop is compared on the first epochs only."""
    c = make_const(4)
    rng = np.random.default_rng(4)
    X0 = rng.uniform(-3, 3, (40, 2))
    prob = native.Problem.from_const(c)
    out = prob.adam_run(X0, train_sites=(0, 1), iterations=60, lr=0.1, history=True)
    ep = out["epochs"].cpu().numpy()
    lh = out["loss_hist"].cpu().numpy()
    vh = out["var_hist"].cpu().numpy()

    def gpu_lg(x):
        L, g, _ = prob.loss_grad_host(np.asarray(x)[None, :], 1)
        return L[0], g[0]

    stopped_early = 0
    for g in range(len(X0)):
        r = O.adam_train(c, X0[g], None, (0, 1), 60, 0.1, loss_grad_fn=gpu_lg)
        assert r["epochs"] == ep[g], (g, r["epochs"], ep[g])
        assert np.abs(np.array(r["loss"]) - lh[:ep[g], g]).max() < 1e-7
        assert np.abs(np.array(r["best_vars"]) - out["best_vars"][g].cpu().numpy()).max() < 1e-7
        assert np.abs(r["final_vars"] - out["final_vars"][g].cpu().numpy()).max() < 1e-7
        assert abs(min(r["loss"]) - float(out["best_loss"][g])) < 1e-7
        for k in range(2):
            assert np.abs(np.array(r["var_data"][k]) - vh[:ep[g] // 10, g, k]).max(initial=0) < 1e-7
        assert np.isnan(lh[ep[g]:, g]).all()
        stopped_early += ep[g] < 60
        pure = O.adam_train(c, X0[g], None, (0, 1), 60, 0.1)
        m = min(5, pure["epochs"], ep[g])
        assert np.abs(np.array(pure["loss"][:m]) - lh[:m, g]).max() < 1e-9
    assert stopped_early >= 3            # ragged convergence is exercised
    prob.close()


def test_adam_partial_training_and_nonprogress(native, O):
    c = make_const(3)
    prob = native.Problem.from_const(c)
    # only chi_0 trained: chi_1 never moves
    out = prob.adam_run(np.array([[1.0, -1.5]]), train_sites=(0,), iterations=30, lr=0.1)
    assert float(out["final_vars"][0, 1]) == -1.5
    # lr so small that every move is < tol: stops after the third epoch (Optimizer.py:202-206)
    out = prob.adam_run(np.array([[4.0, 1.0]]), train_sites=(0, 1), iterations=50, lr=1e-9)
    assert int(out["epochs"][0]) == 3
    prob.close()


def test_argmin(native):
    import torch
    prob = native.Problem.from_const(make_const(2))
    rng = np.random.default_rng(0)
    for n in (1, 5, 1000, 100003):
        v = rng.uniform(0, 1, n)
        v[rng.integers(0, n)] = -1.0
        j = int(np.argmin(v))
        mn, ix = prob.argmin(torch.from_numpy(v).cuda())
        assert int(ix) == j and float(mn) == v[j]
    v = np.array([3.0, 1.0, 1.0, 2.0])
    mn, ix = prob.argmin(torch.from_numpy(v).cuda())
    assert int(ix) == 1                                   # first minimum wins, like np.argmin
    prob.close()


def test_tet_api_loss_and_optimizer(native, O, tmp_path):
    import tet
    from tet.Optimizer import Adam, mp_opt
    c = make_const(4)
    L = tet.Loss(c)
    val = L([1.5, -1.5], site=1)
    assert abs(float(val.numpy()) - O.loss(c, [1.5, -1.5], 1)) < 1e-10
    tr = L([1.5, -1.5], site="x0", single_value=False).numpy()
    assert tr.shape == (30,) and np.abs(tr - O.trace(c, [1.5, -1.5], 0)).max() < 1e-10
    with pytest.raises(ValueError):
        L([1.5, -1.5], site=1.5)
    # the reference's remaining public methods (HamiltonianLoss.py:62-176, 218-233)
    L([1.5, -1.5], site=1)
    assert abs(float(L.loss().numpy()) - O.loss(c, [1.5, -1.5], 1)) < 1e-10
    H = O.hamiltonian(c, [1.5, -1.5])
    assert L.ConstructElement(2, 2) == H[2, 2] and L.ConstructElement(2, 3) == H[2, 3]
    L.derive(); L.getHashArray()
    assert (L.states == O.fock_basis(4, 2)).all() and L.sorted_indices[0] == 0
    opt = tet.Optimizer(const=c, target_site=1, Print=False, iterations=500, train_sites=[0, 1],
                        opt=Adam(learning_rate=0.1, beta_1=0.9, amsgrad=False), data_path=str(tmp_path / "g0"))
    res = opt(0.5, -0.5, write_data=True)
    ref = O.adam_train(c, [0.5, -0.5], 1, [0, 1], 500, 0.1)
    assert len(res["loss"]) == ref["epochs"]
    assert np.abs(np.array(res["best_vars"]) - np.array(ref["best_vars"])).max() < 1e-6
    assert min(res["loss"]) < 0.1
    for name in ("losses.txt", "init_chis.txt", "x0trajectory.txt", "x1trajectory.txt"):
        assert (tmp_path / "g0" / name).exists()
    assert len((tmp_path / "g0" / "losses.txt").read_text().split()) == len(res["loss"])
    grads, loss = opt.get_grads()
    assert len(grads) == 2 and grads[0] is not None
    row = mp_opt(0, (2.5, -2.5), str(tmp_path), c, 1, 500, 0.1, 0.9, False, False, [0, 1])
    assert row.shape == (3,) and row[2] < 0.1


def test_solver_mp_config1(native, O, tmp_path):
    # BASELINE config 1: dimer N=4, grid method, limits [[0.5,2.5],[-2.5,-0.5]] -> 4 guesses, all reach TET
    from tet.solver_mp import solver_mp
    c = make_const(4)
    lims = {"x0lims": [0.5, 2.5], "x1lims": [-2.5, -0.5]}
    res = solver_mp(lims, dict(c), grid=2, lr=0.1, beta_1=0.9, amsgrad=False, write_data=True, method="grid",
                    target_site=1, data_path=str(tmp_path / "data"))
    _, opt_vars, min_loss = O.solve(c, O.get_combinations(lims, [0, 1], "grid", 2), 1, [0, 1], 500, 0.1)
    assert res["min_n"] <= 0.1
    assert abs(res["min_n"] - min_loss) < 1e-6
    assert np.abs(np.array(res["chis"]) - np.array(opt_vars)).max() < 1e-5
    assert (tmp_path / "data" / "iteration_0" / "data_optimizer_3" / "losses.txt").exists()


def test_qtet_cli(tmp_path):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = {"constants": {"max_N": 4, "sites": 2, "max_t": 25, "omegas": [-3, 3], "chis": [0, 0], "coupling": 1, "timesteps": 30},
           "tensorflow_params": {"lr": 0.1, "beta_1": 0.9, "amsgrad": False, "iterations": 1000, "tol": 1e-8, "train_sites": [0, 1]},
           "solver_params": {"target": 1, "Npoints": 2, "epochs_grid": 500, "epochs_bins": 1000},
           "limits": [[0.5, 2.5], [-2.5, -0.5]]}
    (tmp_path / "constants.json").write_text(json.dumps(cfg))
    out = subprocess.run([sys.executable, os.path.join(root, "quantum-targeted-energy-transfer_b200", "qtet.py"),
                          "-p", str(tmp_path / "out"), "-c", str(tmp_path / "constants.json"), "-n", "4", "-m", "grid"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    runs = list((tmp_path / "out").iterdir())
    assert len(runs) == 1
    params = json.loads((runs[0] / "parameters.json").read_text())
    assert params["constants"]["min_n"] <= 0.1
    assert set(params) == {"constants", "tensorflow_params", "solver_params"}
