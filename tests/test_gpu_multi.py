"""Multi-GPU (one process per GPU, NCCL) test of the sharded solver: the guesses are split contiguously over
the ranks, every rank optimises its shard on its own GPU, and ONE all_gather picks the global arg-min
(replaces multiprocessing.Pool + np.argmin, /root/reference/src/tet/solver_mp.py:157-184).
Skipped on boxes with fewer than two GPUs (the host-side logic is covered on CPU with gloo in test_host_cpu.py)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "quantum-targeted-energy-transfer_b200")

_WORKER = r"""
import json, os, sys
import numpy as np
sys.path.insert(0, {pkg!r})
import torch, torch.distributed as dist
local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from tet.solver_mp import solver_mp, randomCombinations
const = {{"max_N": 4, "sites": 2, "max_t": 25, "omegas": [-3, 3], "chis": [0, 0], "coupling": 1, "timesteps": 30}}
lims = {{"x0lims": [-3, 3], "x1lims": [-3, 3]}}
combos = randomCombinations(lims, [0, 1], 257, seed=3)
res = solver_mp(lims, dict(const), lr=0.1, method='grid', epochs_grid=120, target_site=1,
                data_path={out!r} + f'/rank{{dist.get_rank()}}', combinations=combos)
if dist.get_rank() == 0:
    json.dump({{"chis": res["chis"], "min_n": res["min_n"]}}, open({out!r} + '/result.json', 'w'))
dist.barrier()
dist.destroy_process_group()
"""


def test_sharded_solver_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(pkg=PKG, out=str(tmp_path)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    got = json.loads((tmp_path / "result.json").read_text())
    # single-process answer on GPU 0
    sys.path.insert(0, PKG)
    from tet.solver_mp import solver_mp, randomCombinations
    const = {"max_N": 4, "sites": 2, "max_t": 25, "omegas": [-3, 3], "chis": [0, 0], "coupling": 1, "timesteps": 30}
    lims = {"x0lims": [-3, 3], "x1lims": [-3, 3]}
    combos = randomCombinations(lims, [0, 1], 257, seed=3)
    ref = solver_mp(lims, dict(const), lr=0.1, method="grid", epochs_grid=120, target_site=1,
                    data_path=str(tmp_path / "single"), combinations=combos)
    assert got["min_n"] == ref["min_n"]
    assert np.array_equal(np.array(got["chis"]), np.array(ref["chis"]))
    assert got["min_n"] < 0.1


def test_qtet_cli_two_ranks_one_output_tree(tmp_path):
    """`qtet -m bins` under torchrun with two ranks (gloo, both on cuda:0 so that a 1-GPU box can run it): ONE timestamped
    output directory, every guess directory exactly once, one consistent draw of the random 'bins' guesses
    (the reference draws them from the unseeded global RNG, solver_mp.py:70; qtet.py:26 timestamps per process)."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    consts = {"constants": {"max_N": 3, "max_t": 25, "omegas": [-3, 3], "chis": [0, 0], "sites": 2, "coupling": 1, "timesteps": 30},
              "tensorflow_params": {"lr": 0.1, "iterations": 200, "tol": 1e-8, "beta_1": 0.9, "amsgrad": False, "train_sites": [0, 1]},
              "solver_params": {"target": 1, "Npoints": 3, "epochs_grid": 60, "epochs_bins": 60},
              "limits": [[0.5, 3.5], [-3.5, -0.5]]}
    sys.path.insert(0, PKG)
    import tet.constants as tc
    consts["tensorflow_params"] = {k: consts["tensorflow_params"].get(k, v) for k, v in tc.TensorflowParams.items() if k != "DTYPE"}
    consts["solver_params"] = {**{k: v for k, v in tc.solver_params.items() if k != "methods"}, **consts["solver_params"]}
    cj = tmp_path / "constants.json"
    cj.write_text(json.dumps(consts))
    out_dir = tmp_path / "out"
    env = {**os.environ, "QTET_DIST_BACKEND": "gloo"}
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(PKG, "qtet.py"),
                          "-p", str(out_dir), "-c", str(cj), "-m", "bins"], capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    trees = sorted(os.listdir(out_dir))
    assert len(trees) == 1, trees                                   # one data_<timestamp> directory for the whole job
    it0 = out_dir / trees[0] / "iteration_0"
    guesses = sorted(os.listdir(it0))
    n = len(guesses)
    assert n >= 2 and guesses == sorted(f"data_optimizer_{i}" for i in range(n))      # every guess once, none twice
    inits = np.array([np.loadtxt(it0 / f"data_optimizer_{i}" / "init_chis.txt") for i in range(n)])
    assert len(np.unique(inits, axis=0)) == n                       # one draw: distinct bins, no rank-local duplicates
    assert (out_dir / trees[0] / "parameters.json").exists()
