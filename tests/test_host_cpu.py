"""CPU-only tests: the C-ABI library loads and exports every symbol include/qtet.h declares (no compute
calls without a GPU), host-side logic of the tet package, and the multi-process arg-min over gloo."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "quantum-targeted-energy-transfer_b200")


@pytest.fixture(scope="module")
def built_lib():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    return os.path.join(PKG, "lib", "libqtet_b200.so")


def test_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "qtet.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(qtet_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 15
    lib = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/qtet.h but not exported"
    from tet import _native
    assert sorted(_native.EXPORTS) == declared
    lib.qtet_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.qtet_version()


def test_no_cpu_fallback_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tet import _native
    with pytest.raises(_native.QtetError):
        _native.Problem(3, 2, [-3, 3], 1.0, 25, 30)
    # argument validation happens before any CUDA call
    h = ctypes.c_void_p()
    om = (ctypes.c_double * 2)(-3, 3)
    assert _native.lib.qtet_problem_create(0, 2, om, 1.0, 25.0, 30, 0, ctypes.byref(h)) == -1
    assert b"max_N" in _native.lib.qtet_last_error()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), f"{fn} mentions the oracle"


def test_constants_json_roundtrip(tmp_path, built_lib):
    from tet import constants
    assert constants.system_constants["sites"] == 2 and constants.solver_params["target"] == 1
    tf_params = {k: v for k, v in constants.TensorflowParams.items() if k != "DTYPE"}
    d = {"constants": constants.system_constants, "tensorflow_params": tf_params,
         "solver_params": {k: v for k, v in constants.solver_params.items() if k != "methods"},
         "limits": [[-10, 10], [-10, 10]]}
    constants.dumpConstants(d, path=str(tmp_path))
    assert constants.loadConstants(str(tmp_path / "constants.json")) == json.loads(json.dumps(d))
    assert set(constants.TrainableVarsLimits) == {"x0lims", "x1lims"}


def test_data_process(tmp_path, built_lib):
    from tet.data_process import writeData, read_1D_data, createDir, write_min_N
    createDir(str(tmp_path / "a" / "b"), replace_query=False)
    writeData([1.5, 2.5], str(tmp_path / "a" / "b"), "x.txt")
    writeData([3.5], str(tmp_path / "a" / "b"), "x.txt")          # appends
    assert read_1D_data(str(tmp_path / "a" / "b"), "x.txt") == [1.5, 2.5, 3.5]
    assert (tmp_path / "a" / "b" / "x.txt").read_text().splitlines()[0] == "1.500000000000000000e+00"
    xa, xd = np.meshgrid([0.0, 1.0], [2.0, 3.0])
    write_min_N(xa, xd, np.array([[1.0, 2.0], [3.0, 4.0]]), str(tmp_path), "grid.txt")
    rows = np.loadtxt(tmp_path / "grid.txt")
    assert rows.shape == (4, 3) and rows[1].tolist() == [2.0, 1.0, 2.0]


def test_get_combinations_matches_oracle(built_lib):
    from oracle import qtet_oracle as O
    from tet.solver_mp import getCombinations, randomCombinations
    lims = {"x0lims": [-10, 10], "x1lims": [-10, 10]}
    assert getCombinations(lims, [0, 1], "grid", 3) == O.get_combinations(lims, [0, 1], "grid", 3)
    a = getCombinations(lims, [0, 1], "bins", 3, rng=np.random.default_rng(0))
    b = O.get_combinations(lims, [0, 1], "bins", 3, rng=np.random.default_rng(0))
    assert np.array_equal(np.array(a), np.array(b))
    with pytest.raises(ValueError):
        getCombinations(lims, [0, 1], "random", 3)
    r = randomCombinations(lims, [0, 1], 1000, seed=0)
    assert r.shape == (1000, 2) and np.array_equal(r, np.random.default_rng(0).uniform(-10, 10, (1000, 2)))


def test_shard_bounds(built_lib):
    from tet.parallel import shard_bounds
    for n in (0, 1, 7, 8, 100003):
        for ws in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {pkg!r})
import torch.distributed as dist
from tet import parallel
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:{port}', rank=int(sys.argv[1]), world_size=2)
rank, ws = parallel.world()
G, f = 11, 2
rng = np.random.default_rng(0)
table = rng.uniform(0, 4, (G, f + 1))
table[7, f] = -1.0; table[3, f] = -1.0           # tie: the smaller global index (3) must win
lo, hi = parallel.shard_bounds(G, rank, ws)
idx, row = parallel.global_argmin(table[lo:hi], lo, f)
full = parallel.gather_rows(table[lo:hi], G, lo)
assert idx == 3 and np.array_equal(row, table[3]), (idx, row)
assert np.array_equal(full, table)
# an empty shard still takes part in the collective
idx2, row2 = parallel.global_argmin(table[:1] if rank == 0 else table[:0], 0, f)
assert idx2 == 0 and np.array_equal(row2, table[0])
dist.destroy_process_group()
print('ok', rank)
"""


def test_global_argmin_gloo_world2(tmp_path, built_lib):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(pkg=PKG, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=180)
        assert p.returncode == 0, err[-2000:]
        assert "ok" in out


def test_bench_cpu_baseline_sample_is_bounded():
    """bench.py's CPU baseline (the oracle over a process pool) must stay a bounded sample for every workload: at d = 101 an
    unpinned BLAS oversubscribed the cores and one `--workload dimer100` run spent 13 minutes in it."""
    import time
    sys.path.insert(0, ROOT)
    import bench
    w = bench.WORKLOADS["dimer100"]
    const = bench.make_const(w)
    pts = bench.make_points(w)[:4096]
    t0 = time.perf_counter()
    rate, n, dt = bench.cpu_rate(const, pts, cores=2, seconds_target=2.0)
    wall = time.perf_counter() - t0
    assert 16 <= n <= len(pts) and rate > 0
    assert dt < 20.0 and wall < 40.0, (n, dt, wall)


_BCAST_WORKER = r"""
import os, sys, json
sys.path.insert(0, {pkg!r})
import numpy as np
import torch.distributed as dist
dist.init_process_group('gloo')
from tet import parallel
from tet.solver_mp import getCombinations
rank = dist.get_rank()
np.random.seed(1000 + rank)                      # the ranks' global RNGs differ, as they do in real life
lims = {{"x0lims": [-10, 10], "x1lims": [-10, 10]}}
combos = getCombinations(lims, train_sites=[0, 1], method='bins', grid=4) if rank == 0 else None
combos = parallel.broadcast_object(combos)
path = parallel.broadcast_object('data_%d' % os.getpid() if rank == 0 else None)
json.dump({{"combos": np.asarray(combos).tolist(), "path": path}}, open({out!r} + f'/r{{rank}}.json', 'w'))
dist.barrier()
dist.destroy_process_group()
"""


def test_rank0_draw_is_shared_by_all_ranks_gloo_world2(tmp_path):
    """The random 'bins' guesses and the CLI's timestamped directory must be ONE draw for the whole job."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text(_BCAST_WORKER.format(pkg=PKG, out=str(tmp_path)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    a = json.load(open(tmp_path / "r0.json")); b = json.load(open(tmp_path / "r1.json"))
    assert a == b and len(a["combos"]) >= 4
