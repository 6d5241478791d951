"""`getCombinations` / `solver_mp` -- initial guesses, the data-parallel driver and the arg-min
(/root/reference/src/tet/solver_mp.py:17-240) on the batched B200 path.

The process pool (`mp.Pool(cpu_count)`, solver_mp.py:157-172) becomes: shard the guesses contiguously
over the ranks of torch.distributed (one process per GPU; a single process when not launched under
torchrun), run every shard as one device batch, then ONE collective for the global arg-min
(solver_mp.py:182-184).  `cpu_count` is accepted for signature compatibility and ignored.
"""
import os
import time
from itertools import product
from typing import Any

import numpy as np

from . import parallel
from .constants import solver_params, TensorflowParams
from .data_process import createDir, read_1D_data
from .Optimizer import mp_opt, mp_opt_batch, Optimizer


def getCombinations(trainable_vars_limits, train_sites=TensorflowParams['train_sites'], method='bins',
                    grid=solver_params['Npoints'], rng=None):
    """solver_mp.py:17-84.  'grid': cartesian product of linspace(lo, hi, grid) per trained site.
    'bins': the same points, histogram-binned with the box widened by 0.1, one random pick per non-empty
    bin.  The reference draws from the unseeded global numpy RNG (:70); pass `rng` for reproducibility."""
    if method not in solver_params['methods']:
        raise ValueError('Provided method not in list of supported methods [\'grid\', \'bins\']')
    spans = [np.linspace(trainable_vars_limits[f'x{i}lims'][0], trainable_vars_limits[f'x{i}lims'][1], grid)
             for i in train_sites]
    if method == 'grid':
        return list(product(*spans))
    points = np.array(list(product(*spans)))
    columns = [points[:, i] for i in range(points.shape[1])]
    extents = [(trainable_vars_limits[f'x{i}lims'][0] - 0.1, trainable_vars_limits[f'x{i}lims'][1] + 0.1)
               for i in train_sites]
    _, edges = np.histogramdd(columns, bins=grid, range=extents)
    bins_of = list(zip(*[np.digitize(columns[i], edges[i]) for i in range(len(columns))]))
    pick = (rng or np.random).choice
    combos = []
    for wanted in product(*[range(1, grid + 1) for _ in train_sites]):
        members = [pt for pt, bn in zip(points, bins_of) if tuple(bn) == wanted]
        if members:
            combos.append(members[pick(list(range(len(members))))])
    return combos


def randomCombinations(trainable_vars_limits, train_sites, count, seed=0):
    """`count` uniform random guesses inside the limits (BASELINE config 5: 10^5 seeded guesses)."""
    rng = np.random.default_rng(seed)
    lo = np.array([trainable_vars_limits[f'x{i}lims'][0] for i in train_sites], dtype=np.float64)
    hi = np.array([trainable_vars_limits[f'x{i}lims'][1] for i in train_sites], dtype=np.float64)
    return rng.uniform(lo, hi, (count, len(train_sites)))


def solver_mp(trainable_vars_limits: dict, const: dict, grid=solver_params['Npoints'], lr=0.1, beta_1=0.9,
              amsgrad=False, write_data=False, iterations=1, method='bins',
              epochs_bins=solver_params['epochs_bins'], epochs_grid=solver_params['epochs_grid'],
              target_site=solver_params['target'], main_opt=False,
              data_path=os.path.join(os.getcwd(), 'data'), cpu_count=None, combinations=None) -> dict[float, Any]:
    """solver_mp.py:87-240.  Returns `const` with const['chis'] = best variables and const['min_n'] = best
    loss.  `combinations` (extension) overrides the generated initial guesses."""
    from . import _native
    rank, ws = parallel.world()
    createDir(destination=data_path, replace_query=False)
    lims = list(trainable_vars_limits.values())
    _edge = [4.0, 3.5, 3.0, 2.5, 1.5, 0.5, 0.1]
    iteration = 0
    lim_changes = 0
    optimal_vars, min_loss = np.zeros(len(trainable_vars_limits)), const['max_N']
    train_sites = [int(list(trainable_vars_limits.keys())[i][1]) for i in range(len(trainable_vars_limits))]
    sites = const['sites']
    problem = _native.shared_problem(const)              # per-system handle, shared with Loss / Optimizer / mp_opt

    t0 = time.time()
    while iteration < iterations:
        data_path2 = os.path.join(data_path, f'iteration_{iteration}')
        createDir(destination=data_path2, replace_query=False)
        if combinations is not None:
            combos = combinations
        else:
            # 'bins' draws from the unseeded global RNG (solver_mp.py:70): rank 0 draws, every rank shards THAT list
            combos = getCombinations(trainable_vars_limits, train_sites=train_sites, method=method, grid=grid) \
                if rank == 0 else None
            combos = parallel.broadcast_object(combos)
        epochs = epochs_bins if method == 'bins' else epochs_grid
        G = len(combos)
        if rank == 0:
            print(10 * '-', f'Iteration: {iteration}, Method: {method}, Jobs: {G}, lims: {lims}', 10 * '-')
        t2 = time.time()
        lo, hi = parallel.shard_bounds(G, rank, ws)
        mine = np.asarray(combos, dtype=np.float64).reshape(G, -1)[lo:hi]
        if G > 4096 and not write_data:
            # large runs: nothing but the arg-min record leaves the GPU (device reduction + one all_gather)
            import torch
            if hi > lo:
                bv, bl = mp_opt_batch(mine, data_path2, const, target_site, epochs, lr, beta_1, amsgrad, write_data,
                                      train_sites, first_index=lo, problem=problem, quiet=True, on_device=True)
            else:
                dev = torch.device('cuda', problem.device)
                bv = torch.zeros(0, sites, dtype=torch.float64, device=dev)
                bl = torch.zeros(0, dtype=torch.float64, device=dev)
            _, best_row = parallel.global_argmin_device(problem, bv, bl, lo)
        else:
            if hi > lo:
                rows = mp_opt_batch(mine, data_path2, const, target_site, epochs, lr, beta_1, amsgrad, write_data,
                                    train_sites, first_index=lo, problem=problem, quiet=(G > 64))
            else:
                rows = np.zeros((0, sites + 1))
            _, best_row = parallel.global_argmin(rows, lo, sites)          # solver_mp.py:182-184
        dt = time.time() - t2
        optimal_vars = [float(best_row[i]) for i in range(sites)]
        min_loss = float(best_row[sites])
        if rank == 0:
            print(f"Best parameters of tries: loss={min_loss}, optimal_vars = {optimal_vars}")
            print("Code run time: ", dt, " s")
        if min_loss <= const["max_N"] / 2:                                   # solver_mp.py:190-198
            lim_changes += 1
            edge = _edge[iteration]
            lims = [[optimal_vars[i] - edge, optimal_vars[i] + edge] for i in range(len(trainable_vars_limits))]
        else:
            lr += 0.1
        iteration += 1
        if lim_changes >= 5 and min_loss >= const['max_N'] / 2:
            if rank == 0:
                print("Couldn't find TET")
            break
        if float(min_loss) <= 0.1:
            if rank == 0:
                print('TET!')
                print(f'OptimalParams:{optimal_vars}')
            break
    t1 = time.time()
    const['chis'] = optimal_vars
    const['min_n'] = min_loss
    if main_opt:
        # solver_mp.py:218-236 reads 'optimalvars.txt', which nothing in the reference writes; the branch is
        # never enabled by qtet (qtet.py:79-90).  Kept out on purpose.
        raise NotImplementedError("main_opt refinement is a dead branch in the reference (reads a file "
                                  "that is never written, solver_mp.py:231-233)")
    if rank == 0:
        print('Total solver run time: ', t1 - t0)
    return const
