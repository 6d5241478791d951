#!/usr/bin/env python3
"""`qtet` command line -- same flags and outputs as /root/reference/src/qtet.py:13-100
(-p/--data_path, -c/--constants, -n/--ncpus, -m/--method), running on the B200 path.
`-n` is accepted and ignored: the guesses are one device batch (sharded over GPUs under torchrun)."""
import argparse
import multiprocessing as mp
import os
import pathlib
import sys
from datetime import datetime

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import tet.constants                       # noqa: E402
from tet.solver_mp import solver_mp        # noqa: E402
from tet.data_process import createDir     # noqa: E402
from tet import parallel                   # noqa: E402


def _maybe_init_distributed():
    if 'RANK' in os.environ and 'WORLD_SIZE' in os.environ and int(os.environ['WORLD_SIZE']) > 1:
        import torch
        import torch.distributed as dist
        local = int(os.environ.get('LOCAL_RANK', 0))
        torch.cuda.set_device(local % max(1, torch.cuda.device_count()))    # (gloo tests run two ranks on one GPU)
        dist.init_process_group(os.environ.get('QTET_DIST_BACKEND', 'nccl'))


def run(argv=None):
    parser = argparse.ArgumentParser(description="qtet: targeted-energy-transfer optimisation on B200")
    parser.add_argument('-p', '--data_path', type=pathlib.Path, required=True, help='Path to create data dir.')
    parser.add_argument('-c', '--constants', type=pathlib.Path, required=False,
                        help='Path to constants json file (defaults: tet.constants).')
    parser.add_argument('-n', '--ncpus', default=mp.cpu_count(), type=int, required=False,
                        help='Kept for compatibility; ignored (guesses are batched on the GPU).')
    parser.add_argument('-m', '--method', default='bins', choices=tet.constants.solver_params['methods'], type=str,
                        required=False, help='Method of choosing initial guesses. Default "bins".')
    args = parser.parse_args(argv)
    _maybe_init_distributed()
    rank, _ = parallel.world()

    # one output tree for the whole job: rank 0 picks the (second-resolution) timestamp, the others take its choice
    data_path = parallel.broadcast_object(
        args.data_path.joinpath(f'data_{datetime.now().strftime("%Y_%h_%d_%T")}') if rank == 0 else None)
    if not args.data_path.exists() and rank == 0:
        print(f"Input data path {args.data_path} does not exist! Creating it.")
    os.makedirs(data_path, exist_ok=True)

    tf_params = {k: v for k, v in tet.constants.TensorflowParams.items() if k != 'DTYPE'}
    solver_pars = {k: v for k, v in tet.constants.solver_params.items() if k != 'methods'}
    solver_pars['method'] = args.method

    if args.constants is not None:
        if not args.constants.exists():
            raise OSError('Provided path does not exist.')
        given = tet.constants.loadConstants(args.constants)
        if given['constants'].keys() != tet.constants.system_constants.keys() or \
                given['tensorflow_params'].keys() != tf_params.keys():
            raise KeyError('Parameter json file has altered keys, please run "make" to reconfigure it.')
        tet.constants.system_constants = {**tet.constants.system_constants, **given['constants']}
        tet.constants.TensorflowParams = {**tet.constants.TensorflowParams, **given['tensorflow_params']}
        # qtet.py:66-68 merges system_constants (not solver_params) here; kept for identical behaviour
        tet.constants.solver_params = {**tet.constants.system_constants, **given['solver_params']}
        lims = given['limits']
    else:
        if rank == 0:
            print('\n* No json file was provided, using default parameters. *\n')
        lims = [[-10, 10] for _ in range(tet.constants.system_constants['sites'])]

    keys = [f'x{k}lims' for k in tet.constants.TensorflowParams['train_sites']]
    result = solver_mp(
        const=tet.constants.system_constants, trainable_vars_limits=dict(zip(keys, lims)),
        lr=tet.constants.TensorflowParams['lr'], beta_1=tet.constants.TensorflowParams['beta_1'],
        amsgrad=tet.constants.TensorflowParams['amsgrad'], target_site=tet.constants.solver_params['target'],
        data_path=data_path, method=args.method, write_data=True, cpu_count=args.ncpus)

    if rank == 0:
        tet.constants.dumpConstants(
            dictionary={'constants': result, 'tensorflow_params': tf_params, 'solver_params': solver_pars},
            path=data_path, name='parameters')
    return result


if __name__ == "__main__":
    run()
