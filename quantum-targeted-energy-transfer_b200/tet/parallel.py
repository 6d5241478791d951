"""Data parallelism over initial guesses / parameter points: one process per GPU, contiguous shards,
no collective during evaluation, ONE tiny collective at the end (global arg-min).

Replaces multiprocessing.Pool + np.argmin of /root/reference/src/tet/solver_mp.py:157-184.
Works with torch.distributed backend 'nccl' (device tensors over NVLink/NVSwitch) and 'gloo' (CPU tests).
"""
import numpy as np


def world():
    """(rank, world_size) -- (0, 1) when torch.distributed is not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def broadcast_object(obj, src: int = 0):
    """The object held by rank `src`, on every rank (identity when not distributed).  Used for the things that must be ONE
    consistent draw / decision across the ranks: the random initial guesses of the 'bins' method (the reference draws them
    from the unseeded global numpy RNG, solver_mp.py:70) and the timestamped output directory of the CLI (qtet.py:26)."""
    rank, ws = world()
    if ws == 1:
        return obj
    import torch.distributed as dist
    box = [obj if rank == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def shard_bounds(n: int, rank: int, world_size: int):
    """Contiguous, balanced shard [lo, hi) of n items for `rank` (first n % world shards get one more)."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_argmin(local_rows: np.ndarray, lo: int, loss_col: int):
    """Each rank holds rows [lo, lo+n) of the [G, f+1] result array.  Returns (global_index, row) of the
    smallest loss over all ranks; ties -> smallest global index (what np.argmin on the gathered array gives,
    solver_mp.py:183-184).  One all_gather of a (f+3)-double record per rank."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    width = local_rows.shape[1] if local_rows.ndim == 2 else loss_col + 1
    rec = np.full(width + 2, np.inf)
    if local_rows.shape[0] > 0:
        j = int(np.argmin(local_rows[:, loss_col]))
        rec[0] = local_rows[j, loss_col]
        rec[1] = float(lo + j)
        rec[2:] = local_rows[j]
    else:
        rec[1] = float(2 ** 52)
        rec[2:] = 0.0
    if ws == 1:
        return int(rec[1]), rec[2:].copy()
    backend = dist.get_backend()
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    mine = torch.from_numpy(rec).to(dev)
    gathered = [torch.empty_like(mine) for _ in range(ws)]
    dist.all_gather(gathered, mine)
    table = torch.stack(gathered).cpu().numpy()
    order = np.lexsort((table[:, 1], table[:, 0]))          # by loss, then by global index
    best = table[order[0]]
    return int(best[1]), best[2:].copy()


def global_argmin_device(problem, best_vars, best_loss, lo: int):
    """Device-side form of `global_argmin` for large shards: the shard's arg-min is taken by the library's own reduction
    kernel (qtet_argmin; first minimum wins, like np.argmin), only ONE (f+2)-double record per rank leaves the GPU -- through
    the all_gather (NCCL over NVLink) -- instead of the whole [G/ws, f+1] result array (solver_mp.py:182-184).
    best_vars [n, f], best_loss [n]: CUDA tensors of this rank's shard.  Returns (global_index, row[f+1])."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    f = best_vars.shape[1]
    dev = best_loss.device
    rec = torch.empty(f + 2, dtype=torch.float64, device=dev)
    if best_loss.numel() > 0:
        mn, ix = problem.argmin(best_loss)
        rec[0:1] = mn
        rec[1:2] = ix.to(torch.float64) + float(lo)
        rec[2:] = best_vars.index_select(0, ix)[0]
    else:
        rec[0] = float('inf'); rec[1] = float(2 ** 52); rec[2:] = 0.0
    if ws > 1:
        gathered = [torch.empty_like(rec) for _ in range(ws)]
        dist.all_gather(gathered, rec)
        table = torch.stack(gathered).cpu().numpy()
    else:
        table = rec.cpu().numpy()[None, :]
    order = np.lexsort((table[:, 1], table[:, 0]))          # by loss, then by global index
    best = table[order[0]]
    return int(best[1]), np.concatenate([best[2:], best[0:1]])


def gather_rows(local_rows: np.ndarray, n_total: int, lo: int):
    """All-gather variable-length shards of rows into the full [n_total, width] array (used only when the
    caller asks for every guess' result, e.g. write_data; the solver itself needs just the arg-min)."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    if ws == 1:
        return local_rows
    width = local_rows.shape[1]
    backend = dist.get_backend()
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    pad = -(-n_total // ws)
    buf = np.zeros((pad, width))
    buf[:local_rows.shape[0]] = local_rows
    mine = torch.from_numpy(buf).to(dev)
    gathered = [torch.empty_like(mine) for _ in range(ws)]
    dist.all_gather(gathered, mine)
    out = np.zeros((n_total, width))
    for r in range(ws):
        a, b = shard_bounds(n_total, r, ws)
        out[a:b] = gathered[r].cpu().numpy()[:b - a]
    return out
