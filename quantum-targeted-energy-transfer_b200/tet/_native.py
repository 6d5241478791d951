"""ctypes binding of libqtet_b200.so (C ABI in include/qtet.h).

There is no CPU fallback: importing this module fails loudly when the library is missing, and every
compute call fails loudly when no CUDA device is visible.  PyTorch is used only as the carrier of
device memory, streams and (elsewhere) torch.distributed.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QTET_B200_LIB", os.path.join(_HERE, "..", "lib", "libqtet_b200.so"))

QTET_PATH_AUTO, QTET_PATH_GENERIC, QTET_PATH_SPLIT, QTET_PATH_DIRECT = 0, 1, 2, 3
PATH_NAMES = {1: "generic", 2: "split", 3: "direct"}


class QtetError(RuntimeError):
    pass


def _load():
    path = os.path.abspath(LIB_PATH)
    if not os.path.exists(path):
        raise ImportError(
            f"libqtet_b200.so not found at {path}: build it with "
            "`python quantum-targeted-energy-transfer_b200/build_native.py` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(path)
    P = C.c_void_p
    D = C.POINTER(C.c_double)
    sig = {
        "qtet_problem_create": (C.c_int, [C.c_int, C.c_int, D, C.c_double, C.c_double, C.c_int, C.c_int, C.POINTER(P)]),
        "qtet_problem_destroy": (None, [P]),
        "qtet_problem_dim": (C.c_int, [P]),
        "qtet_problem_states": (C.c_int, [P, D]),
        "qtet_problem_time_grid": (C.c_int, [P, D]),
        "qtet_hamiltonian_host": (C.c_int, [P, D, D]),
        "qtet_problem_reserve": (C.c_int, [P, C.c_int64]),
        "qtet_set_path": (C.c_int, [P, C.c_int]),
        "qtet_set_locality_sort": (C.c_int, [P, C.c_int]),
        "qtet_get_path": (C.c_int, [P]),
        "qtet_loss_grad": (C.c_int, [P, P, C.c_int64, C.c_int, P, P, P, P]),
        "qtet_loss_grad_host": (C.c_int, [P, P, C.c_int64, C.c_int, P, P, P]),
        "qtet_trace": (C.c_int, [P, P, C.c_int64, C.c_int, P, P]),
        "qtet_adam_run": (C.c_int, [P, P, C.c_int64, C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_double, C.c_double,
                                    C.c_double, C.c_double, C.c_int, C.c_double, P, P, P, P, P, P]),
        "qtet_argmin": (C.c_int, [P, C.c_int64, P, P, P]),
        "qtet_fp64_peak": (C.c_int, [C.c_int, C.c_double, D]),
        "qtet_fp64_peak2": (C.c_int, [C.c_int, C.c_double, D, D]),
        "qtet_profile_enable": (C.c_int, [P, C.c_int]),
        "qtet_profile_read": (C.c_int, [P, D, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
        "qtet_stats": (C.c_int, [P, C.POINTER(C.c_int64), C.c_int]),
        "qtet_launch_count": (C.c_int64, []),
        "qtet_last_error": (C.c_char_p, []),
        "qtet_version": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
EXPORTS = ["qtet_problem_create", "qtet_problem_destroy", "qtet_problem_dim", "qtet_problem_states",
           "qtet_problem_time_grid", "qtet_hamiltonian_host", "qtet_problem_reserve", "qtet_set_path", "qtet_set_locality_sort", "qtet_get_path",
           "qtet_loss_grad", "qtet_loss_grad_host", "qtet_trace", "qtet_adam_run", "qtet_argmin",
           "qtet_fp64_peak", "qtet_fp64_peak2", "qtet_profile_enable", "qtet_profile_read", "qtet_stats", "qtet_launch_count", "qtet_last_error",
           "qtet_version"]


def _check(rc: int, what: str) -> int:
    if rc < 0:
        raise QtetError(f"{what} failed ({rc}): {lib.qtet_last_error().decode()}")
    return rc


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise QtetError("no CUDA device visible: the QTET B200 path has no CPU fallback")
    return torch


def launch_count() -> int:
    return int(lib.qtet_launch_count())


def fp64_peak_tflops(device: int = 0, seconds: float = 0.3) -> float:
    _torch()
    out = C.c_double(0.0)
    _check(lib.qtet_fp64_peak(device, seconds, C.byref(out)), "qtet_fp64_peak")
    return out.value


def fp64_peaks_tflops(device: int = 0, seconds: float = 0.5) -> dict:
    """{'dfma': ..., 'dmma': ...}: both FP64 paths measured on the device (the roofline uses the higher one)."""
    _torch()
    a, b = C.c_double(0.0), C.c_double(0.0)
    _check(lib.qtet_fp64_peak2(device, seconds, C.byref(a), C.byref(b)), "qtet_fp64_peak2")
    return {"dfma": a.value, "dmma": b.value}


_SHARED = {}


def shared_problem(const: dict, device=None):
    """One native handle per (system constants, device) for the whole process.  The reference builds a fresh Loss -- and
    with it a fresh graph -- for every Optimizer / every guess (Optimizer.py:108, 295); here that would mean one set of
    device workspaces per guess, so Loss, Optimizer, mp_opt and solver_mp all draw their handle from this cache.  Handles
    obtained here must not be close()d by the borrower; `release_shared_problems()` frees them all."""
    torch = _torch()
    dev = torch.cuda.current_device() if device is None else int(device)
    key = (int(const["max_N"]), int(const["sites"]), tuple(float(o) for o in const["omegas"]), float(const["coupling"]),
           int(np.int32(const["max_t"])), int(const["timesteps"]), dev)
    prob = _SHARED.get(key)
    if prob is None or getattr(prob, "_h", None) is None:
        prob = _SHARED[key] = Problem.from_const(const, dev)
    return prob


def release_shared_problems():
    for prob in _SHARED.values():
        prob.close()
    _SHARED.clear()


class Problem:
    """Opaque handle: everything about a system that does not depend on chi
    (replaces Loss.__init__, /root/reference/src/tet/HamiltonianLoss.py:20-47)."""

    def __init__(self, max_N, sites, omegas, coupling, max_t, timesteps, device=None):
        torch = _torch()
        if device is None:
            device = torch.cuda.current_device()
        self.device = int(device)
        self.max_N, self.sites, self.timesteps = int(max_N), int(sites), int(timesteps)
        om = np.ascontiguousarray(np.asarray(omegas, dtype=np.float64))
        if om.shape != (self.sites,):
            raise ValueError("omegas must have one entry per site")
        h = C.c_void_p()
        _check(lib.qtet_problem_create(self.max_N, self.sites, om.ctypes.data_as(C.POINTER(C.c_double)),
                                       float(coupling), float(max_t), self.timesteps, self.device, C.byref(h)),
               "qtet_problem_create")
        self._h = h
        self.dim = int(lib.qtet_problem_dim(h))

    @classmethod
    def from_const(cls, const: dict, device=None):
        # max_t is cast to int32 by the reference (HamiltonianLoss.py:29)
        return cls(const["max_N"], const["sites"], const["omegas"], const["coupling"],
                   int(np.int32(const["max_t"])), const["timesteps"], device)

    def close(self):
        if getattr(self, "_h", None):
            lib.qtet_problem_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- introspection -----------------------------------------------------------------
    def states(self) -> np.ndarray:
        out = np.empty((self.dim, self.sites))
        _check(lib.qtet_problem_states(self._h, out.ctypes.data_as(C.POINTER(C.c_double))), "qtet_problem_states")
        return out

    def time_grid(self) -> np.ndarray:
        out = np.empty(self.timesteps)
        _check(lib.qtet_problem_time_grid(self._h, out.ctypes.data_as(C.POINTER(C.c_double))), "qtet_problem_time_grid")
        return out

    def hamiltonian(self, chis) -> np.ndarray:
        x = np.ascontiguousarray(np.asarray(chis, dtype=np.float64))
        if x.shape != (self.sites,):
            raise ValueError("chis must have one entry per site")
        out = np.empty((self.dim, self.dim))
        _check(lib.qtet_hamiltonian_host(self._h, x.ctypes.data_as(C.POINTER(C.c_double)),
                                         out.ctypes.data_as(C.POINTER(C.c_double))), "qtet_hamiltonian_host")
        return out

    def reserve(self, batch: int):
        """Size the workspaces for batches of up to `batch` points now (no allocation / synchronisation in later calls)."""
        _check(lib.qtet_problem_reserve(self._h, int(batch)), "qtet_problem_reserve")

    def set_locality_sort(self, on: bool = True):
        """Evaluate loss_grad batches (device tensors, B >= 4096) in locality order; results are bit-identical, random batches ~20 % faster."""
        _check(lib.qtet_set_locality_sort(self._h, 1 if on else 0), "qtet_set_locality_sort")

    def set_path(self, path: int):
        _check(lib.qtet_set_path(self._h, int(path)), "qtet_set_path")

    def get_path(self) -> int:
        return int(lib.qtet_get_path(self._h))

    def profile(self, on: bool):
        _check(lib.qtet_profile_enable(self._h, 1 if on else 0), "qtet_profile_enable")

    def profile_read(self):
        """{kernel: (ms_sum, launches, points)} for 'ql_scalar' (thread-per-point QL), 'apply' (rotation replay + evolution + gradient),
        'generic_or_householder' (fused generic kernel, or the Householder kernel of the dense large-d path)."""
        ms = (C.c_double * 3)(); n = (C.c_int64 * 3)(); pts = (C.c_int64 * 3)()
        _check(lib.qtet_profile_read(self._h, ms, n, pts), "qtet_profile_read")
        return {k: (ms[i], int(n[i]), int(pts[i])) for i, k in enumerate(("ql_scalar", "apply", "generic_or_householder"))}

    def stats(self, reset: bool = False) -> dict:
        """Cumulative health counters of this handle's kernels (see qtet_stats in include/qtet.h)."""
        out = (C.c_int64 * 4)()
        _check(lib.qtet_stats(self._h, out, 1 if reset else 0), "qtet_stats")
        return {"direct_fallback": int(out[0]), "stream_overflow": int(out[1]), "nonconverged": int(out[2])}

    # ---- the hot path --------------------------------------------------------------------
    def _site(self, site) -> int:
        # HamiltonianLoss.py:49-60
        if type(site) is int or isinstance(site, np.integer):
            s = int(site)
        elif type(site) is str:
            s = int(site[-1])
        else:
            raise ValueError("Invalid type for site variable. Must be int.")
        if not 0 <= s < self.sites:
            raise ValueError(f"site {s} out of range for {self.sites} sites")
        return s

    def _dev_chis(self, chis):
        torch = _torch()
        x = chis if isinstance(chis, torch.Tensor) else torch.as_tensor(np.asarray(chis, dtype=np.float64))
        x = x.to(device=f"cuda:{self.device}", dtype=torch.float64).contiguous()
        if x.dim() == 1:
            x = x.unsqueeze(0)
        if x.dim() != 2 or x.shape[1] != self.sites:
            raise ValueError(f"chis must be [B, {self.sites}]")
        return x

    def loss_grad(self, chis, site=None, need_grad=True, need_tstar=True, out=None):
        """Device path: chis [B,f] (torch CUDA float64 or array-like) -> (loss[B], grad[B,f], t_star[B])
        as torch CUDA tensors.  Asynchronous on the current torch stream."""
        torch = _torch()
        s = self.sites - 1 if site is None else self._site(site)
        x = self._dev_chis(chis)
        B = x.shape[0]
        dev = x.device
        if out is not None:
            loss, grad, ts = out
        else:
            loss = torch.empty(B, dtype=torch.float64, device=dev)
            grad = torch.empty(B, self.sites, dtype=torch.float64, device=dev) if need_grad else None
            ts = torch.empty(B, dtype=torch.int32, device=dev) if need_tstar else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        _check(lib.qtet_loss_grad(self._h, x.data_ptr(), B, s, loss.data_ptr(),
                                  grad.data_ptr() if grad is not None else None,
                                  ts.data_ptr() if ts is not None else None, stream), "qtet_loss_grad")
        return loss, grad, ts

    def loss_grad_host(self, chis: np.ndarray, site=None):
        """Host path (numpy in, numpy out; H2D + kernels + D2H inside one C call)."""
        _torch()
        s = self.sites - 1 if site is None else self._site(site)
        x = np.ascontiguousarray(np.atleast_2d(np.asarray(chis, dtype=np.float64)))
        if x.shape[1] != self.sites:
            raise ValueError(f"chis must be [B, {self.sites}]")
        B = x.shape[0]
        loss = np.empty(B); grad = np.empty((B, self.sites)); ts = np.empty(B, dtype=np.int32)
        _check(lib.qtet_loss_grad_host(self._h, x.ctypes.data, B, s, loss.ctypes.data, grad.ctypes.data,
                                       ts.ctypes.data), "qtet_loss_grad_host")
        return loss, grad, ts

    def trace(self, chis, site=None):
        torch = _torch()
        s = self.sites - 1 if site is None else self._site(site)
        x = self._dev_chis(chis)
        out = torch.empty(x.shape[0], self.timesteps, dtype=torch.float64, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _check(lib.qtet_trace(self._h, x.data_ptr(), x.shape[0], s, out.data_ptr(), stream), "qtet_trace")
        return out

    def adam_run(self, init_chis, site=None, train_sites=(0, 1), iterations=1000, lr=0.1, beta_1=0.9,
                 beta_2=0.999, epsilon=1e-7, amsgrad=False, tol=1e-8, history=False):
        """Batched Optimizer.train (/root/reference/src/tet/Optimizer.py:99-242) for all rows of
        init_chis [B,f].  Returns a dict of torch CUDA tensors."""
        torch = _torch()
        s = self.sites - 1 if site is None else self._site(site)
        x = self._dev_chis(init_chis).clone()
        B = x.shape[0]
        dev = x.device
        mask = (C.c_int32 * self.sites)(*[1 if k in train_sites else 0 for k in range(self.sites)])
        best = torch.empty(B, self.sites, dtype=torch.float64, device=dev)
        best_loss = torch.empty(B, dtype=torch.float64, device=dev)
        epochs = torch.empty(B, dtype=torch.int32, device=dev)
        lh = vh = None
        if history:
            lh = torch.empty(max(iterations, 1), B, dtype=torch.float64, device=dev)
            vh = torch.empty(max(iterations // 10, 1), B, self.sites, dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        n = _check(lib.qtet_adam_run(self._h, x.data_ptr(), B, s, mask, int(iterations), float(lr), float(beta_1),
                                     float(beta_2), float(epsilon), int(bool(amsgrad)), float(tol), best.data_ptr(),
                                     best_loss.data_ptr(), epochs.data_ptr(),
                                     lh.data_ptr() if lh is not None else None,
                                     vh.data_ptr() if vh is not None else None, stream), "qtet_adam_run")
        return {"final_vars": x, "best_vars": best, "best_loss": best_loss, "epochs": epochs,
                "loss_hist": lh, "var_hist": vh, "epochs_run": n}

    def argmin(self, values):
        torch = _torch()
        v = values.contiguous()
        mn = torch.empty(1, dtype=torch.float64, device=v.device)
        ix = torch.empty(1, dtype=torch.int64, device=v.device)
        stream = torch.cuda.current_stream(v.device).cuda_stream
        _check(lib.qtet_argmin(v.data_ptr(), v.numel(), mn.data_ptr(), ix.data_ptr(), stream), "qtet_argmin")
        return mn, ix
