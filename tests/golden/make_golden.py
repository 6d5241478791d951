#!/usr/bin/env python3
"""Generate golden vectors from the UNMODIFIED reference sources.

Runs only in the build container (needs /root/reference, which does not exist on the GPU box).
TensorFlow / Keras are not installable here (no network, Python 3.12 vs the reference's <3.12
pin), so the reference's own ``src/tet/HamiltonianLoss.py`` and ``src/tet/Optimizer.py`` are
imported unmodified on top of a small numpy stand-in for the ``tf.*`` / ``keras`` symbols they
touch (defined below -- our code, not the reference's).  Two arithmetic modes:

* ``c64``  : ``tf.complex64`` -> numpy complex64, i.e. the reference's real precision
             (HamiltonianLoss.py:205-212 evolves in complex64);
* ``c128`` : ``tf.complex64`` -> numpy complex128, i.e. the reference's *formulas* at full
             precision.  This is the tight pin for the FP64 oracle / CUDA path.

``tf.GradientTape.gradient`` is emulated by central differences (two step sizes, Richardson)
of the reference's own loss in c128 mode; ``tf.keras.optimizers.Adam`` by the published Keras
update rule.  Output: tests/golden/reference_golden.json (committed).

Usage:  python tests/golden/make_golden.py
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference/src/tet"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.json")


# --------------------------------------------------------------------------------------
# numpy stand-in for the tensorflow / keras surface used by the reference
# --------------------------------------------------------------------------------------
class _Variable:
    def __init__(self, initial_value=0.0, dtype=None, name=None, trainable=True):
        self.v = np.float64(getattr(initial_value, "v", initial_value))
        self.trainable = trainable
        self.name = name

    def numpy(self):
        return np.float64(self.v)

    def assign(self, x):
        self.v = np.float64(getattr(x, "v", x))
        return self

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.v, dtype=dtype)

    def __float__(self):
        return float(self.v)

    def __mul__(self, o):  return self.v * o
    def __rmul__(self, o): return o * self.v
    def __add__(self, o):  return self.v + o
    def __radd__(self, o): return o + self.v
    def __sub__(self, o):  return self.v - o
    def __rsub__(self, o): return o - self.v
    def __neg__(self):     return -self.v


class _TensorArray:
    def __init__(self, dtype, size=0, **kw):
        self.dtype = dtype
        self.items = [None] * size

    def write(self, i, value):
        self.items[i] = np.asarray(value).astype(self.dtype)
        return self

    def stack(self):
        return np.array(self.items, dtype=self.dtype)


class _Adam:
    """Keras Adam (3rd-party rule, restated): m += (g-m)(1-b1); v += (g^2-v)(1-b2);
    a = lr*sqrt(1-b2^t)/(1-b1^t); x -= a*m/(sqrt(v)+eps), eps = 1e-7."""
    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad=False):
        self.learning_rate = _Variable(learning_rate)
        self.beta_1, self.beta_2, self.epsilon, self.amsgrad = beta_1, beta_2, epsilon, amsgrad
        self.iterations = _Variable(0)
        self.slots = {}

    def variables(self):
        return []

    def apply_gradients(self, grads_and_vars):
        t = int(self.iterations.v) + 1
        lr = float(self.learning_rate.v)
        a = lr * np.sqrt(1 - self.beta_2 ** t) / (1 - self.beta_1 ** t)
        for g, var in grads_and_vars:
            if g is None:
                continue
            s = self.slots.setdefault(id(var), {"m": 0.0, "v": 0.0, "vhat": 0.0})
            s["m"] += (g - s["m"]) * (1 - self.beta_1)
            s["v"] += (g * g - s["v"]) * (1 - self.beta_2)
            if self.amsgrad:
                s["vhat"] = max(s["vhat"], s["v"])
                var.assign(var.v - a * s["m"] / (np.sqrt(s["vhat"]) + self.epsilon))
            else:
                var.assign(var.v - a * s["m"] / (np.sqrt(s["v"]) + self.epsilon))
        self.iterations.assign(t)


_last_call = {}


def _function(*dargs, **dkw):
    def deco(fn):
        def wrapped(*a, **k):
            _last_call["fn"] = (fn, a, k)
            return fn(*a, **k)
        return wrapped
    if len(dargs) == 1 and callable(dargs[0]) and not dkw:
        return deco(dargs[0])
    return deco


class _Scalar(np.ndarray):
    """0-d float64 array with .numpy() like an eager tf.Tensor."""
    def numpy(self):
        return np.float64(self)


def _as_tensor(x):
    return np.asarray(x, dtype=np.float64).view(_Scalar)


class _GradientTape:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def gradient(self, loss, variables):
        fn, a, k = _last_call["fn"]
        out = []
        for var in variables:
            if not var.trainable:
                out.append(None)
                continue
            x0 = var.v
            ds = []
            for h in (2e-5, 1e-5):
                var.assign(x0 + h); fp = float(fn(*a, **k))
                var.assign(x0 - h); fm = float(fn(*a, **k))
                ds.append((fp - fm) / (2 * h))
            var.assign(x0)
            out.append(np.float64((4 * ds[1] - ds[0]) / 3))      # Richardson, O(h^4)
        return out


def make_tf(complex_dtype):
    tf = types.ModuleType("tensorflow")
    tf.__version__ = "2.9.0-numpy-standin"
    tf.float64, tf.int32, tf.complex64 = np.float64, np.int32, complex_dtype
    tf.Tensor = np.ndarray
    tf.Variable = _Variable
    tf.TensorArray = _TensorArray
    tf.constant = lambda v, dtype=None: np.asarray(v, dtype=dtype)
    tf.get_static_value = lambda x: np.asarray(x).item()
    tf.reshape = lambda x, shape: np.reshape(x, shape)
    tf.cast = lambda x, dtype: np.asarray(x).astype(dtype)
    tf.convert_to_tensor = lambda x, dtype=None: np.asarray(x, dtype=dtype)
    tf.tensordot = lambda a, b, axes: np.tensordot(a, b, axes)
    tf.reduce_sum = lambda x: np.sum(x)
    tf.reduce_max = lambda x: _as_tensor(np.max(x))
    tf.reduce_min = lambda x: _as_tensor(np.min(x))
    tf.exp = np.exp
    tf.complex = lambda re, im: complex_dtype(complex(re, im))
    tf.zeros_like = np.zeros_like
    tf.function = _function
    tf.GradientTape = _GradientTape
    tf.linalg = types.SimpleNamespace(eigh=np.linalg.eigh)
    tf.math = types.SimpleNamespace(conj=np.conj, real=np.real)
    tf.keras = types.SimpleNamespace(optimizers=types.SimpleNamespace(Adam=_Adam))
    tf.compat = types.SimpleNamespace(v1=types.SimpleNamespace(
        logging=types.SimpleNamespace(set_verbosity=lambda *_: None, ERROR=0)))
    keras = types.ModuleType("keras")
    backend = types.ModuleType("keras.backend")
    backend.clear_session = lambda: None
    backend.set_value = lambda var, val: var.assign(val)
    keras.backend = backend
    return tf, keras, backend


def load_reference(complex_dtype):
    """Import the reference's tet.constants / HamiltonianLoss / Optimizer from their own files."""
    for name in [n for n in sys.modules if n == "tet" or n.startswith("tet.")]:
        del sys.modules[name]
    tf, keras, backend = make_tf(complex_dtype)
    sys.modules["tensorflow"] = tf
    sys.modules["keras"] = keras
    sys.modules["keras.backend"] = backend
    pkg = types.ModuleType("tet"); pkg.__path__ = [REF]
    sys.modules["tet"] = pkg
    mods = {}
    for name in ("constants", "data_process", "HamiltonianLoss", "Optimizer"):
        spec = importlib.util.spec_from_file_location(f"tet.{name}", os.path.join(REF, f"{name}.py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"tet.{name}"] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods


# --------------------------------------------------------------------------------------
def const_of(N, omegas, coupling=1, max_t=25, timesteps=30):
    return {"max_N": N, "sites": len(omegas), "max_t": max_t, "omegas": list(omegas),
            "chis": [0] * len(omegas), "coupling": coupling, "timesteps": timesteps}


def main():
    gold = {"generator": "tests/golden/make_golden.py", "reference": "JAndronis/Quantum-Targeted-Energy-Transfer "
            "src/tet/HamiltonianLoss.py + Optimizer.py run unmodified on a numpy stand-in for tensorflow/keras",
            "basis": [], "hamiltonian": [], "loss": [], "train": []}

    ref64 = load_reference(np.complex64)
    Loss64 = ref64["HamiltonianLoss"].Loss
    # --- basis order + hash order (HamiltonianLoss.py:62-121)
    for (N, f) in [(3, 2), (4, 2), (20, 2), (2, 3), (3, 3), (10, 3), (3, 4)]:
        L = Loss64(const_of(N, [-3, 3] + [3] * (f - 2)))
        gold["basis"].append({"max_N": N, "sites": f, "dim": int(L.dim),
                              "states": L.states.astype(int).tolist(),
                              "sorted_indices": [int(i) for i in L.sorted_indices]})
    # --- Hamiltonians, bit-exact float64 (HamiltonianLoss.py:124-176)
    for (N, om, chis, lam) in [(3, [-3, 3], [1.5, -2.0], 1), (4, [-3, 3], [0.3, 0.7], 1),
                               (20, [-3, 3], [0.3, -0.3], 1), (6, [-2.5, 1.5], [-4.25, 7.5], 0.75),
                               (2, [-3, 3, 3], [1.0, 2.0, 3.0], 1), (4, [-3, 1, 3], [0.5, -1.5, 2.5], 1.25),
                               (10, [-3, 3, 3], [0.1, -0.2, 0.3], 1)]:
        c = const_of(N, om, lam)
        L = Loss64(c); L.chis = chis
        H = np.asarray(L.createHamiltonian(), dtype=np.float64)
        gold["hamiltonian"].append({"const": c, "chis": chis, "H_hex": [x.hex() for x in H.ravel().tolist()]})

    # --- loss and trace, in both arithmetic modes (HamiltonianLoss.py:179-233)
    cases = [
        (3, [-3, 3], [0.0, 0.0]), (3, [-3, 3], [2.0, -2.0]), (4, [-3, 3], [1.5, -1.5]),
        (4, [-3, 3], [0.5, -2.5]), (4, [-3, 3], [-7.25, 3.5]), (20, [-3, 3], [0.3, -0.3]),
        (20, [-3, 3], [1.25, -0.75]), (20, [-3, 3], [-9.5, 8.25]),
        (2, [-3, 3, 3], [1.0, 2.0, 3.0]), (4, [-3, 3, 3], [1.5, 0.25, -1.5]),
        (10, [-3, 3, 3], [0.6, 0.1, -0.6]),
    ]
    ref128 = load_reference(np.complex128)
    Loss128 = ref128["HamiltonianLoss"].Loss
    for (N, om, chis) in cases:
        c = const_of(N, om)
        f = len(om)
        for site in sorted({0, f - 1, f // 2}):
            rec = {"const": c, "chis": chis, "site": site}
            for tag, LossC in (("c64", Loss64), ("c128", Loss128)):
                L = LossC(c)
                rec[f"loss_{tag}"] = float(L(list(chis), site=site))
                rec[f"trace_{tag}"] = [float(x) for x in L(list(chis), site=site, single_value=False)]
            gold["loss"].append(rec)
        # string form of the site argument (HamiltonianLoss.py:56-57)
        L = Loss128(c)
        assert float(L(list(chis), site=f"x{f-1}")) == gold["loss"][-1]["loss_c128"]

    # --- Optimizer.train trajectories (Optimizer.py:99-242), c128 arithmetic, FD gradient
    Opt = ref128["Optimizer"].Optimizer
    train_cases = [
        (4, [-3, 3], [0.5, -0.5], [0, 1], 500, 0.1),
        (4, [-3, 3], [2.5, -2.5], [0, 1], 500, 0.1),
        (4, [-3, 3], [0.5, -2.5], [0, 1], 500, 0.1),
        (4, [-3, 3], [2.5, -0.5], [0, 1], 500, 0.1),
        (3, [-3, 3], [1.0, -3.0], [0, 1], 60, 0.1),       # does not reach TET inside 60 epochs
        (4, [-3, 3], [1.0, -1.5], [0], 200, 0.1),          # only chi_D trained
        (2, [-3, 3, 3], [2.0, 0.5, -2.0], [0, 1, 2], 40, 0.1),
        (4, [-3, 3], [3.16, -0.62], [0, 1], 80, 0.1),      # step-revert rule fires (Optimizer.py:184-186)
        (4, [-3, 3], [-0.33, -3.5], [0, 1], 30, 0.1),      # revert at epoch 3
        (3, [-3, 3], [4.0, 1.0], [0, 1], 50, 1e-9),        # >2 sub-tol moves -> early return (Optimizer.py:202-206)
    ]
    for (N, om, x0, train_sites, iters, lr) in train_cases:
        c = const_of(N, om)
        f = len(om)
        o = Opt(const=c, target_site=f - 1, Print=False, iterations=iters, train_sites=train_sites,
                opt=_Adam(learning_rate=lr, beta_1=0.9, amsgrad=False), data_path="/tmp/_unused")
        res = o(*x0, write_data=False)
        gold["train"].append({"const": c, "init": x0, "train_sites": train_sites, "iterations": iters, "lr": lr,
                              "loss": [float(x) for x in res["loss"]],
                              "best_vars": [float(x) for x in res["best_vars"]],
                              "var_data": [[float(y) for y in row] for row in res["var_data"]]})
        print("train", N, x0, "epochs", len(res["loss"]), "best", res["best_vars"], min(res["loss"]))

    with open(OUT, "w") as fh:
        json.dump(gold, fh, indent=0)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
