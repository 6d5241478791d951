"""`Optimizer` / `mp_opt` -- the reference's per-guess optimiser API
(/root/reference/src/tet/Optimizer.py:19-321) on the batched B200 path.

The reference builds one TensorFlow graph + one Keras Adam per initial guess and runs them in a process
pool.  Here a guess is one row of a batch: `Optimizer.train` is the batch-of-one case of
`Problem.adam_run`, and `mp_opt_batch` runs every guess of a solver iteration in one call.  All the
stop / revert / learning-rate rules of Optimizer.py:146-224 are applied per guess on the device.
"""
import os
import time

import numpy as np

from . import _native
from .constants import TensorflowParams
from .data_process import createDir, writeData
from .HamiltonianLoss import Loss


class Adam:
    """Configuration carrier standing in for `tf.keras.optimizers.Adam(learning_rate, beta_1, amsgrad)`
    (Optimizer.py:38, :302).  Keras defaults: beta_2=0.999, epsilon=1e-7."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad=False):
        self.learning_rate = learning_rate
        self.beta_1, self.beta_2, self.epsilon, self.amsgrad = beta_1, beta_2, epsilon, amsgrad
        self.iterations = 0


class _Var:
    """Minimal stand-in for a tf.Variable holding one chi (`.numpy()`, `.assign()`)."""

    def __init__(self, value, trainable, name=None):
        self.value, self.trainable, self.name = float(value), trainable, name

    def numpy(self):
        return np.float64(self.value)

    def assign(self, v):
        self.value = float(v.numpy() if hasattr(v, 'numpy') else v)
        return self


class Optimizer:
    """Same constructor as Optimizer.py:33-40.  `opt` is a `tet.Optimizer.Adam` (or any object with
    learning_rate / beta_1 / beta_2 / epsilon / amsgrad attributes)."""

    def __init__(self, const: dict, target_site: int, DataExist=False, Print=True,
                 iterations=TensorflowParams['iterations'], train_sites=TensorflowParams['train_sites'],
                 opt=None, data_path=os.path.join(os.getcwd(), 'data_optimizer'), problem=None):
        self.const = const
        self.coupling = const['coupling']
        self.max_t = const['max_t']
        self.max_n = const['max_N']
        self.omegas = const['omegas']
        self.sites = const['sites']
        self.train_sites = train_sites
        self.target_site = target_site
        self.DataExist = DataExist
        self.data_path = data_path
        self.iter = iterations
        self.opt = opt if opt is not None else Adam()
        self.lr = self.opt.learning_rate
        self.Print = Print
        self.vars = [None for _ in range(len(const['chis']))]
        self.DTYPE = 'float64'
        self.tol = TensorflowParams['tol']
        self.loss = None
        self._problem = problem
        self._adam_state = None
        self.results = None

    # -- Optimizer.py:64-77
    def __call__(self, *args, write_data=False):
        self.init_chis = list(args)
        if self.DataExist:
            return None
        if write_data:
            createDir(self.data_path, replace_query=False)
        self._train(write_data=bool(write_data))
        return self.results

    def _ensure_loss(self):
        if self.loss is None:
            self.loss = Loss(const=self.const) if self._problem is None else _LossOn(self.const, self._problem)
        return self.loss

    def _init_vars(self, initial_chis):
        for i in range(len(self.const['chis'])):           # Optimizer.py:118-130
            if self.vars[i] is None:
                self.vars[i] = _Var(initial_chis[i], trainable=(i in self.train_sites), name=f'chi{i}')

    # -- Optimizer.py:80-82
    def compute_loss(self):
        return self._ensure_loss()(self.vars, site=self.target_site)

    # -- Optimizer.py:85-91: (grads, loss); None for non-trainable variables
    def get_grads(self):
        import torch
        L = self._ensure_loss()
        x = np.array([[v.numpy() for v in self.vars]], dtype=np.float64)
        loss, grad, _ = L.problem.loss_grad(x, L.problem._site(self.target_site))
        g = grad[0].cpu()
        grads = [g[i] if self.vars[i].trainable else None for i in range(len(self.vars))]
        return grads, loss[0].cpu()

    # -- Optimizer.py:94-96 (Keras Adam update on the host-side variables; single-guess convenience)
    def apply_grads(self, grads):
        st = self._adam_state
        if st is None:
            st = self._adam_state = {'t': 0, 'm': [0.0] * len(self.vars), 'v': [0.0] * len(self.vars),
                                     'vh': [0.0] * len(self.vars)}
        o = self.opt
        st['t'] += 1
        o.iterations = st['t']
        a = float(o.learning_rate) * np.sqrt(1.0 - o.beta_2 ** st['t']) / (1.0 - o.beta_1 ** st['t'])
        for i, (g, var) in enumerate(zip(grads, self.vars)):
            if g is None:
                continue
            g = float(g)
            st['m'][i] += (g - st['m'][i]) * (1.0 - o.beta_1)
            st['v'][i] += (g * g - st['v'][i]) * (1.0 - o.beta_2)
            den = st['v'][i]
            if o.amsgrad:
                st['vh'][i] = max(st['vh'][i], st['v'][i]); den = st['vh'][i]
            var.assign(var.value - a * st['m'][i] / (np.sqrt(den) + o.epsilon))

    # -- Optimizer.py:99-242
    def train(self, initial_chis):
        L = self._ensure_loss()
        self._init_vars(initial_chis)
        t0 = time.time()
        x0 = np.array([[v.numpy() for v in self.vars]], dtype=np.float64)
        out = L.problem.adam_run(x0, site=L.problem._site(self.target_site), train_sites=self.train_sites,
                                 iterations=self.iter, lr=float(self.lr), beta_1=self.opt.beta_1,
                                 beta_2=self.opt.beta_2, epsilon=self.opt.epsilon, amsgrad=self.opt.amsgrad,
                                 tol=self.tol, history=True)
        n = int(out['epochs'][0].item())
        mylosses = [self.max_n] + [float(v) for v in out['loss_hist'][:n, 0].cpu().numpy()]
        nsamp = n // 10
        vh = out['var_hist'][:nsamp, 0, :].cpu().numpy() if nsamp else np.zeros((0, len(self.vars)))
        var_data = [[np.float64(vh[s, k]) for s in range(nsamp)] for k in range(len(self.vars))]
        best = out['best_vars'][0].cpu().numpy()
        final = out['final_vars'][0].cpu().numpy()
        for k, var in enumerate(self.vars):
            var.assign(final[k])
        best_vars = [_Var(b, trainable=False) for b in best]
        self.opt.iterations = n
        if self.Print:
            dt = time.time() - t0
            for e in range(0, n, 50):                    # Optimizer.py:155-158 (progress lines: loss and variables BEFORE the step)
                xs = x0[0] if e == 0 else vh[e // 10 - 1]
                print(f'Loss:{mylosses[e + 1]}, ', *[f'x{j}: {xs[j]}, ' for j in range(len(self.vars))], f', epoch:{e}')
            if n < self.iter and not abs(mylosses[n]) < 0.1:
                print('Stopped training because a trained variable moved by less than tol =', self.tol,
                      'in more than two steps')             # Optimizer.py:202-208
            print(*[f"\nApproximate value of chi_{j}: {best_vars[j].numpy()}" for j in range(len(self.vars))],
                  "\nLoss:", min(mylosses), "\nOptimizer Iterations:", n, "\nTraining Time:", dt,
                  "\n" + 60 * "-", "\nParameters:", "\nOmegas", self.omegas, "| N:", self.max_n,
                  "| Sites: ", self.sites, "| Total timesteps:", self.max_t,
                  "| Coupling Lambda:", self.coupling, "\n" + 60 * "-")
        return mylosses, var_data, best_vars

    # -- Optimizer.py:245-263
    def _train(self, write_data: bool):
        mylosses, var_data, best_vars = self.train(self.init_chis)
        if write_data:
            write_guess_files(self.data_path, mylosses[1:], self.init_chis, var_data)
        self.results = {'loss': mylosses[1:], 'var_data': var_data,
                        'best_vars': [v.numpy() for v in best_vars]}


class _LossOn(Loss):
    """Loss bound to an existing native problem handle (avoids rebuilding it per guess)."""

    def __init__(self, const, problem):
        self.const = const
        self.max_N_np = const['max_N']
        self.NpointsT = const['timesteps']
        self.sites = const['sites']
        self.chis = const['chis']
        self.targetState = const['sites'] - 1
        self.problem = problem
        self.dim = problem.dim
        self.states = problem.states()


def write_guess_files(data_path, losses, init_chis, var_data):
    """Optimizer.py:248-257: losses.txt, init_chis.txt, x{i}trajectory.txt."""
    writeData(data=losses, destination=data_path, name_of_file='losses.txt')
    writeData(data=init_chis, destination=data_path, name_of_file='init_chis.txt')
    for i in range(len(var_data)):
        writeData(data=var_data[i], destination=data_path, name_of_file=f'x{i}trajectory.txt')


# ----------------------------- pool worker replacement --------------------------------------
def mp_opt(i: int, combination: list, iteration_path: str, const: dict, target_site: int, iterations: int,
           lr: float, beta_1: float, amsgrad_bool: bool, write_data: bool, train_sites: list) -> np.ndarray:
    """Optimizer.py:268-321 for ONE guess: returns np.array([*best_vars, min(loss)])."""
    rows = mp_opt_batch([combination], iteration_path, const, target_site, iterations, lr, beta_1,
                        amsgrad_bool, write_data, train_sites, first_index=i)
    return rows[0]


def mp_opt_batch(combinations, iteration_path, const, target_site, iterations, lr, beta_1, amsgrad_bool,
                 write_data, train_sites, first_index=0, problem=None, quiet=False, on_device=False):
    """All guesses of one solver iteration in ONE device batch (replaces pool.starmap_async(mp_opt, args),
    solver_mp.py:157-172).  Returns the [G, f+1] array solver_mp.py:182 builds from the pool results."""
    f = len(const['chis'])
    combos = np.asarray(combinations, dtype=np.float64).reshape(len(combinations), -1)
    G = combos.shape[0]
    x0 = np.zeros((G, f))                                  # Optimizer.py:307-312
    for col, site in enumerate(train_sites):
        x0[:, site] = combos[:, col]
    own = False                                            # handles come from the per-system cache, never closed here
    prob = problem if problem is not None else _native.shared_problem(const)
    out = prob.adam_run(x0, site=prob._site(target_site), train_sites=train_sites, iterations=iterations,
                        lr=lr, beta_1=beta_1, amsgrad=amsgrad_bool, tol=TensorflowParams['tol'],
                        history=bool(write_data))
    if on_device and not write_data:
        # large shards: the results stay on the GPU, the caller reduces them there (parallel.global_argmin_device)
        if own:
            prob.close()
        return out['best_vars'], out['best_loss']
    best = out['best_vars'].cpu().numpy()
    best_loss = out['best_loss'].cpu().numpy()
    if write_data:
        ep = out['epochs'].cpu().numpy()
        lh = out['loss_hist'].cpu().numpy()
        vh = out['var_hist'].cpu().numpy()
        for g in range(G):
            path = os.path.join(iteration_path, f'data_optimizer_{first_index + g}')   # Optimizer.py:292
            createDir(path, replace_query=False)
            n = int(ep[g])
            var_data = [[vh[s, g, k] for s in range(n // 10)] for k in range(f)]
            write_guess_files(path, lh[:n, g], list(x0[g]), var_data)
    if not quiet:
        for g in range(G):
            print(f'Job {first_index + g}: Done')          # Optimizer.py:319
    if own:
        prob.close()
    return np.concatenate([best, best_loss[:, None]], axis=1)
