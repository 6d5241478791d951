"""`Loss` -- the reference's loss object (/root/reference/src/tet/HamiltonianLoss.py:12-233) on the
B200 kernels.  Same constructor and call signature; returns 0-d float64 tensors (`.numpy()` works as on
an eager tf.Tensor).  Extra, batched entry points (`batch`, `batch_trace`) expose what the kernels do
natively: many parameter points per call."""
import numpy as np

from . import _native

DTYPE = 'float64'


def _as_float(x):
    if hasattr(x, 'numpy'):
        x = x.numpy()
    return float(np.asarray(x))


class Loss:
    """Loss(const) with const as in tet.constants.system_constants (HamiltonianLoss.py:20-47)."""

    def __init__(self, const, device=None):
        self.const = const
        self.max_N_np = const['max_N']
        self.NpointsT = const['timesteps']
        self.sites = const['sites']
        self.chis = const['chis']
        self.targetState = const['sites'] - 1          # HamiltonianLoss.py:36
        self.problem = _native.shared_problem(const, device)    # one handle per system for the whole process
        self.dim = self.problem.dim
        self.states = self.problem.states()            # HamiltonianLoss.py:43, 62-110

    def __call__(self, chis, site=0, single_value=True):
        """HamiltonianLoss.py:49-60.  `site`: int, or a string ending in the site digit ('x1')."""
        import torch
        self.chis = chis
        self.targetState = self.problem._site(site)
        x = np.array([[_as_float(c) for c in chis]], dtype=np.float64)
        if single_value:
            loss, _, _ = self.problem.loss_grad(x, self.targetState, need_grad=False, need_tstar=False)
            return loss[0].cpu()
        return self.problem.trace(x, self.targetState)[0].cpu()

    # ---- batched API (what replaces the process pool) -------------------------------------
    def batch(self, chis, site=None):
        """chis [B,f] -> (loss[B], grad[B,f], t_star[B]) torch CUDA tensors."""
        return self.problem.loss_grad(chis, self.sites - 1 if site is None else site)

    def batch_trace(self, chis, site=None):
        return self.problem.trace(chis, self.sites - 1 if site is None else site)

    def landscape(self, xd, xa, site=None, vary=None, destination=None, name_of_file='min_n.txt'):
        """The (x_D, x_A, min_n) landscape the authors' plotting scripts read (HamiltonianLoss.py:232-233 evaluated over a
        grid + data_process.py:35-58 `write_min_N`): ONE batched sweep of the loss over the outer product of the 1-D arrays
        `xd` (rows) and `xa` (columns).  `vary` = the two sites whose chi are swept (default: donor 0 and acceptor f-1), the
        others keep const['chis'].  Returns min_n [len(xd), len(xa)]; with `destination` the file is written in the
        reference's format (one row `x_D x_A min_n` per point, C order, %.18e)."""
        from .data_process import write_min_N
        xd = np.asarray(xd, dtype=np.float64).ravel()
        xa = np.asarray(xa, dtype=np.float64).ravel()
        s0, s1 = (0, self.sites - 1) if vary is None else vary
        XD, XA = np.meshgrid(xd, xa, indexing='ij')
        pts = np.tile(np.array([_as_float(c) for c in self.const['chis']], dtype=np.float64), (XD.size, 1))
        pts[:, s0] = XD.ravel()
        pts[:, s1] = XA.ravel()
        target = self.sites - 1 if site is None else self.problem._site(site)
        loss, _, _ = self.problem.loss_grad(pts, target, need_grad=False, need_tstar=False)
        min_n = loss.cpu().numpy().reshape(XD.shape)
        if destination is not None:
            if len(xd) != len(xa):
                raise ValueError("write_min_N (data_process.py:35-58) indexes a SQUARE landscape")
            write_min_N(xa=XA, xd=XD, min_n=min_n, destination=destination, name_of_file=name_of_file)
        return min_n

    def createHamiltonian(self):
        """HamiltonianLoss.py:168-176 for the current self.chis (device-built, returned as numpy)."""
        return self.problem.hamiltonian([_as_float(c) for c in self.chis])

    # ---- the reference's remaining public methods, kept so that code poking at a Loss object still runs ----------
    def derive(self):
        """HamiltonianLoss.py:62-110: (re)fill self.states with the Fock basis in the reference's order."""
        self.states = self.problem.states()

    def getHash(self, state):
        """HamiltonianLoss.py:112-116."""
        s = 0
        for i in range(self.sites):
            s += np.sqrt(100 * (i + 1) + 3) * state[i]
        return s

    def getHashArray(self):
        """HamiltonianLoss.py:118-121."""
        self.T = np.array([self.getHash(self.states[i]) for i in range(self.dim)])
        self.sorted_indices = np.argsort(self.T)

    def ConstructElement(self, n, m):
        """HamiltonianLoss.py:124-165: H[n, m] for the current self.chis."""
        return self.createHamiltonian()[n, m]

    def loss(self, single_value=True):
        """HamiltonianLoss.py:218-233 for the current self.chis / self.targetState."""
        return self(self.chis, self.targetState, single_value)
