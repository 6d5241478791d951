"""`tet` -- drop-in replacement of the reference's Python package for the loss+gradient hot path
(reference: /root/reference/src/tet/__init__.py:1-12).  Same module and callable names; the
TensorFlow loss/autodiff step and the multiprocessing pool are replaced by libqtet_b200.so (sm_100a)."""
__all__ = ["HamiltonianLoss", "data_process", "constants", "solver_mp", "Optimizer"]

from .constants import system_constants, TensorflowParams, solver_params
from .HamiltonianLoss import Loss
from .Optimizer import Optimizer
from .solver_mp import solver_mp
