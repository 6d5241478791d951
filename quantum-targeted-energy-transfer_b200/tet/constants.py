"""Configuration dictionaries and their JSON round trip.

Mirror of /root/reference/src/tet/constants.py:21-122 (same dictionary names, keys and default values,
same JSON layout) so that constants files and `qtet -c` keep working.  The only difference: 'DTYPE' is the
string 'float64' instead of `tensorflow.float64` (constants.py:3, :42) -- TensorFlow is not a dependency.
"""
import json
import os

# constants.py:21-28 -- physical system
system_constants = {
    'max_N': 3,            # number of bosons
    'sites': None,         # number of oscillators f (filled in below)
    'max_t': 25,           # evolution time window
    'omegas': [-3, 3],     # oscillator frequencies, one per site
    'chis': [0, 0],        # nonlinearity parameters, one per site
    'coupling': 1,         # nearest-neighbour coupling lambda
    'timesteps': 30,       # samples of the time grid
}
system_constants['sites'] = len(system_constants['omegas'])

# constants.py:42-48 -- optimiser
TensorflowParams = {
    'DTYPE': 'float64',
    'lr': 0.1,
    'beta_1': 0.9,
    'amsgrad': False,
    'iterations': 1000,
    'tol': 1e-8,
    'train_sites': [0, 1],
}

# constants.py:51-52
acceptor = system_constants['sites'] - 1
donor = 0

# constants.py:69-73 -- solver
solver_params = {
    'methods': ['grid', 'bins'],
    'target': acceptor,
    'Npoints': 2,
    'epochs_grid': 500,
    'epochs_bins': 1000,
}

# constants.py:80-82 -- default search box per trained site
keys = [f'x{i}lims' for i in TensorflowParams['train_sites']]
lims = [[-10, 10]] * len(keys)
TrainableVarsLimits = dict(zip(keys, lims))

# constants.py:88
plotting_params = {'plotting_resolution': 100}


def dumpConstants(dictionary: dict, path=os.getcwd(), name='constants'):
    """Write `dictionary` to <path>/<name>.json, indent=1 (constants.py:93-104)."""
    with open(os.path.join(path, name + '.json'), 'w') as fh:
        json.dump(dictionary, fh, indent=1)


def loadConstants(path='system_constants.json') -> dict:
    """Read a JSON written by dumpConstants (constants.py:109-122)."""
    with open(path, 'r') as fh:
        return json.load(fh)
