"""Text-file helpers, behaviour-compatible with /root/reference/src/tet/data_process.py:7-105
(`%.18e`, one value per line, append when the file exists)."""
import os
import sys

import numpy as np


def writeData(data, destination, name_of_file):
    """data_process.py:7-30: savetxt with '%.18e' ('%s' for strings); appends to an existing file."""
    arr = np.array(data)
    target = os.path.join(destination, name_of_file)
    fmt = '%s' if arr.dtype.name.startswith('str') else '%.18e'
    mode = 'a' if os.path.exists(target) else 'w'
    with open(target, mode) as fh:
        np.savetxt(fh, arr, fmt=fmt)


def write_min_N(xa, xd, min_n, destination, name_of_file):
    """data_process.py:35-58: rows (x_D, x_A, min_n) of a square landscape, C order."""
    z = np.asarray(min_n).flatten(order='C')
    x = np.asarray(xd).flatten(order='C')
    y = np.asarray(xa).flatten(order='C')
    n = len(min_n)
    rows = np.stack([x[:n * n], y[:n * n], z[:n * n]], axis=1)
    writeData(data=rows, destination=destination, name_of_file=name_of_file)


def read_1D_data(destination, name_of_file):
    """data_process.py:63-76: first token of every line as float."""
    out = []
    with open(os.path.join(destination, name_of_file), 'r') as fh:
        for line in fh:
            out.append(float(line.split()[0]))
    return out


def createDir(destination, replace_query=True):
    """data_process.py:81-105: mkdir; with replace_query ask before reusing an existing directory."""
    if not replace_query:
        os.makedirs(destination, exist_ok=True)
        return
    try:
        os.mkdir(destination)
    except OSError as error:
        print(error)
        while True:
            query = input("Directory exists, replace it? [y/n] ")
            first = query[:1].lower()
            if first not in ('y', 'n'):
                print('Please answer with yes or no')
                continue
            os.makedirs(destination, exist_ok=True)
            break
        if first == 'n':
            sys.exit(0)
