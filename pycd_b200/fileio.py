"""POSCAR reader and `<name>.log` report writer.

Semantics follow the reference's PyCD/io.py:12-82 (read_poscar) and :163-199
(generate_report); the log layout is load-bearing because material_run parses
`precomputed_array.log` (PyCD/material_run.py:75-80).
"""
from datetime import datetime

import numpy as np


def read_poscar(path):
    """VASP5 POSCAR -> dict(lattice_matrix [A], element_types, num_elements, total_elements,
    coordinate_type, coordinates).  A VASP4 file (no element line) takes its element
    names from the comment line, like the reference (io.py:46-47)."""
    with open(path, 'r') as fh:
        lines = fh.read().splitlines()
    has_names = lines[5].split()[0].isalpha()
    lattice = np.array([[float(x) for x in lines[i].split()[:3]] for i in (2, 3, 4)])
    names = lines[5].split() if has_names else lines[0].split()
    off = 1 if has_names else 0
    counts = np.array([int(x) for x in lines[5 + off].split()], dtype=int)
    kind = lines[6 + off].split()[0]
    total = int(counts.sum())
    coords = np.array([[float(x) for x in lines[7 + off + i].split()[:3]] for i in range(total)])
    return {'lattice_matrix': lattice, 'element_types': names, 'num_elements': counts,
            'total_elements': total, 'coordinate_type': kind, 'coordinates': coords}


def format_elapsed(delta):
    """'Time elapsed: ...' line of io.py:189-197."""
    secs = delta.seconds
    out = 'Time elapsed: '
    if delta.days:
        out += '%2d days, ' % delta.days
    out += '%2d hours' % ((secs // 3600) % 24)
    out += ', %2d minutes' % ((secs // 60) % 60)
    out += ', %2d seconds' % (secs % 60)
    return out


def generate_report(start_time, dst_path, file_name, print_time_elapsed, prefix=None):
    """Writes <dst_path>/<file_name>.log = prefix + elapsed-time line."""
    with open(dst_path / (file_name + '.log'), 'w') as fh:
        if prefix:
            fh.write(prefix)
        if print_time_elapsed:
            fh.write(format_elapsed(datetime.now() - start_time))
