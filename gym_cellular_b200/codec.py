"""Host-side mixed-radix tabular codec for single elements (Python ints, unbounded).

Same contract as the reference's gym_cellular/envs/utils/generalized_space_transformations.py:1-23
(cell 0 is the least-significant digit, digit c = value_c - min(space_c), radix c = len(space_c)) and
utils/space_transformations.py:4-23 (fixed radix).  The batched device versions are gc_encode /
gc_decode and the `index` output of gc_step.
"""
from functools import reduce

import numpy as np


def generalized_cellular2tabular(alist, intracellular_space_set):
    if len(alist) != len(intracellular_space_set):
        raise AssertionError("one value per cell expected")
    digits = [int(v) - min(sp) for v, sp in zip(alist, intracellular_space_set)]
    radices = [len(sp) for sp in intracellular_space_set]
    # Horner from the most significant cell down
    return reduce(lambda acc, dr: acc * dr[1] + dr[0], zip(reversed(digits), reversed(radices)), 0)


def generalized_tabular2cellular(anint, intracellular_space_set):
    out = []
    rest = int(anint)
    for sp in intracellular_space_set:
        rest, digit = divmod(rest, len(sp))
        out.append(digit + min(sp))
    return out


def cellular2tabular(alist, intracellular_size, n_cells):
    assert len(alist) == n_cells and max(alist) < intracellular_size and min(alist) >= 0
    return generalized_cellular2tabular(alist, [range(intracellular_size)] * n_cells)


def tabular2cellular(anint, intracellular_size, n_cells):
    assert 0 <= anint < intracellular_size ** n_cells
    return np.array(generalized_tabular2cellular(anint, [range(intracellular_size)] * n_cells), dtype=int)
