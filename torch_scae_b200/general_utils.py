"""Tiny helpers (reference general_utils.py:9-17)."""
import operator
from functools import reduce

import numpy as np


def prod(iterable):
    return reduce(operator.mul, iterable, 1)


def combined_shape(length, shape=None):
    if shape is None:
        return (length,)
    return (length, shape) if np.isscalar(shape) else (length, *shape)
