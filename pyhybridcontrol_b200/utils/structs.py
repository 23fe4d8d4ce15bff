"""Attribute-access dictionaries (the part of the reference's ``structdict`` package the hot path relies on:
attribute == item access, reference: structdict/accessors.py:57-87) and the ``ParNotSet`` sentinel
(reference: utils/func_utils.py:17-27)."""
import numpy as np


class ParNotSetType(object):
    _inst = None

    def __new__(cls):
        if cls._inst is None:
            cls._inst = super(ParNotSetType, cls).__new__(cls)
        return cls._inst

    def __bool__(self):
        return False

    def __repr__(self):
        return "ParNotSet"


ParNotSet = ParNotSetType()


class StructDict(dict):
    """dict whose items are also attributes."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, value):
        if key.startswith("_"):
            object.__setattr__(self, key, value)
        else:
            self[key] = value

    def __delattr__(self, key):
        try:
            del self[key]
        except KeyError:
            raise AttributeError(key)

    def get_sub_list(self, keys):
        return [self[k] for k in keys]

    def get_sub_struct(self, keys):
        return type(self)((k, self[k]) for k in keys)

    def copy(self):
        return type(self)(self)

    def deepcopy(self):
        import copy
        return copy.deepcopy(self)


def atleast_2d_col(arr, dtype=None):
    """scalars -> (1,1), vectors -> column (reference: utils/matrix_utils.py:31-39)."""
    arr = np.asanyarray(arr, dtype=dtype)
    if arr.ndim == 0:
        return arr.reshape(1, 1)
    if arr.ndim == 1:
        return arr[:, np.newaxis]
    return arr
