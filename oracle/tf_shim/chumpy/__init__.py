"""Unpickling stand-in for chumpy.Ch leaves of the SMPL pickle: exposes `.r` and `.shape` (batch_smpl.py:19-20,47)."""
import numpy as np


class Ch(object):
    def __init__(self, x=None):
        self.x = x

    def __setstate__(self, st):
        self.__dict__.update(st if isinstance(st, dict) else {"x": st})

    @property
    def r(self):
        return np.asarray(self.x)

    @property
    def shape(self):
        return self.r.shape
