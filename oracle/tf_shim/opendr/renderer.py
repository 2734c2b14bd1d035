"""opendr.renderer.ColoredRenderer as renderer.py uses it (:123-128, :153-197, :231-232)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from oracle import np_oracle  # noqa: E402


class ColoredRenderer(object):
    def __init__(self):
        self.camera = None
        self.frustum = None
        self.background_image = None
        self.v = self.f = self.vc = self.bgcolor = None

    def set(self, **kw):
        for k, val in kw.items():
            setattr(self, k, np.asarray(val))

    @property
    def r(self):
        assert self.bgcolor is None or np.all(np.asarray(self.bgcolor) == 1.0), "stand-in renders on white (renderer.py:153)"
        fr = self.frustum
        faces = np.asarray(self.f).astype(np.int64)
        return np_oracle.rasterise(self.v, faces, self.vc, self.camera.f, self.camera.c, int(fr['height']), int(fr['width']),
                                   float(fr['near']), float(fr['far']), self.background_image)
