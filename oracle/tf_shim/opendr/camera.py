"""opendr.camera.ProjectPoints as renderer.py uses it (:57-62, :126): a pinhole camera x = f X / Z + c."""
import numpy as np


class ProjectPoints(object):
    def __init__(self, f=None, rt=None, t=None, k=None, c=None, v=None):
        self.f = np.asarray(f, np.float64)
        self.rt = np.zeros(3) if rt is None else np.asarray(rt, np.float64)
        self.t = np.zeros(3) if t is None else np.asarray(t, np.float64)
        self.k = np.zeros(5) if k is None else np.asarray(k, np.float64)
        self.c = np.asarray(c, np.float64)
        self.v = v
        # the reference only ever builds the identity pose without distortion (renderer.py:57-62)
        assert not self.rt.any() and not self.t.any() and not self.k.any(), "stand-in covers rt = t = k = 0 only"
