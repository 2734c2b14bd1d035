"""TEST INFRASTRUCTURE ONLY -- the three cv2 functions renderer.py's visualiser path calls (:99-104, :204-208, :216-217)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import np_oracle  # noqa: E402


def Rodrigues(rvec):
    return np_oracle.rodrigues(rvec), None


def split(img):
    return [img[:, :, i] for i in range(img.shape[2])]


def merge(channels):
    return np.dstack(channels)
