"""TEST INFRASTRUCTURE ONLY -- plyfile.PlyData.read for the one file renderer.py opens (template-bodyparts.ply, :157-168):
a binary little-endian PLY whose vertex element is x y z nx ny nz (float32) red green blue (uint8)."""
import numpy as np


class PlyElement(object):
    def __init__(self, data):
        self.data = data


class PlyData(object):
    def __init__(self, elements):
        self.elements = elements

    @staticmethod
    def read(f):
        raw = f.read()
        end = raw.index(b"end_header\n") + len(b"end_header\n")
        hdr = raw[:end].decode("ascii")
        assert "binary_little_endian" in hdr
        n = int([ln for ln in hdr.splitlines() if ln.startswith("element vertex")][0].split()[-1])
        dt = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                       ("red", "u1"), ("green", "u1"), ("blue", "u1")])
        return PlyData([PlyElement(np.frombuffer(raw, dtype=dt, count=n, offset=end))])
