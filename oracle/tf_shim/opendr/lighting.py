"""opendr.lighting.LambertianPointLight as renderer.py uses it (:171-195)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from oracle import np_oracle  # noqa: E402


def LambertianPointLight(f, v, num_verts, light_pos, vc, light_color, double_sided=False):
    assert not double_sided and len(v) == num_verts
    return np_oracle.lambertian_point_light(v, f, light_pos, vc, light_color)
