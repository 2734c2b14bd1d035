"""GPU mesh visualiser: drop-in for the reference's ``renderer.SMPLRenderer`` (renderer.py:23-115, 146-197, 221-256).

The reference renders the decoder's ``verts`` with OpenDR (an OpenGL renderer driven from chumpy objects); this module
keeps its interface -- ``SMPLRenderer(img_size, flength, face_path)``, ``__call__(verts, cam, img, do_alpha, far, near,
color_id, img_size, render_seg)``, ``rotated(verts, deg, ...)``, uint8 images out -- and runs the whole thing as two CUDA
kernels behind the C ABI (``smpl_b200_renderer_create`` / ``smpl_b200_render``, csrc/render_kernels.cu).  There is no
CPU path.

Additions: ``verts`` may be a CUDA tensor and may carry a batch dimension ((N, V, 3) -> (N, h, w, C)); pass
``as_tensor=True`` to keep the images on the device.

Documented deviations
  * ``color_id``: renderer.py:241-244 indexes ``colors.values()`` (python-2 dict order, unspecified); here 0 and ``None``
    are 'light_blue' (the reference's ``None`` branch, :239-240) and odd ids 'light_pink'.
  * a non-positive ``near`` (renderer.py:65 yields -0.2 for meshes nearer than 25 units) is invalid for the GL frustum
    OpenDR builds; here it clips at z > 0 only.  Faces with a vertex at z <= 0 are dropped rather than clipped.
  * OpenDR's ``overdraw`` anti-aliasing of silhouette edges is not restated.
  * file paths are explicit arguments with the reference's relative names as defaults; when the default files are
    absent the copies extracted from the reference (data/ref_fixtures.npz) are used.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import weakref

import numpy as np
import torch

from . import _lib, smpl_io

# renderer.py:16-20
colors = {
    'light_blue': [0.65098039, 0.74117647, 0.85882353],
    'light_pink': [.9, .7, .7],
}
# simple_renderer's three LambertianPointLights, renderer.py:171-195: position, colour
_LIGHTS = (([-200.0, -100.0, -100.0], [1.0, 1.0, 1.0]), ([800.0, 10.0, 300.0], [1.0, 1.0, 1.0]),
           ([-500.0, 500.0, 1000.0], [0.7, 0.7, 0.7]))


def _rotateY(points, angle):
    """renderer.py:138-143."""
    ry = np.array([[np.cos(angle), 0., np.sin(angle)], [0., 1., 0.], [-np.sin(angle), 0., np.cos(angle)]])
    return np.dot(points, ry)


def _rodrigues(rvec) -> np.ndarray:
    """cv2.Rodrigues(rvec)[0] (renderer.py:99-104): rotation matrix of an axis-angle vector."""
    r = np.asarray(rvec, np.float64).reshape(3)
    th = float(np.sqrt((r * r).sum()))
    if th < 1e-300:
        return np.eye(3)
    k = r / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return math.cos(th) * np.eye(3) + (1 - math.cos(th)) * np.outer(k, k) + math.sin(th) * K


def _load_faces(face_path: str) -> np.ndarray:
    if os.path.exists(face_path):
        return np.load(face_path)                                                # renderer.py:27
    return smpl_io.golden_fixtures()["faces"]


def _load_part_colors(ply_path: str) -> np.ndarray:
    """renderer.py:157-168: the red / green / blue properties of the PLY's vertices."""
    if os.path.exists(ply_path):
        raw = open(ply_path, "rb").read()
        end = raw.index(b"end_header\n") + len(b"end_header\n")
        hdr = raw[:end].decode("ascii")
        if "binary_little_endian" not in hdr:
            raise ValueError("%s: only binary little-endian PLY files are read" % ply_path)
        n = int([ln for ln in hdr.splitlines() if ln.startswith("element vertex")][0].split()[-1])
        dt = np.dtype([("xyz", "<f4", 3), ("n", "<f4", 3), ("rgb", "u1", 3)])
        return np.frombuffer(raw, dtype=dt, count=n, offset=end)["rgb"].copy()
    return smpl_io.golden_fixtures()["ply_rgb"]


class SMPLRenderer(object):
    """renderer.py:23-115."""

    def __init__(self, img_size=224, flength=500., face_path="keras_smpl/smpl_faces.npy", device=None,
                 bodypart_ply="template-bodyparts.ply"):
        if not torch.cuda.is_available():
            raise _lib.SmplB200Error("SMPLRenderer needs a CUDA device; there is no CPU path")
        self.faces = np.ascontiguousarray(_load_faces(face_path)).astype(np.int32)
        self.w = img_size
        self.h = img_size
        self.flength = flength
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if self.device.type != "cuda":
            raise _lib.SmplB200Error("SMPLRenderer needs a CUDA device; there is no CPU path")
        self._ply = bodypart_ply
        self._part_colors = None
        self._num_verts = int(self.faces.max()) + 1
        handle = C.c_void_p()
        lib = _lib.load()
        _lib.check(lib.smpl_b200_renderer_create(self.device.index or 0, self.faces.ctypes.data_as(C.POINTER(C.c_int32)),
                                                 int(self.faces.shape[0]), self._num_verts, C.byref(handle)),
                   "smpl_b200_renderer_create")
        self._handle = handle
        self._finalizer = weakref.finalize(self, lib.smpl_b200_renderer_destroy, handle)

    # -- helpers -------------------------------------------------------------------------------------------------
    def _albedo(self, color_id, render_seg):
        if render_seg:                                                           # renderer.py:156-168
            if self._part_colors is None:
                rgb = _load_part_colors(self._ply)
                self._part_colors = torch.as_tensor(rgb.astype(np.float32) / np.float32(255.0), device=self.device).contiguous()
            return self._part_colors, 1
        if color_id is None:                                                     # renderer.py:239-244
            color = colors['light_blue']
        else:
            color = list(colors.values())[color_id % len(colors)]
        return torch.tensor(color, dtype=torch.float32, device=self.device), 0

    def __call__(self, verts, cam=None, img=None, do_alpha=False, far=None, near=None, color_id=0, img_size=None,
                 render_seg=False, as_tensor=False, yrot=0.0):
        """cam is 3D [f, px, py] (renderer.py:43-85).  Returns uint8 (h, w, 3|4), or (N, h, w, 3|4) for batched verts."""
        v = torch.as_tensor(verts, dtype=torch.float32, device=self.device)
        batched = v.dim() == 3
        v = (v if batched else v[None]).contiguous()
        n = v.shape[0]
        if v.shape[1] != self._num_verts or v.shape[2] != 3:
            raise ValueError("verts must be (%d, 3) or (N, %d, 3), got %s" % (self._num_verts, self._num_verts, tuple(v.shape)))
        if img is not None:
            h, w = img.shape[:2]                                                 # :47-48
        elif img_size is not None:
            h, w = img_size[0], img_size[1]                                      # :49-51
        else:
            h, w = self.h, self.w                                                # :52-54
        h, w = int(h), int(w)
        if cam is None:
            cam = [self.flength, w / 2., h / 2.]                                 # :56-57
        cam_t = torch.as_tensor(np.asarray(cam, np.float32) if not torch.is_tensor(cam) else cam, dtype=torch.float32,
                                device=self.device).reshape(-1, 3)
        cam_t = cam_t.expand(n, 3).contiguous()
        z = v[:, :, 2]
        near_t = torch.clamp_min(z.amin(dim=1) - 25, -0.2) if near is None else torch.full((n,), float(near), device=self.device)
        far_t = torch.clamp_min(z.amax(dim=1) + 25, 25) if far is None else torch.full((n,), float(far), device=self.device)
        near_far = torch.stack([near_t, far_t], dim=1).to(torch.float32).contiguous()   # :64-67
        bg = None
        if img is not None:                                                      # :231-232: img / 255. if img.max() > 1 else img
            im = torch.as_tensor(img, device=self.device)
            if im.shape[-1] != 3:
                raise ValueError("background image must be (h, w, 3)")
            imf = im.to(torch.float32)
            k = imf if float(imf.max()) > 1 else imf * 255.0
            bg = torch.clamp(torch.round(k), 0, 255).to(torch.uint8).contiguous()
        albedo, per_vertex = self._albedo(color_id, render_seg)
        lights = np.zeros((0, 6), np.float32)
        if not render_seg:                                                       # :169-195
            lights = np.array([list(_rotateY(np.array(p), yrot)) + c for p, c in _LIGHTS], np.float32)
        channels = 4 if do_alpha else 3                                          # :252-255
        out = torch.empty((n, h, w, channels), dtype=torch.uint8, device=self.device)
        lib = _lib.load()
        ws_bytes = int(lib.smpl_b200_render_workspace_bytes(self._handle, n))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(lib.smpl_b200_render(self._handle, C.c_void_p(v.data_ptr()), C.c_void_p(cam_t.data_ptr()),
                                            C.c_void_p(near_far.data_ptr()), n, h, w, C.c_void_p(albedo.data_ptr()),
                                            per_vertex, lights.ctypes.data_as(C.POINTER(C.c_float)), int(lights.shape[0]),
                                            C.c_void_p(bg.data_ptr()) if bg is not None else None, 0, channels,
                                            C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr()), ws_bytes, stream),
                       "smpl_b200_render")
        res = out if batched else out[0]
        return res if as_tensor else res.cpu().numpy()

    def rotated(self, verts, deg, cam=None, axis='y', img=None, do_alpha=True, far=None, near=None, color_id=0,
                img_size=None, as_tensor=False):
        """renderer.py:87-115: rotate the mesh about its centroid, then render."""
        if axis == 'y':
            around = _rodrigues(np.array([0, math.radians(deg), 0]))
        elif axis == 'x':
            around = _rodrigues(np.array([math.radians(deg), 0, 0]))
        else:
            around = _rodrigues(np.array([0, 0, math.radians(deg)]))
        v = torch.as_tensor(verts, dtype=torch.float32, device=self.device)
        center = v.mean(dim=-2, keepdim=True)
        new_v = (v - center) @ torch.as_tensor(around, dtype=torch.float32, device=self.device) + center
        return self.__call__(new_v, cam, img=img, do_alpha=do_alpha, far=far, near=near, img_size=img_size,
                             color_id=color_id, as_tensor=as_tensor)

