"""Host-side model / table I/O for the SMPL decoder path.

Everything here is plain numpy: it reads the same input files the reference reads
and hands flat fp32/int32 arrays to the C-ABI library (see include/smpl_b200.h).

Reference behaviour mirrored (file:line relative to /root/reference):
  * SMPL pickle keys and reshapes .......... keras_smpl/batch_smpl.py:31-94
  * part-vertex tables ...................... keras_smpl/projects_to_seg.py:16-24,34-37
  * mean parameters (h5) .................... keras_smpl/concat_mean_param.py:8-25
"""
from __future__ import annotations

import io
import os
import pickle
import sys
import types
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

NUM_VERTS = 6890
NUM_JOINTS = 24
NUM_BETAS = 10
NUM_POSE_BASIS = 207
NUM_PARTS = 31

# Standard SMPL kinematic tree (kintree_table[0]); entry 0 is 2**32-1 in the pkl and is never read
# (batch_smpl.py:71, :206-211).
SMPL_PARENTS = np.array([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21], np.int32)

# The reference's own data files (part tables, mean-parameter h5 bytes, template geometry of the PLY), extracted by
# tools/extract_reference_data.py; used when the CWD-relative files the reference opens are not present.
_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "ref_fixtures.npz")


@dataclass
class SmplHostModel:
    """The SMPL constants in the layout the reference's ``SMPLLayer.build`` produces."""

    v_template: np.ndarray      # (V,3) f32
    shapedirs: np.ndarray       # (10, V*3) f32, column = v*3+c          batch_smpl.py:50-55
    posedirs: np.ndarray        # (207, V*3) f32                          batch_smpl.py:64-68
    J_regressor: np.ndarray     # (V,24) f32 (dense, already transposed)  batch_smpl.py:58-61
    parents: np.ndarray         # (24,) int32, parents[0] = -1           batch_smpl.py:71
    lbs_weights: np.ndarray     # (V,24) f32                              batch_smpl.py:76-79
    joint_regressor: np.ndarray  # (V,19) f32 cocoplus                    batch_smpl.py:82-85
    meta: dict = field(default_factory=dict)

    def validate(self) -> "SmplHostModel":
        V = self.v_template.shape[0]
        assert self.v_template.shape == (V, 3)
        assert self.shapedirs.shape == (NUM_BETAS, V * 3), self.shapedirs.shape
        assert self.posedirs.shape == (NUM_POSE_BASIS, V * 3), self.posedirs.shape
        assert self.J_regressor.shape == (V, NUM_JOINTS)
        assert self.lbs_weights.shape == (V, NUM_JOINTS)
        assert self.parents.shape == (NUM_JOINTS,)
        assert self.joint_regressor.shape[0] == V
        for i in range(1, NUM_JOINTS):
            assert 0 <= self.parents[i] < i, "kinematic tree must be topologically ordered"
        return self


# ----------------------------------------------------------------------------------------------
# python-2 SMPL pickle (chumpy + scipy.sparse objects), loaded without chumpy installed
# ----------------------------------------------------------------------------------------------
class _ChStub(object):
    """Stand-in for chumpy.ch.Ch: keeps whatever state the pickle holds, exposes ``.r``
    (the reference's ``undo_chumpy`` reads ``x.r``, batch_smpl.py:19-20)."""

    def __init__(self, *a, **k):
        self._state = {}

    def __setstate__(self, state):
        self._state = state if isinstance(state, dict) else {"x": state}

    def __getstate__(self):
        return self._state

    @property
    def r(self):
        st = self._state
        for key in ("x", b"x", "_x", "r"):
            if key in st:
                return np.asarray(st[key])
        for v in st.values():
            if isinstance(v, np.ndarray):
                return v
        raise ValueError("chumpy object without array payload: keys=%r" % (list(st),))

    @property
    def shape(self):
        return self.r.shape


class _SmplUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("chumpy"):
            return _ChStub
        if module.startswith("scipy.sparse"):
            import scipy.sparse as sp

            if hasattr(sp, name):
                return getattr(sp, name)
        if module == "numpy.core.multiarray" or module == "numpy._core.multiarray":
            import numpy.core.multiarray as m  # noqa

            return getattr(m, name)
        return super().find_class(module, name)


def _undo_chumpy(x):
    return x if isinstance(x, np.ndarray) else np.asarray(x.r)


def _dense(x):
    return np.asarray(x.todense()) if hasattr(x, "todense") else np.asarray(x)


def load_smpl_pkl(path: str, dtype=np.float32) -> SmplHostModel:
    """Load ``neutral_smpl_with_cocoplus_reg.pkl`` (batch_smpl.py:33-87).

    Raises IOError/FileNotFoundError on a missing file exactly as the reference's ``open`` does.
    """
    with open(path, "rb") as f:
        dd = _SmplUnpickler(f, encoding="latin1").load()
    dd = {(k.decode() if isinstance(k, bytes) else k): v for k, v in dd.items()}
    v_template = _undo_chumpy(dd["v_template"]).astype(dtype)
    V = v_template.shape[0]
    sd = _undo_chumpy(dd["shapedirs"])
    num_betas = sd.shape[-1]
    shapedirs = np.reshape(sd, [-1, num_betas]).T.astype(dtype)
    J_regressor = _dense(dd["J_regressor"]).T.astype(dtype)
    pdirs = _undo_chumpy(dd["posedirs"])
    posedirs = np.reshape(pdirs, [-1, pdirs.shape[-1]]).T.astype(dtype)
    parents = np.asarray(dd["kintree_table"])[0].astype(np.int64).astype(np.int32)  # 2**32-1 -> -1
    parents[0] = -1
    weights = _undo_chumpy(dd["weights"]).astype(dtype)
    if "cocoplus_regressor" in dd:
        jr = _dense(dd["cocoplus_regressor"]).T.astype(dtype)
    else:
        jr = np.zeros((V, 19), dtype)
    return SmplHostModel(np.ascontiguousarray(v_template), np.ascontiguousarray(shapedirs),
                         np.ascontiguousarray(posedirs), np.ascontiguousarray(J_regressor), parents,
                         np.ascontiguousarray(weights), np.ascontiguousarray(jr),
                         meta={"source": path}).validate()


def save_smpl_pkl(model: SmplHostModel, path: str, protocol: int = 2, use_chumpy_stubs: bool = True) -> None:
    """Write a pickle with the *same object structure* as the HMR/SMPL release (chumpy ``Ch`` leaves,
    scipy CSC regressors, uint32 kintree) so the loader's format handling can be exercised without
    the real (licensed, absent) file."""
    import scipy.sparse as sp

    V = model.v_template.shape[0]
    fake_mods = {}
    if use_chumpy_stubs:
        ch_pkg = types.ModuleType("chumpy")
        ch_mod = types.ModuleType("chumpy.ch")

        class Ch(object):  # pickled by reference as chumpy.ch.Ch
            def __init__(self, x=None):
                self.x = x

            def __getstate__(self):
                return {"x": self.x, "_dirty_vars": set(), "_itr": None}

            def __setstate__(self, st):
                self.__dict__.update(st)

        Ch.__module__ = "chumpy.ch"
        Ch.__qualname__ = "Ch"
        ch_mod.Ch = Ch
        ch_pkg.ch = ch_mod
        fake_mods = {"chumpy": ch_pkg, "chumpy.ch": ch_mod}
        wrap = lambda a: Ch(np.asarray(a, np.float64))  # noqa: E731
    else:
        wrap = lambda a: np.asarray(a, np.float64)  # noqa: E731
    kintree = np.zeros((2, NUM_JOINTS), np.uint32)
    kintree[0] = model.parents.astype(np.int64) % (1 << 32)
    kintree[1] = np.arange(NUM_JOINTS)
    dd = {
        "v_template": wrap(model.v_template),
        "shapedirs": wrap(model.shapedirs.T.reshape(V, 3, -1)),
        "posedirs": wrap(model.posedirs.T.reshape(V, 3, -1)),
        "J_regressor": sp.csc_matrix(model.J_regressor.T.astype(np.float64)),
        "kintree_table": kintree,
        "weights": wrap(model.lbs_weights),
        "cocoplus_regressor": sp.csc_matrix(model.joint_regressor.T.astype(np.float64)),
        "f": np.zeros((1, 3), np.uint32),
        "bs_style": "lbs",
        "bs_type": "lrotmin",
    }
    saved = {k: sys.modules.get(k) for k in fake_mods}
    sys.modules.update(fake_mods)
    try:
        with open(path, "wb") as f:
            pickle.dump(dd, f, protocol=protocol)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


# ----------------------------------------------------------------------------------------------
# fixtures shipped with the repo (extracted from the reference's data files)
# ----------------------------------------------------------------------------------------------
_fix_cache = {}


def golden_fixtures(path: Optional[str] = None) -> dict:
    path = os.path.normpath(path or _GOLDEN)
    if path not in _fix_cache:
        with np.load(path) as z:
            _fix_cache[path] = {k: z[k] for k in z.files}
    return _fix_cache[path]


# Approximate neutral-SMPL T-pose joint locations in metres (public SMPL knowledge, rounded); only used
# to give the synthetic model a plausible skeleton for the PLY template.
_APPROX_JOINTS = np.array([
    [-0.002, -0.223, 0.028], [0.069, -0.314, 0.021], [-0.069, -0.314, 0.022], [-0.004, -0.114, 0.002],
    [0.103, -0.690, 0.017], [-0.107, -0.696, 0.015], [0.001, 0.021, 0.027], [0.089, -1.088, -0.027],
    [-0.092, -1.092, -0.023], [0.003, 0.074, 0.028], [0.115, -1.144, 0.093], [-0.117, -1.143, 0.096],
    [0.000, 0.288, -0.015], [0.082, 0.194, -0.006], [-0.077, 0.192, -0.011], [0.005, 0.353, 0.037],
    [0.173, 0.225, -0.015], [-0.176, 0.226, -0.020], [0.433, 0.213, -0.042], [-0.429, 0.212, -0.043],
    [0.683, 0.222, -0.046], [-0.685, 0.220, -0.047], [0.767, 0.214, -0.056], [-0.768, 0.213, -0.057]],
    np.float64)


def make_synthetic_smpl(seed: int = 0, v_template: Optional[np.ndarray] = None, lbs_nnz: int = 4,
                        shape_scale: float = 0.01, pose_scale: float = 0.003) -> SmplHostModel:
    """Deterministic SMPL-shaped model: real template geometry, standard
    tree, sparse J_regressor rows summing to 1, LBS weights with <= ``lbs_nnz`` non-zeros per vertex
    summing to 1, Gaussian blend shapes.  Used wherever the real pkl (absent) would be."""
    rng = np.random.default_rng(seed)
    if v_template is None:
        v_template = golden_fixtures()["v_template"].astype(np.float64).copy()
        v_template[:, 1] -= 0.18           # PLY origin -> SMPL-like origin (pelvis near y=-0.22)
    v_template = np.asarray(v_template, np.float64)
    V = v_template.shape[0]
    joints = _APPROX_JOINTS.copy()
    joints[:, 1] += 0.07
    d2 = ((v_template[:, None, :] - joints[None, :, :]) ** 2).sum(-1)          # (V,24)
    # J_regressor: each joint = convex combination of its 24 nearest vertices
    J_reg = np.zeros((V, NUM_JOINTS))
    for j in range(NUM_JOINTS):
        nn = np.argsort(d2[:, j])[:24]
        w = rng.uniform(0.2, 1.0, size=nn.shape)
        J_reg[nn, j] = w / w.sum()
    # LBS weights: softmin over joint distances, keep the lbs_nnz largest (1..lbs_nnz non-zeros)
    logits = -d2 / (2 * 0.06 ** 2)
    logits -= logits.max(1, keepdims=True)
    w = np.exp(logits)
    order = np.argsort(-w, axis=1)
    W = np.zeros((V, NUM_JOINTS))
    rows = np.arange(V)[:, None]
    keep = order[:, :lbs_nnz]
    W[rows, keep] = w[rows, keep]
    W[W < 1e-3 * W.max(1, keepdims=True)] = 0.0
    W /= W.sum(1, keepdims=True)
    shapedirs = rng.standard_normal((NUM_BETAS, V * 3)) * shape_scale
    # make shape directions spatially smooth-ish (scale with distance from pelvis) so beta changes limb length
    radial = (v_template - joints[0]).reshape(-1)
    shapedirs[0] = 0.04 * radial
    shapedirs[1] = 0.03 * np.tile(np.array([1.0, 0.0, 1.0]), V) * radial
    posedirs = rng.standard_normal((NUM_POSE_BASIS, V * 3)) * pose_scale
    # cocoplus regressor: 19 keypoints, each a convex combination of 12 vertices near a joint / extremity
    coco = np.zeros((V, 19))
    anchor = [8, 5, 2, 1, 4, 7, 21, 19, 17, 16, 18, 20, 12, 15, 15, 15, 15, 15, 15]
    for k, j in enumerate(anchor):
        c = joints[j] + (rng.standard_normal(3) * 0.03 if k >= 14 else 0.0)
        nn = np.argsort(((v_template - c) ** 2).sum(1))[:12]
        ww = rng.uniform(0.2, 1.0, size=12)
        coco[nn, k] = ww / ww.sum()
    f32 = np.float32
    return SmplHostModel(np.ascontiguousarray(v_template, f32), np.ascontiguousarray(shapedirs, f32),
                         np.ascontiguousarray(posedirs, f32), np.ascontiguousarray(J_reg, f32),
                         SMPL_PARENTS.copy(), np.ascontiguousarray(W, f32), np.ascontiguousarray(coco, f32),
                         meta={"source": "synthetic", "seed": seed}).validate()


# ----------------------------------------------------------------------------------------------
# part tables
# ----------------------------------------------------------------------------------------------
def part_vertices_filename(vertex_sampling: Optional[int]) -> str:
    """projects_to_seg.py:18-21 (CWD-relative literals in the reference)."""
    if vertex_sampling is None:
        return "./keras_smpl/part_vertices.pkl"
    return "./keras_smpl/" + str(vertex_sampling) + "_sampled_part_vertices.pkl"


def load_part_vertices(path: str) -> List[List[int]]:
    with open(path, "rb") as f:
        parts = pickle.load(f)
    return [list(map(int, p)) for p in parts]


def golden_part_vertices(vertex_sampling: Optional[int]) -> List[List[int]]:
    fx = golden_fixtures()
    vs = 1 if vertex_sampling is None else int(vertex_sampling)
    ptr, idx = fx["parts%d_ptr" % vs], fx["parts%d_idx" % vs]
    return [idx[ptr[k]:ptr[k + 1]].tolist() for k in range(len(ptr) - 1)]


def write_part_vertices_pkl(parts: Sequence[Sequence[int]], path: str, protocol: int = 0) -> None:
    with open(path, "wb") as f:
        pickle.dump([list(map(int, p)) for p in parts], f, protocol=protocol)


def sampled_part_table(parts: Sequence[Sequence[int]], vertex_sampling: Optional[int]):
    """CSR (ptr, idx) of *sampled-space* vertex indices per part: ``index // vertex_sampling``
    (projects_to_seg.py:36-37)."""
    vs = 1 if vertex_sampling is None else int(vertex_sampling)
    ptr = np.zeros(len(parts) + 1, np.int32)
    ptr[1:] = np.cumsum([len(p) for p in parts])
    idx = np.concatenate([np.asarray(p, np.int64) // vs for p in parts]).astype(np.int32) if ptr[-1] else \
        np.zeros(0, np.int32)
    return ptr, idx


# ----------------------------------------------------------------------------------------------
# mean parameters
# ----------------------------------------------------------------------------------------------
_H5_SIG = b"\x89HDF\r\n\x1a\n"
_H5_SHAPE_OFF, _H5_POSE_OFF, _H5_LEN = 4192, 4272, 4848


def read_mean_params_h5(path_or_bytes) -> dict:
    """Read ``pose`` (72 f8) and ``shape`` (10 f8) from neutral_smpl_mean_params.h5 without h5py/deepdish.

    The file the reference ships is a 4848-byte PyTables file whose two contiguous float64 datasets sit
    at fixed byte offsets (shape @4192, pose @4272 = file end - 82*8).  Any other layout is rejected
    loudly rather than guessed."""
    if isinstance(path_or_bytes, (bytes, bytearray, np.ndarray)):
        raw = bytes(bytearray(np.asarray(path_or_bytes, np.uint8)) if isinstance(path_or_bytes, np.ndarray)
                    else path_or_bytes)
    else:
        with open(path_or_bytes, "rb") as f:
            raw = f.read()
    if raw[:8] != _H5_SIG:
        raise ValueError("not an HDF5 file (bad signature)")
    if len(raw) != _H5_LEN or b"shape" not in raw or b"pose" not in raw:
        raise ValueError("unsupported mean-params h5 layout (expected the 4848-byte file shipped with the "
                         "reference); convert it to npz with keys 'pose','shape' instead")
    shape = np.frombuffer(raw, "<f8", count=10, offset=_H5_SHAPE_OFF).copy()
    pose = np.frombuffer(raw, "<f8", count=72, offset=_H5_POSE_OFF).copy()
    if not (np.all(np.isfinite(shape)) and np.all(np.isfinite(pose)) and np.abs(pose).max() < 10):
        raise ValueError("mean-params h5 payload failed sanity checks")
    return {"pose": pose, "shape": shape}


def load_mean_params(path: Optional[str] = None) -> dict:
    """``dd.io.load('./neutral_smpl_mean_params.h5')`` equivalent (concat_mean_param.py:9-10).  Accepts the
    h5 itself or an npz with the same two keys; with ``path=None`` tries the reference's CWD-relative
    literal and then the repo's golden copy of the same bytes."""
    if path is None:
        path = "./neutral_smpl_mean_params.h5"
        if not os.path.exists(path):
            return read_mean_params_h5(golden_fixtures()["h5_bytes"])
    if path.endswith(".npz"):
        with np.load(path) as z:
            return {"pose": z["pose"].astype(np.float64), "shape": z["shape"].astype(np.float64)}
    return read_mean_params_h5(path)


def mean_param_vector(img_wh: float, mean: Optional[dict] = None) -> np.ndarray:
    """The (1,86) initial/mean vector: cam [wh/2, wh/2, wh/2, wh/1.6], pose with global rotation zeroed,
    shape (concat_mean_param.py:12-25; set_cam_params.py:29-47).  float64 like the reference, cast later."""
    mean = mean or load_mean_params()
    pose = np.array(mean["pose"], np.float64)
    pose[:3] = 0.0
    out = np.zeros((1, 86))
    out[0, 4:] = np.hstack((pose, np.asarray(mean["shape"], np.float64)))
    out[0, 0] = img_wh / 2.0
    out[0, 1] = img_wh / 2.0
    out[0, 2] = img_wh / 2.0
    out[0, 3] = img_wh / 1.6
    return out
