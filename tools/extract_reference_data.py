#!/usr/bin/env python
"""Extract the reference's *data* fixtures into indirect_learning_pose-shape_b200/data/ref_fixtures.npz.

Run once in the build container (needs /root/reference, which does not exist on
the GPU box).  Only data is extracted, never source:

  * v_template  : the 6890 T-pose vertices of template-bodyparts.ply (real SMPL
                  topology/geometry; used as the synthetic model's template
                  because neutral_smpl_with_cocoplus_reg.pkl is not shipped,
                  see /root/reference/.MISSING_LARGE_BLOBS)
  * ply_rgb     : per-vertex colours of the same file (part colouring)
  * parts{1,2,5}_ptr/idx : CSR form of keras_smpl/part_vertices.pkl,
                  2_sampled_part_vertices.pkl, 5_sampled_part_vertices.pkl
                  (original vertex ids, as read at projects_to_seg.py:18-24)
  * mean_pose / mean_shape : the float64 datasets of neutral_smpl_mean_params.h5
                  (read at concat_mean_param.py:9-15), plus their byte offsets
  * h5_bytes    : the raw 4848-byte h5 file, so the offset reader can be tested
                  on the GPU box / without the reference tree
  * faces       : keras_smpl/smpl_faces.npy, the 13776 triangles of the SMPL mesh
                  (loaded at renderer.py:27; input of the mesh visualiser)
"""
import os
import pickle
import sys

import numpy as np

REF = os.environ.get("SMPL_REF_DIR", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "indirect_learning_pose-shape_b200", "data",
                   "ref_fixtures.npz")


def read_ply_vertices(path):
    raw = open(path, "rb").read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    hdr = raw[:end].decode("ascii")
    assert "element vertex 6890" in hdr and "binary_little_endian" in hdr
    dt = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                   ("r", "u1"), ("g", "u1"), ("b", "u1")])
    v = np.frombuffer(raw, dtype=dt, count=6890, offset=end)
    xyz = np.stack([v["x"], v["y"], v["z"]], 1).astype(np.float32)
    rgb = np.stack([v["r"], v["g"], v["b"]], 1).astype(np.uint8)
    return xyz, rgb


def csr(lists):
    ptr = np.zeros(len(lists) + 1, np.int32)
    ptr[1:] = np.cumsum([len(x) for x in lists])
    idx = np.concatenate([np.asarray(x, np.int32) for x in lists])
    return ptr, idx


def main():
    out = {}
    out["v_template"], out["ply_rgb"] = read_ply_vertices(os.path.join(REF, "template-bodyparts.ply"))
    for vs, name in ((1, "part_vertices.pkl"), (2, "2_sampled_part_vertices.pkl"), (5, "5_sampled_part_vertices.pkl")):
        with open(os.path.join(REF, "keras_smpl", name), "rb") as f:
            parts = pickle.load(f)
        assert len(parts) == 31
        out["parts%d_ptr" % vs], out["parts%d_idx" % vs] = csr(parts)
    raw = open(os.path.join(REF, "neutral_smpl_mean_params.h5"), "rb").read()
    assert len(raw) == 4848 and raw[:8] == b"\x89HDF\r\n\x1a\n"
    out["h5_bytes"] = np.frombuffer(raw, np.uint8)
    out["mean_shape"] = np.frombuffer(raw, "<f8", count=10, offset=4192).copy()
    out["mean_pose"] = np.frombuffer(raw, "<f8", count=72, offset=4272).copy()
    faces = np.load(os.path.join(REF, "keras_smpl", "smpl_faces.npy"))
    assert faces.shape == (13776, 3) and faces.max() == 6889
    out["faces"] = faces.astype(np.int32)
    np.savez_compressed(OUT, **out)
    print("wrote", os.path.normpath(OUT), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    sys.exit(main())
