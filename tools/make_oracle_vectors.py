#!/usr/bin/env python
"""Write tests/golden/oracle_vectors.npz: outputs of the ORACLE (oracle/np_oracle.py) on the shipped mean parameters
and two seeded random samples, with the seeded synthetic SMPL model.  The reference itself cannot be run (python 2.7 /
TF 1.x / the SMPL pickle are absent), so these vectors do not pin parity to the reference; they only guard the oracle
and the synthetic model generator against drift."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import np_oracle  # noqa: E402

pkg = importlib.import_module("indirect_learning_pose-shape_b200")
synth = importlib.import_module("indirect_learning_pose-shape_b200.synth")
host = pkg.smpl_io.make_synthetic_smpl(seed=0)
parts = pkg.smpl_io.golden_part_vertices(5)
p = np.concatenate([pkg.smpl_io.mean_param_vector(48).astype(np.float32), synth.make_params(2, 48, seed=2024)], 0)
out = np_oracle.decode(host, p, 48, 5, parts, silhouette_wh=48)
dst = os.path.join(ROOT, "tests", "golden", "oracle_vectors.npz")
os.makedirs(os.path.dirname(dst), exist_ok=True)
np.savez_compressed(dst, params=p, verts_sub=out["verts"][:, ::53], J_transformed=out["J_transformed"],
                    projects=out["projects"], mask=out["mask"], seg_labels=out["seg"].argmax(-1).astype(np.uint8),
                    sil=out["silhouette"][..., 1].astype(np.float32))
print("wrote", dst, os.path.getsize(dst), "bytes")
