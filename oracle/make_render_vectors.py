#!/usr/bin/env python
"""ORACLE TOOLING -- TEST INFRASTRUCTURE ONLY.

Runs the reference's OWN, UNMODIFIED renderer.py (SMPLRenderer.__call__ / rotated -> render_model -> simple_renderer) on
seeded meshes and writes tests/golden/render_vectors.npz.  `opendr`, `cv2` and `plyfile` do not exist in this image; the
import names resolve to the stand-ins of oracle/tf_shim/ (see oracle/tf_shim/opendr/__init__.py for what that pins and
what it cannot).  renderer.py is python-2 code: `colors.values()[color_id % ...]` (:241-244) does not run on python 3, so
every call passes color_id=None, which renderer.py itself maps to 'light_blue' (:239-240).

Run from the repository root IN THE BUILD CONTAINER (the GPU box has no /root/reference):

    python oracle/make_render_vectors.py

Meshes (camera frame: x right, y down, z forward, shifted in front of the camera as renderer.get_original does, :268-271):
  tmpl   the T-pose template of template-bodyparts.ply
  posed  the synthetic SMPL model (smpl_io.make_synthetic_smpl(seed=0)) decoded at seeded random parameters
Cases: lit / render_seg / background image with alpha / white background with alpha / rotated 60 degrees about y /
explicit camera, non-square image and clip planes that cut the mesh.
"""
import os
import sys

sys.dont_write_bytecode = True      # /root/reference is read-only and must stay untouched

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
SHIM = os.path.join(ROOT, "oracle", "tf_shim")
OUT = os.path.join(ROOT, "tests", "golden", "render_vectors.npz")


def camera_frame(verts, tz):
    v = np.array(verts, np.float64)
    v[:, 1] *= -1.0
    v[:, 2] *= -1.0
    return (v + np.array([0.0, 0.0, tz])).astype(np.float32)


def main():
    if not os.path.isfile(os.path.join(REF, "renderer.py")):
        raise SystemExit("reference tree not found at %s" % REF)
    sys.path.insert(0, ROOT)
    import importlib
    pkg = importlib.import_module("indirect_learning_pose-shape_b200")       # inputs only: fixtures, synthetic model, params
    synth = importlib.import_module("indirect_learning_pose-shape_b200.synth")
    from oracle import np_oracle
    fx = pkg.smpl_io.golden_fixtures()
    host = pkg.smpl_io.make_synthetic_smpl(seed=0)
    posed = np_oracle.smpl_layer_call(host, synth.make_params(1, 48, seed=11))[0]

    sys.path.insert(0, REF)
    sys.path.insert(0, SHIM)
    os.chdir(REF)                      # renderer.py opens 'keras_smpl/smpl_faces.npy' and 'template-bodyparts.ply'
    import renderer as ref_renderer
    import opendr
    assert ref_renderer.__file__.startswith(REF) and opendr.__file__.startswith(SHIM)

    rng = np.random.RandomState(5)
    out = {}
    meshes = {"tmpl": camera_frame(fx["v_template"], 6.0), "posed": camera_frame(posed, 5.0)}
    R = ref_renderer.SMPLRenderer(img_size=96, flength=230.)
    bg = rng.randint(0, 256, size=(96, 96, 3)).astype(np.uint8)
    out["background"] = bg
    for name, v in meshes.items():
        out[name + "_verts"] = v
        out[name + "_lit"] = R(v, color_id=None)
        out[name + "_seg"] = R(v, color_id=None, render_seg=True)
        out[name + "_bg_alpha"] = R(v, color_id=None, img=bg, do_alpha=True)
        out[name + "_alpha"] = R(v, color_id=None, do_alpha=True)
        out[name + "_rot60"] = R.rotated(v, 60, color_id=None)
        out[name + "_cam"] = R(v, cam=[260.0, 50.5, 44.25], color_id=None, img_size=(80, 112), near=1.0,
                               far=float(v[:, 2].mean()))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == "__main__":
    main()
