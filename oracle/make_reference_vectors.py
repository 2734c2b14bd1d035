#!/usr/bin/env python
"""ORACLE TOOLING — TEST INFRASTRUCTURE ONLY.

Runs the reference's OWN, UNMODIFIED source files (/root/reference/keras_smpl/*.py and focal_loss.py) on seeded inputs
and writes tests/golden/reference_vectors.npz.  TensorFlow/Keras/deepdish/cPickle/chumpy do not exist in this image, so
the import names resolve to oracle/tf_shim/ (torch-CPU implementations of exactly the TF/Keras symbols those files
call, with TF's documented semantics; see oracle/tf_shim/README.md).  Every Python statement of the reference's decoder
path is therefore executed as written; only the floating-point kernels underneath are torch's instead of TensorFlow's.

Run from the repository root IN THE BUILD CONTAINER (the GPU box has no /root/reference):

    python oracle/make_reference_vectors.py

Cases (model = smpl_io.make_synthetic_smpl(seed=0) written as an HMR-layout pickle and loaded by SMPLLayer.build itself):
  c5   N=2  wh=48 vertex_sampling=5   SMPLLayer -> orthographic_project -> compute_mask -> projects_to_seg, with
            d(sum(seg*G))/d(params) by autograd through the reference's code, and softmax -> categorical_focal_loss
  v2   N=1  wh=64 vertex_sampling=2 (2_sampled_part_vertices.pkl), the same chain with its gradient
  c1   N=1  wh=48 vertex_sampling=None, params = load_mean_set_cam_params(zeros) (the shipped mean params; config C1)
  sil  N=1  wh=48 projects_to_silhouette on the c1 projections (the reference hard-codes 6890 vertices) + gradient
  a1   concat_mean_param / set_cam_params / load_mean_set_cam_params outputs at wh = 48 and 64
"""
import os
import sys
import tempfile

sys.dont_write_bytecode = True      # /root/reference is read-only and must stay untouched

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
SHIM = os.path.join(ROOT, "oracle", "tf_shim")
OUT = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")


def main():
    if not os.path.isdir(os.path.join(REF, "keras_smpl")):
        raise SystemExit("reference tree not found at %s" % REF)
    sys.path.insert(0, ROOT)
    import importlib
    pkg = importlib.import_module("indirect_learning_pose-shape_b200")       # inputs only: synthetic model + seeded params
    synth = importlib.import_module("indirect_learning_pose-shape_b200.synth")
    host = pkg.smpl_io.make_synthetic_smpl(seed=0)
    tmp = tempfile.mkdtemp(prefix="refvec_")
    pkl = os.path.join(tmp, "neutral_smpl_with_cocoplus_reg.pkl")
    pkg.smpl_io.save_smpl_pkl(host, pkl)

    sys.path.insert(0, REF)
    sys.path.insert(0, SHIM)
    os.chdir(REF)                      # the reference opens './keras_smpl/*.pkl' and './neutral_smpl_mean_params.h5'
    import tensorflow as tf            # the shim
    from keras_smpl.batch_smpl import SMPLLayer
    from keras_smpl.projection import orthographic_project
    from keras_smpl.compute_mask import compute_mask
    from keras_smpl.projects_to_seg import projects_to_seg
    from keras_smpl.projects_to_silhouette import projects_to_silhouette
    from keras_smpl.concat_mean_param import concat_mean_param
    from keras_smpl.set_cam_params import set_cam_params, load_mean_set_cam_params
    from focal_loss import categorical_focal_loss
    assert tf.__file__.startswith(SHIM)
    for mod in ("keras_smpl.batch_smpl", "keras_smpl.projects_to_seg", "keras_smpl.compute_mask", "focal_loss"):
        assert sys.modules[mod].__file__.startswith(REF), sys.modules[mod].__file__

    out = {}
    torch.set_num_threads(os.cpu_count() or 1)

    def run_decoder(params_np, wh, vs, tag, with_grad):
        n = params_np.shape[0]
        x = torch.tensor(params_np, requires_grad=with_grad)
        layer = SMPLLayer(pkl, batch_size=n)
        layer.build(None)
        verts = layer.call(tf.Tensor(x))
        pwd = orthographic_project([verts, tf.Tensor(x)], vs)
        mask = compute_mask(pwd)
        seg = projects_to_seg([pwd, mask], wh, vs)
        out[tag + "_params"] = params_np
        out[tag + "_verts"] = verts.numpy()
        out[tag + "_J_transformed"] = layer.J_transformed.numpy()
        out[tag + "_projects"] = pwd.numpy()
        out[tag + "_mask"] = mask.numpy()
        out[tag + "_seg"] = seg.numpy()
        return x, verts, pwd, mask, seg

    # ---- c5 ----------------------------------------------------------------------------------------------------
    wh, vs, n = 48, 5, 2
    p = synth.make_params(n, wh, seed=2024)
    x, verts, pwd, mask, seg = run_decoder(p, wh, vs, "c5", True)
    G = torch.randn(seg.t.shape, generator=torch.Generator().manual_seed(7))
    (seg.t * G).sum().backward(retain_graph=True)
    out["c5_G"] = G.numpy()
    out["c5_g_params"] = x.grad.numpy().copy()
    # the op after the path: Reshape -> softmax (model.py:119-120) -> categorical focal loss (focal_loss.py)
    x.grad = None
    labels = torch.randint(0, 32, (n, wh * wh), generator=torch.Generator().manual_seed(8))
    y_true = torch.nn.functional.one_hot(labels, 32).float()
    y_pred = torch.softmax(seg.t.reshape(n, wh * wh, 32), dim=-1)          # Keras Activation('softmax')
    for weighted in (False, True):
        loss = categorical_focal_loss(gamma=2.0, weight_classes=weighted)(tf.Tensor(y_true), tf.Tensor(y_pred))
        out["c5_focal%d" % weighted] = loss.numpy()
    loss.t.sum().backward()
    out["c5_labels"] = labels.numpy().astype(np.uint8)
    out["c5_focal1_g_params"] = x.grad.numpy().copy()

    # ---- v2: the third part table, another resolution ----------------------------------------------------------------
    p2 = synth.make_params(1, 64, seed=2025)
    x2, _, _, _, seg2 = run_decoder(p2, 64, 2, "v2", True)
    G2 = torch.randn(seg2.t.shape, generator=torch.Generator().manual_seed(10))
    (seg2.t * G2).sum().backward()
    out["v2_G"] = G2.numpy()
    out["v2_g_params"] = x2.grad.numpy().copy()

    # ---- a1 + c1 -----------------------------------------------------------------------------------------------
    for w in (48, 64):
        z = tf.Tensor(torch.zeros(3, 86))
        f = tf.Tensor(torch.arange(3 * 7, dtype=torch.float32).reshape(3, 7))
        out["a1_concat_%d" % w] = concat_mean_param(f, w).numpy()
        out["a1_setcam_%d" % w] = set_cam_params(tf.Tensor(torch.ones(3, 86) * 0.25), w).numpy()
        out["a1_loadmean_%d" % w] = load_mean_set_cam_params(z, w).numpy()
    p1 = out["a1_loadmean_48"][:1].copy()
    x1, verts1, pwd1, mask1, seg1 = run_decoder(p1, 48, None, "c1", False)

    # ---- silhouette (6890 vertices hard-coded, projects_to_silhouette.py:33) -------------------------------------------
    pw = torch.tensor(pwd1.numpy(), requires_grad=True)
    sil = projects_to_silhouette(tf.Tensor(pw), 48)
    Gs = torch.randn(sil.t.shape, generator=torch.Generator().manual_seed(9))
    (sil.t * Gs).sum().backward()
    out["sil_projects"] = pwd1.numpy()
    out["sil_out"] = sil.numpy()
    out["sil_G"] = Gs.numpy()
    out["sil_g_projects"] = pw.grad.numpy().copy()

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print("wrote %s (%.1f KB): %s" % (OUT, os.path.getsize(OUT) / 1e3, ", ".join(sorted(out))))


if __name__ == "__main__":
    main()
