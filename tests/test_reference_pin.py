"""Pins the oracle to the REFERENCE'S OWN SOURCE.

tests/golden/reference_vectors.npz was written by oracle/make_reference_vectors.py, which imports the unmodified files
/root/reference/keras_smpl/{batch_smpl,projection,compute_mask,projects_to_seg,projects_to_silhouette,concat_mean_param,
set_cam_params}.py and focal_loss.py and executes them on seeded inputs, with `tensorflow` / `keras` resolved to the
torch-CPU stand-ins of oracle/tf_shim/ (TF and Keras are absent from the image).  So every Python statement of the
reference path ran as written; only the float kernels underneath are torch's.  These tests hold the NumPy restatement
(oracle/np_oracle.py) and its autograd twin (oracle/torch_oracle.py) to those vectors, check the stand-in ops against
TensorFlow's documented semantics, and -- in the build container, where /root/reference exists -- regenerate the
vectors and require them to be identical to the committed file.
"""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import np_oracle, torch_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")


@pytest.fixture(scope="module")
def ref():
    return np.load(GOLDEN)


@pytest.mark.parametrize("tag,vs,wh", [("c5", 5, 48), ("c1", None, 48), ("v2", 2, 64)])
def test_oracle_forward_matches_reference_source(ref, host_model, parts_by_vs, tag, vs, wh):
    p = ref[tag + "_params"]
    o = np_oracle.decode(host_model, p, wh, vs, parts_by_vs[vs])
    # geometry: same op order, different BLAS summation order only
    assert np.abs(o["verts"] - ref[tag + "_verts"]).max() <= 1e-6
    assert np.abs(o["J_transformed"] - ref[tag + "_J_transformed"]).max() <= 1e-6
    assert np.abs(o["projects"] - ref[tag + "_projects"]).max() <= 1e-5       # |u| ~ 24..48: 1e-5 = 2.6 ulp
    # integer-valued / discrete outputs: identical
    assert np.array_equal(o["mask"], ref[tag + "_mask"])
    assert np.array_equal(o["seg"].argmax(-1), ref[tag + "_seg"].argmax(-1))
    # rasterisers on the reference's own projections and mask: rounding of exp/sqrt only
    assert np.array_equal(np_oracle.compute_mask(ref[tag + "_projects"]), ref[tag + "_mask"])
    assert np.array_equal(np_oracle.compute_mask(ref[tag + "_projects"], fast=False), ref[tag + "_mask"])
    seg = np_oracle.projects_to_seg([ref[tag + "_projects"], ref[tag + "_mask"]], wh, vs, parts_by_vs[vs])
    assert np.abs(seg - ref[tag + "_seg"]).max() <= 5e-7
    assert np.abs(o["seg"] - ref[tag + "_seg"]).max() <= 2e-5                   # |ds| <= |dd| ~ projections' 1e-5


def test_c1_is_the_shipped_mean_params(ref, pkg):
    """Config C1 decodes exactly load_mean_set_cam_params(0): the reference's h5 mean pose/shape + camera init."""
    assert np.array_equal(ref["c1_params"], pkg.smpl_io.mean_param_vector(48).astype(np.float32))
    assert np.array_equal(ref["c1_params"][0, :4], np.float32([24, 24, 24, 30]))
    assert np.all(ref["c1_params"][0, 4:7] == 0)


def test_mean_and_camera_params_bit_exact(ref, pkg):
    mv = pkg.smpl_io.load_mean_params()
    feats = np.arange(21, dtype=np.float32).reshape(3, 7)
    for w in (48, 64):
        assert np.array_equal(np_oracle.concat_mean_param(feats, w, mv), ref["a1_concat_%d" % w])
        assert np.array_equal(np_oracle.set_cam_params(np.full((3, 86), 0.25, np.float32), w), ref["a1_setcam_%d" % w])
        assert np.array_equal(np_oracle.load_mean_set_cam_params(np.zeros((3, 86), np.float32), w, mv),
                              ref["a1_loadmean_%d" % w])


def test_silhouette_matches_reference_source(ref):
    sil = np_oracle.projects_to_silhouette(ref["sil_projects"], 48)
    assert np.abs(sil - ref["sil_out"]).max() <= 5e-7
    x = torch.tensor(ref["sil_projects"], dtype=torch.float64, requires_grad=True)
    (torch_oracle.projects_to_silhouette(x, 48) * torch.tensor(ref["sil_G"], dtype=torch.float64)).sum().backward()
    g, r = x.grad.numpy(), ref["sil_g_projects"].astype(np.float64)
    bad = np.abs(g - r) > 2e-4 * np.abs(r).max()
    assert bad.mean() <= 1e-3, (bad.mean(), np.abs(g - r).max())


@pytest.mark.parametrize("tag,vs,wh", [("c5", 5, 48), ("v2", 2, 64)])
def test_oracle_gradient_matches_reference_autograd(ref, host_model, parts_by_vs, tag, vs, wh):
    """d sum(seg * G) / d params: torch autograd through the reference's own statements vs the oracle's torch twin."""
    C = torch_oracle.TorchSmplConstants(host_model, torch.float32)
    x = torch.tensor(ref[tag + "_params"], requires_grad=True)
    o = torch_oracle.decode(C, x, wh, vs, parts_by_vs[vs])
    (o["seg"] * torch.tensor(ref[tag + "_G"])).sum().backward()
    got, want = x.grad.numpy().astype(np.float64), ref[tag + "_g_params"].astype(np.float64)
    scale = np.abs(want).max(axis=0, keepdims=True) + 1e-6
    err = np.abs(got - want) / scale
    assert np.median(err) <= 1e-4 and err.max() <= 2e-2, (np.median(err), err.max())


def test_focal_loss_matches_reference_source(ref):
    seg = ref["c5_seg"].reshape(2, 48 * 48, 32)
    y = np.eye(32, dtype=np.float32)[ref["c5_labels"]]
    for weighted in (False, True):
        got = np_oracle.categorical_focal_loss(y, np_oracle.softmax_last_axis(seg), 2.0, weighted)
        want = ref["c5_focal%d" % weighted]
        assert np.abs(got - want).max() <= 2e-6 * max(1.0, np.abs(want).max())


def test_shim_ops_follow_tensorflow_semantics():
    """The stand-in ops the vectors depend on, against TF's documented behaviour."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
    try:
        saved = {k: sys.modules.pop(k, None) for k in ("tensorflow", "keras", "keras.backend", "keras.layers")}
        tf = importlib.import_module("tensorflow")
        assert tf.__file__.startswith(os.path.join(ROOT, "oracle", "tf_shim"))
        T = tf.Tensor
        # round half to even
        assert tf.round(T(torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5]))).numpy().tolist() == [0.0, 2.0, 2.0, -0.0, -2.0]
        # argmax: smallest index among equal maxima
        assert int(tf.argmax(T(torch.tensor([[1.0], [3.0], [3.0], [2.0]])), axis=0).numpy()[0]) == 1
        # reduce_max: gradient split evenly among ties
        x = torch.tensor([[1.0, 2.0, 2.0]], requires_grad=True)
        tf.reduce_max(T(x), axis=1).t.sum().backward()
        assert x.grad.tolist() == [[0.0, 0.5, 0.5]]
        # clip_by_value: gradient passes on the closed interval
        x = torch.tensor([-0.5, 0.0, 0.5, 1.0, 1.5], requires_grad=True)
        tf.clip_by_value(T(x), 0, 1).t.sum().backward()
        assert x.grad.tolist() == [0.0, 1.0, 1.0, 1.0, 0.0]
        # gather with (?,1) indices keeps the index shape; where returns row-major coordinates
        g = tf.gather(T(torch.arange(12.0).reshape(4, 3)), tf.where(T(torch.tensor([False, True, False, True]))))
        assert list(g.t.shape) == [2, 1, 3] and g.numpy()[:, 0, 0].tolist() == [3.0, 9.0]
        # meshgrid 'xy': t1 varies along columns
        a, b = tf.meshgrid(tf.range(0, 3), tf.range(0, 2))
        assert a.numpy().tolist() == [[0, 1, 2], [0, 1, 2]] and b.numpy().tolist() == [[0, 0, 0], [1, 1, 1]]
        # scatter_nd accumulates, pad / reverse / tile layouts
        s = tf.scatter_nd(T(torch.tensor([[1], [3], [1]])), T(torch.tensor([1.0, 2.0, 4.0])), [5])
        assert s.numpy().tolist() == [0.0, 5.0, 0.0, 2.0, 0.0]
        assert tf.pad(T(torch.ones(1, 2, 2)), [[0, 0], [0, 1], [3, 0]]).numpy().shape == (1, 3, 5)
        assert tf.reverse(T(torch.arange(6.0).reshape(1, 3, 2)), axis=[1]).numpy()[0, 0].tolist() == [4.0, 5.0]
        # norm = sqrt(sum(x*x)), fp32
        v = torch.tensor([[3.0, 4.0]])
        assert float(tf.norm(T(v), axis=1).numpy()[0]) == 5.0
    finally:
        sys.path.pop(0)
        for k in ("tensorflow", "keras", "keras.backend", "keras.layers"):
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


@pytest.mark.skipif(not os.path.isdir("/root/reference/keras_smpl"), reason="reference tree only exists in the build container")
def test_vectors_regenerate_identically(tmp_path):
    """Re-run the reference's source and require the committed fixture to be what it produces (drift guard)."""
    out = tmp_path / "regen.npz"
    code = ("import sys; sys.argv=['x']; import importlib.util as u; "
            "s=u.spec_from_file_location('mrv', %r); m=u.module_from_spec(s); s.loader.exec_module(m); "
            "m.OUT=%r; m.main()" % (os.path.join(ROOT, "oracle", "make_reference_vectors.py"), str(out)))
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT, capture_output=True, timeout=600)
    a, b = np.load(GOLDEN), np.load(str(out))
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
        assert a[k].shape == b[k].shape, k
        if a[k].dtype.kind == "f":
            # same machine -> identical; another BLAS thread count may reorder sums: allow rounding noise only
            assert np.abs(a[k].astype(np.float64) - b[k]).max() <= 1e-5 * max(1.0, np.abs(a[k]).max()), k
        else:
            assert np.array_equal(a[k], b[k]), k
