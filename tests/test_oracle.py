"""Oracle self-consistency (CPU).  The reference has no tests or golden vectors and cannot run here (python2/TF1
absent, SMPL pickle not shipped): PARITY UNPINNED.  What can be pinned is pinned: the oracle against its fp64 twin,
its torch twin, finite differences, closed-form properties, the reference's shipped data files, and drift vectors.
"""
import os

import numpy as np
import pytest
import torch

from oracle import np_oracle, torch_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_vectors.npz")


def test_fp32_vs_fp64(host_model, make_params):
    p = make_params(4, 48, seed=1)
    a = np_oracle.smpl_layer_call(host_model, p, return_all=True)
    b = np_oracle.smpl_layer_call(host_model, p.astype(np.float64), return_all=True)
    assert a["verts"].dtype == np.float32 and b["verts"].dtype == np.float64
    for k in ("verts", "J_transformed", "joints", "v_posed"):
        assert np.abs(a[k] - b[k]).max() < 5e-6, k


def test_zero_pose_gives_shape_blend(host_model):
    p = np.zeros((2, 86), np.float64)
    p[:, 76:] = [[1, -2, 0.5, 0, 0, 0, 0, 0, 0, 3], [0] * 10]
    out = np_oracle.smpl_layer_call(host_model, p, return_all=True)
    assert np.abs(out["verts"] - out["v_shaped"]).max() < 1e-7      # 1e-8 shift inside the Rodrigues norm
    assert np.abs(out["verts"][1] - host_model.v_template).max() < 1e-7
    assert np.abs(out["J_transformed"] - out["J"]).max() < 1e-7


def test_root_rotation_is_rigid(host_model, make_params):
    """A pure global rotation rotates the shaped mesh about the root joint (batch_smpl.py:192,204)."""
    p = np.zeros((1, 86), np.float64)
    p[0, 4:7] = [0.3, -0.7, 0.2]
    out = np_oracle.smpl_layer_call(host_model, p, return_all=True)
    R = np_oracle.batch_rodrigues(p[:, 4:7])[0]
    J0 = out["J"][0, 0]
    expect = (out["v_shaped"][0] - J0) @ R.T + J0
    assert np.abs(out["verts"][0] - expect).max() < 1e-6


def test_rodrigues_orthonormal():
    th = np.random.default_rng(0).standard_normal((50, 3))
    R = np_oracle.batch_rodrigues(th)
    assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 1e-6
    assert np.abs(np.linalg.det(R) - 1).max() < 1e-6
    R0 = np_oracle.batch_rodrigues(np.zeros((1, 3), np.float32))      # theta = 0: finite, identity (Q2)
    assert np.isfinite(R0).all() and np.abs(R0[0] - np.eye(3)).max() < 1e-6


def test_torch_twin_matches_numpy(host_model, parts_by_vs, make_params):
    p = make_params(2, 48, seed=2)
    a = np_oracle.decode(host_model, p, 48, 5, parts_by_vs[5], silhouette_wh=32)
    C = torch_oracle.TorchSmplConstants(host_model)
    b = torch_oracle.decode(C, torch.tensor(p), 48, 5, parts_by_vs[5], silhouette_wh=32)
    assert np.abs(a["verts"] - b["verts"].numpy()).max() < 2e-6
    assert np.abs(a["projects"] - b["projects"].numpy()).max() < 5e-5
    same = (a["mask"] == b["mask"].numpy()).all(1)
    assert same.any()
    assert np.abs(a["seg"][same] - b["seg"].numpy()[same]).max() < 1e-4
    sil_t = torch_oracle.projects_to_silhouette(torch.tensor(a["projects"]), 32).numpy()
    assert np.abs(a["silhouette"] - np_oracle.projects_to_silhouette(a["projects"], 32)).max() == 0
    assert np.abs(sil_t - np_oracle.projects_to_silhouette(a["projects"], 32)).max() < 2e-6


def test_mask_fast_equals_literal(host_model, make_params):
    p = make_params(2, 48, seed=3)
    pr = np_oracle.orthographic_project([np_oracle.smpl_layer_call(host_model, p), p], 5)
    assert np.array_equal(np_oracle.compute_mask(pr, fast=True), np_oracle.compute_mask(pr, fast=False))
    rng = np.random.default_rng(0)
    pr = (rng.random((2, 400, 3)) * 80 - 8).astype(np.float32)       # profiling_renderer.py:26-style input
    pr[0, :30, 2] = 1.0
    pr[0, :30, :2] = [5.0, 6.0]
    assert np.array_equal(np_oracle.compute_mask(pr, fast=True), np_oracle.compute_mask(pr, fast=False))


def test_mask_semantics():
    pr = np.array([[[3.2, 4.4, 0.1], [3.4, 4.3, 0.9], [2.6, 3.6, 0.9], [10.5, 2.5, 0.0], [11.5, 2.5, 5.0],
                    [-3.0, 2.0, 9.0]]], np.float32)
    m = np_oracle.compute_mask(pr, fast=False)[0]
    # vertices 0,1,2 share pixel (3,4): the largest z wins, first index on the tie (compute_mask.py:100-102)
    assert m[0] == 500 and m[1] == 1 and m[2] == 500
    # same vertices re-ordered: the tie now resolves to the other one (index 1 is also the "empty pixel" vote, :98-99)
    m2 = np_oracle.compute_mask(pr[:, [0, 2, 1, 3, 4, 5]], fast=False)[0]
    assert m2[0] == 500 and m2[1] == 1 and m2[2] == 500
    # half to even: 10.5 -> 10, 11.5 -> 12, 2.5 -> 2 ; out-of-grid vertex is never visible
    assert m[3] == 1 and m[4] == 1 and m[5] == 500


def test_seg_properties(parts_by_vs):
    rng = np.random.default_rng(0)
    Vs, wh = 1378, 24
    pr = np.concatenate([rng.random((1, Vs, 2)) * wh, rng.standard_normal((1, Vs, 1))], 2).astype(np.float32)
    mask = np.ones((1, Vs), np.float32)
    seg = np_oracle.projects_to_seg([pr, mask], wh, 5, parts_by_vs[5])
    assert seg.shape == (1, wh, wh, 32)
    assert (seg >= 0).all() and (seg <= 1).all()
    # translation equivariance + vertical flip: shifting v by +1 row moves the picture one row UP after the flip
    pr2 = pr.copy()
    pr2[..., 1] += 1.0
    seg2 = np_oracle.projects_to_seg([pr2, mask], wh, 5, parts_by_vs[5])
    assert np.abs(seg2[:, :-1, :, 1:] - seg[:, 1:, :, 1:]).max() < 1e-5
    # invisible weights only sharpen: weight 500 everywhere gives a (nearly) empty picture
    seg3 = np_oracle.projects_to_seg([pr, mask * 500], wh, 5, parts_by_vs[5])
    assert seg3[..., 1:].sum() < seg[..., 1:].sum() * 0.05


def test_finite_difference_gradients(host_model, parts_by_vs):
    """torch twin autograd vs central differences (fp64) through decode -> project -> silhouette."""
    C = torch_oracle.TorchSmplConstants(host_model, torch.float64)
    from importlib import import_module
    sm = import_module("indirect_learning_pose-shape_b200.synth")
    p = sm.make_params(1, 16, seed=4).astype(np.float64)
    w = np.random.default_rng(1).standard_normal((1, 16, 16, 2))

    def f(x):
        o = torch_oracle.smpl_layer_call(C, x)
        pr = torch_oracle.orthographic_project([o, x], 5)
        return (torch_oracle.projects_to_silhouette(pr, 16) * torch.tensor(w)).sum()

    x = torch.tensor(p, requires_grad=True)
    f(x).backward()
    g = x.grad.numpy()[0]
    for i in (0, 2, 5, 20, 50, 77, 85):
        e = np.zeros_like(p)
        e[0, i] = 1e-6
        fd = (f(torch.tensor(p + e)).item() - f(torch.tensor(p - e)).item()) / 2e-6
        assert abs(fd - g[i]) <= 1e-4 * max(1.0, abs(g[i])), (i, fd, g[i])


def test_mean_param_functions(smpl_io):
    mean = smpl_io.load_mean_params()
    feats = np.zeros((3, 2048), np.float32)
    out = np_oracle.concat_mean_param(feats, 48, mean)
    assert out.shape == (3, 2134)
    m = out[0, 2048:]
    assert np.allclose(m[:4], [24, 24, 24, 30]) and np.all(m[4:7] == 0)      # pose[:3] = 0 (concat_mean_param.py:14)
    assert abs(m[76] - 0.20561) < 1e-5 and abs(m[77] - 0.335563) < 1e-5     # shipped h5 values
    z = np.zeros((2, 86), np.float32)
    assert np.array_equal(np_oracle.load_mean_set_cam_params(z, 48, mean)[0], m)
    s = np_oracle.set_cam_params(z, 64)
    assert np.allclose(s[0, :4], [32, 32, 32, 40]) and np.all(s[0, 4:] == 0)


def test_golden_vectors_no_drift(host_model, parts_by_vs, smpl_io):
    """tests/golden/oracle_vectors.npz was written by tools/make_oracle_vectors.py from THIS oracle (the reference
    cannot run): it guards the oracle and the synthetic model against silent drift, nothing more."""
    z = np.load(GOLDEN)
    p = z["params"]
    out = np_oracle.decode(host_model, p, 48, 5, parts_by_vs[5], silhouette_wh=48)
    assert np.abs(out["verts"][:, ::53] - z["verts_sub"]).max() < 1e-6
    assert np.abs(out["J_transformed"] - z["J_transformed"]).max() < 1e-6
    assert np.abs(out["projects"] - z["projects"]).max() < 2e-5
    assert (out["mask"] != z["mask"]).mean() < 2e-3
    assert (out["seg"].argmax(-1) != z["seg_labels"]).mean() < 2e-3
    assert np.abs(out["silhouette"][..., 1] - z["sil"]).max() < 1e-4


# ---------------------------------------------------------------------------------------------------------------
# softmax + categorical focal loss (the op after the path; focal_loss.py:10-48)
# ---------------------------------------------------------------------------------------------------------------
def _focal_inputs(n=2, npix=40, C=32, seed=11):
    rng = np.random.default_rng(seed)
    seg = rng.random((n, npix, C)).astype(np.float64)           # rasteriser scores live in [0, 1]
    lab = rng.integers(0, C, (n, npix))
    y = np.eye(C)[lab]
    return seg, y, lab


@pytest.mark.parametrize("weighted", [False, True])
def test_focal_loss_closed_form_and_twins(weighted):
    seg, y, lab = _focal_inputs()
    p = np_oracle.softmax_last_axis(seg)
    got = np_oracle.categorical_focal_loss(y, p, 2.0, weighted)
    # one-hot: only the true class survives the sum (focal_loss.py:45 comment)
    pt = np.take_along_axis(p, lab[..., None], axis=2)[..., 0]
    w = np_oracle.focal_class_weights(32, np.float64)[lab] if weighted else 1.0
    assert np.allclose(got, w * (1 - pt) ** 2 * -np.log(pt), rtol=1e-12, atol=0)
    tw = torch_oracle.softmax_focal_loss(torch.tensor(y), torch.tensor(seg), 2.0, weighted).numpy()
    assert np.allclose(tw, got, rtol=1e-12, atol=1e-15)
    f32 = np_oracle.categorical_focal_loss(y.astype(np.float32), np_oracle.softmax_last_axis(seg.astype(np.float32)), 2.0, weighted)
    assert f32.dtype == np.float32 and np.allclose(f32, got, rtol=2e-5, atol=1e-6)


def test_focal_loss_gradient_matches_finite_differences():
    seg, y, _ = _focal_inputs(n=1, npix=6, C=8, seed=5)
    x = torch.tensor(seg, requires_grad=True)
    wgt = torch.tensor(np.random.default_rng(1).standard_normal((1, 6)))
    (torch_oracle.softmax_focal_loss(torch.tensor(y), x, 2.0, True) * wgt).sum().backward()
    g = x.grad.numpy()
    eps = 1e-6
    for idx in [(0, 0, 0), (0, 3, 5), (0, 5, 7)]:
        sp, sm = seg.copy(), seg.copy()
        sp[idx] += eps
        sm[idx] -= eps
        fp = (np_oracle.categorical_focal_loss(y, np_oracle.softmax_last_axis(sp), 2.0, True) * wgt.numpy()).sum()
        fm = (np_oracle.categorical_focal_loss(y, np_oracle.softmax_last_axis(sm), 2.0, True) * wgt.numpy()).sum()
        assert abs((fp - fm) / (2 * eps) - g[idx]) <= 1e-6 * max(1.0, abs(g[idx]))
