"""The oracle against INDEPENDENT implementations of the same mathematics (CPU, no GPU).

The reference (py2 / TF1) cannot run here and ships no golden vectors, so the oracle is "parity unpinned" with respect to
TensorFlow's own rounding.  What can still be pinned is the mathematics, against code the oracle's author did not write:
  * Rodrigues' formula            -> scipy.spatial.transform.Rotation.from_rotvec
  * kinematic chain + skinning    -> the textbook SMPL form  verts = sum_j w_vj G_j(theta) G_j(0)^-1 [v_posed; 1]
                                     (Loper et al. 2015, eq. 2-4) with 4x4 homogeneous matrices and numpy.linalg.inv,
                                     instead of the reference's "subtract the rotated rest joint" shortcut
                                     (batch_smpl.py:222-226)
  * point z-buffer                -> a dictionary keyed by rounded pixel, written from the prose of compute_mask.py
  * weighted nearest-vertex maps  -> scipy.spatial.cKDTree queries per weight class (exact Euclidean NN in float64)
  * softmax + focal loss          -> scipy.special.log_softmax
"""
import numpy as np
import pytest
from scipy.spatial import cKDTree
from scipy.spatial.transform import Rotation
from scipy.special import log_softmax

from oracle import np_oracle


def test_rodrigues_matches_scipy():
    rng = np.random.default_rng(0)
    theta = rng.standard_normal((500, 3)) * 1.5
    theta[:5] *= 1e-3                                            # small angles; the 1e-8 inside the norm is harmless here
    got = np_oracle.batch_rodrigues(theta.astype(np.float64))
    ref = Rotation.from_rotvec(theta).as_matrix()
    assert np.abs(got - ref).max() <= 5e-8                      # the 1e-8 added to each component inside the norm (batch_smpl.py:265)


def test_chain_and_skinning_match_textbook_smpl(host_model, make_params):
    p = make_params(3, 48, seed=2).astype(np.float64)
    r = np_oracle.smpl_layer_call(host_model, p, return_all=True)
    parents = np.asarray(host_model.parents)
    W = np.asarray(host_model.lbs_weights, np.float64)
    for n in range(p.shape[0]):
        R = Rotation.from_rotvec(p[n, 4:76].reshape(24, 3)).as_matrix()
        J = r["J"][n]
        G = np.zeros((24, 4, 4))
        G0 = np.zeros((24, 4, 4))
        for j in range(24):
            local = np.eye(4)
            local[:3, :3] = R[j]
            local[:3, 3] = J[j] - (J[parents[j]] if j else 0.0)
            rest = np.eye(4)
            rest[:3, 3] = local[:3, 3]
            G[j] = local if j == 0 else G[parents[j]] @ local
            G0[j] = rest if j == 0 else G0[parents[j]] @ rest
        rel = np.stack([G[j] @ np.linalg.inv(G0[j]) for j in range(24)])      # textbook relative transforms
        vp = np.concatenate([r["v_posed"][n], np.ones((W.shape[0], 1))], 1)
        verts = np.einsum("vj,jab,vb->va", W, rel, vp)[:, :3]
        assert np.abs(G[:, :3, 3] - r["J_transformed"][n]).max() <= 1e-7
        assert np.abs(verts - r["verts"][n]).max() <= 1e-7


def test_mask_matches_dictionary_zbuffer(host_model, make_params):
    p = make_params(4, 48, seed=9)
    verts = np_oracle.smpl_layer_call(host_model, p)
    for vs in (None, 5):
        pr = np_oracle.orthographic_project([verts, p], vs)
        got = np_oracle.compute_mask(pr)
        for n in range(pr.shape[0]):
            best = {}
            for i, (u, v, z) in enumerate(pr[n]):
                key = (float(np.rint(u)), float(np.rint(v)))                   # round half to even, like tf.round
                if not (0 <= key[0] <= 63 and 0 <= key[1] <= 63):
                    continue
                if key not in best or z > best[key][0]:                        # largest z, first index on ties
                    best[key] = (z, i)
            ref = np.full(pr.shape[1], 500.0, np.float32)
            for _, i in best.values():
                ref[i] = 1.0
            if len(best) < 64 * 64:
                ref[1] = 1.0                                                    # empty pixels vote for index 1 (Q5)
            assert np.array_equal(got[n], ref)


@pytest.mark.parametrize("vs", [5, None])
def test_seg_matches_kdtree(host_model, parts_by_vs, make_params, vs):
    wh = 48
    p = make_params(2, wh, seed=4)
    verts = np_oracle.smpl_layer_call(host_model, p)
    pr = np_oracle.orthographic_project([verts, p], vs).astype(np.float64)
    mask = np_oracle.compute_mask(pr.astype(np.float32)).astype(np.float64)
    got = np_oracle.projects_to_seg([pr, mask], wh, vs, parts_by_vs[vs])
    cols, rows = np.meshgrid(np.arange(wh), np.arange(wh))
    grid = np.stack([cols.ravel(), rows.ravel()], 1).astype(np.float64)
    for n in range(pr.shape[0]):
        scores = np.zeros((wh * wh, 31))
        for k, part in enumerate(parts_by_vs[vs]):
            ids = np.asarray(part) // (vs or 1)
            for w in np.unique(mask[n, ids]):                                   # one exact NN query per weight class
                pts = pr[n, ids[mask[n, ids] == w], :2]
                d, _ = cKDTree(pts).query(grid)
                scores[:, k] = np.maximum(scores[:, k], np.exp(-d * w))
        bg = 1.0 - np.clip(scores.sum(1), 0, 1)
        ref = np.concatenate([bg[:, None], scores], 1).reshape(wh, wh, 32)[::-1]
        assert np.abs(got[n] - ref).max() <= 1e-12


def test_silhouette_matches_kdtree(host_model, make_params):
    wh = 64
    p = make_params(2, wh, seed=6)
    pr = np_oracle.orthographic_project([np_oracle.smpl_layer_call(host_model, p), p], 2).astype(np.float64)
    got = np_oracle.projects_to_silhouette(pr, wh)
    cols, rows = np.meshgrid(np.arange(wh), np.arange(wh))
    grid = np.stack([cols.ravel(), rows.ravel()], 1).astype(np.float64)
    for n in range(pr.shape[0]):
        d, _ = cKDTree(pr[n, :, :2]).query(grid)
        s = np.exp(-d / 1.2).reshape(wh, wh)[::-1]
        assert np.abs(got[n, ..., 1] - s).max() <= 1e-12
        assert np.abs(got[n, ..., 0] - (1 - s)).max() <= 1e-12


def test_focal_loss_matches_log_softmax():
    rng = np.random.default_rng(12)
    seg = rng.random((2, 30, 32))
    lab = rng.integers(0, 32, (2, 30))
    y = np.eye(32)[lab]
    got = np_oracle.categorical_focal_loss(y, np_oracle.softmax_last_axis(seg), 2.0, True)
    lp = np.take_along_axis(log_softmax(seg, axis=-1), lab[..., None], 2)[..., 0]
    w = np_oracle.focal_class_weights(32, np.float64)[lab]
    assert np.abs(got - w * (1 - np.exp(lp)) ** 2 * -lp).max() <= 1e-12
