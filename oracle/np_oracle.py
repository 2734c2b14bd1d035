"""ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned to the reference's own source (see below); TensorFlow itself cannot run here.

NumPy restatement of the reference's keras_smpl decoder path, statement by statement, evaluated the
way the reference evaluates it (dense matmuls, brute-force O(wh^2 * V) rasterisers, O(4096 * V) mask).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product (indirect_learning_pose-shape_b200/) never does and fails loudly without its CUDA
library.

How it is pinned: the reference is python-2.7 / Keras 2.1 / TensorFlow 1.x code (README.md:20-28) with no
tests, golden vectors or recorded outputs; TensorFlow, Keras, h5py, deepdish, chumpy and python2 are absent
from this image and the SMPL model file it needs (neutral_smpl_with_cocoplus_reg.pkl) is not shipped
(/root/reference/.MISSING_LARGE_BLOBS).  oracle/make_reference_vectors.py therefore imports the reference's
OWN, UNMODIFIED keras_smpl/*.py and focal_loss.py with `tensorflow` / `keras` / `deepdish` / `cPickle` resolved to
the torch-CPU stand-ins of oracle/tf_shim/ (the ~70 TF/Keras symbols those files call, with TF's documented
semantics), runs them on seeded inputs and commits the results as tests/golden/reference_vectors.npz;
tests/test_reference_pin.py holds this restatement (and oracle/torch_oracle.py's gradients) to those vectors
(masks and labels identical, vertices 3e-7, scores 2.4e-7, initialisation functions bit-exact).  Not pinned, because
nothing here can run it: the floating-point rounding inside TensorFlow's own kernels (matmul summation order, exp).
It is additionally validated by (i) an fp64 twin of itself, (ii) finite differences of the torch twin,
(iii) closed-form properties, (iv) independent scipy implementations (tests/test_oracle_independent.py) and
(v) the drift guard tests/golden/oracle_vectors.npz.

All file:line citations are relative to /root/reference.
"""
from __future__ import annotations

import numpy as np

NUM_CAM = 4          # batch_smpl.py:90
NUM_THETAS = 72      # batch_smpl.py:73
MASK_GRID = 64       # compute_mask.py:44 (hard-coded, independent of img_wh)
MASK_INVISIBLE = 500.0   # compute_mask.py:68


# ------------------------------------------------------------------------------------------------
# batch_smpl.py
# ------------------------------------------------------------------------------------------------
def batch_skew(vec):
    """batch_smpl.py:230-253: scatter of [-z, y, z, -x, -y, x] into columns [1,2,3,5,6,7] of a 3x3."""
    n = vec.shape[0]
    res = np.zeros((n, 9), vec.dtype)
    res[:, 1] = -vec[:, 2]
    res[:, 2] = vec[:, 1]
    res[:, 3] = vec[:, 2]
    res[:, 5] = -vec[:, 0]
    res[:, 6] = -vec[:, 1]
    res[:, 7] = vec[:, 0]
    return res.reshape(n, 3, 3)


def batch_rodrigues(theta):
    """batch_smpl.py:255-276.  theta: (N*24, 3)."""
    dt = theta.dtype
    tp = theta + dt.type(1e-8)                                                   # :265 (inside the norm only)
    angle = np.sqrt(np.sum(tp * tp, axis=1, dtype=dt))[:, None]                  # tf.norm = sqrt(sum(x*x))
    r = (theta / angle)[:, :, None]                                              # :266
    angle = angle[:, :, None]                                                    # :268
    cos = np.cos(angle)
    sin = np.sin(angle)
    outer = np.matmul(r, np.transpose(r, (0, 2, 1)))                             # :272
    eyes = np.tile(np.eye(3, dtype=dt)[None], (theta.shape[0], 1, 1))            # :273
    R = cos * eyes + (dt.type(1) - cos) * outer + sin * batch_skew(r[:, :, 0])   # :274-275
    return R


def batch_global_rigid_transformation(Rs, Js, parent):
    """batch_smpl.py:168-228 with rotate_base=False.  Returns (new_J (N,24,3), A (N,24,4,4))."""
    dt = Rs.dtype
    N = Rs.shape[0]
    root_rotation = Rs[:, 0, :, :]                                               # :192
    Js = Js[..., None]                                                           # :195

    def make_A(R, t):                                                            # :197-202
        R_homo = np.pad(R, [[0, 0], [0, 1], [0, 0]])
        t_homo = np.concatenate([t, np.ones((N, 1, 1), dt)], 1)
        return np.concatenate([R_homo, t_homo], 2)

    A0 = make_A(root_rotation, Js[:, 0])                                         # :204
    results = [A0]
    for i in range(1, parent.shape[0]):                                          # :206-211
        j_here = Js[:, i] - Js[:, parent[i]]
        A_here = make_A(Rs[:, i], j_here)
        res_here = np.matmul(results[parent[i]], A_here)
        results.append(res_here)
    results = np.stack(results, axis=1)                                          # :214
    new_J = results[:, :, :3, 3]                                                 # :216
    Js_w0 = np.concatenate([Js, np.zeros((N, 24, 1, 1), dt)], 2)                 # :222
    init_bone = np.matmul(results, Js_w0)                                        # :223
    init_bone = np.pad(init_bone, [[0, 0], [0, 0], [0, 0], [3, 0]])              # :225
    A = results - init_bone                                                      # :226
    return new_J, A


def smpl_layer_call(model, x, return_all=False, joint_type="lsp"):
    """SMPLLayer.call (batch_smpl.py:96-153).  ``model`` is a SmplHostModel-like object whose arrays
    already have the layouts ``build`` produces (:31-94); ``x`` is (N,86).  dtype follows ``x``."""
    dt = x.dtype
    c = lambda a: np.asarray(a, dt)  # noqa: E731
    v_template, shapedirs, posedirs = c(model.v_template), c(model.shapedirs), c(model.posedirs)
    J_regressor, lbs_weights = c(model.J_regressor), c(model.lbs_weights)
    parents = np.asarray(model.parents)
    N = x.shape[0]
    V = v_template.shape[0]
    thetas = x[:, NUM_CAM:NUM_THETAS + NUM_CAM]                                  # :98
    betas = x[:, NUM_CAM + NUM_THETAS:]                                          # :99
    v_shaped = np.reshape(betas @ shapedirs, [-1, V, 3]) + v_template            # :106-108
    Jx = v_shaped[:, :, 0] @ J_regressor                                         # :112-115
    Jy = v_shaped[:, :, 1] @ J_regressor
    Jz = v_shaped[:, :, 2] @ J_regressor
    J = np.stack([Jx, Jy, Jz], axis=2)
    Rs = np.reshape(batch_rodrigues(np.reshape(thetas, [-1, 3])), [-1, 24, 3, 3])  # :119-120
    pose_feature = np.reshape(Rs[:, 1:, :, :] - np.eye(3, dtype=dt), [-1, 207])  # :122
    v_posed = np.reshape(pose_feature @ posedirs, [-1, V, 3]) + v_shaped         # :126-128
    J_transformed, A = batch_global_rigid_transformation(Rs, J, parents)         # :131
    W = np.reshape(np.tile(lbs_weights, [N, 1]), [N, -1, 24])                    # :135-136
    T = np.reshape(np.matmul(W, np.reshape(A, [N, 24, 16])), [N, -1, 4, 4])      # :138-140
    v_posed_homo = np.concatenate([v_posed, np.ones([N, V, 1], dt)], 2)          # :141-142
    v_homo = np.matmul(T, v_posed_homo[..., None])                               # :143
    verts = v_homo[:, :, :3, 0]                                                  # :145
    if not return_all:
        return verts
    joint_regressor = c(model.joint_regressor)
    if joint_type == "lsp":                                                      # :86-87
        joint_regressor = joint_regressor[:, :14]
    joints = np.stack([verts[:, :, k] @ joint_regressor for k in range(3)], axis=2)  # :147-151 (commented out upstream)
    return dict(verts=verts, J_transformed=J_transformed, A=A, Rs=Rs, J=J, v_shaped=v_shaped, v_posed=v_posed,
                pose_feature=pose_feature, joints=joints)


# ------------------------------------------------------------------------------------------------
# projection.py
# ------------------------------------------------------------------------------------------------
def orthographic_project(inputs, vertex_sampling):
    """projection.py:54-81."""
    verts, smpl = inputs
    k_u, k_v, u0, v0 = smpl[:, 0:1], smpl[:, 1:2], smpl[:, 2:3], smpl[:, 3:4]
    if vertex_sampling is not None:
        verts = verts[:, ::vertex_sampling, :]                                   # :67-68
    u = u0 + verts[:, :, 0] * k_u                                                # :77 (mul then add: two ops)
    v = v0 + verts[:, :, 1] * k_v                                                # :78
    return np.stack([u, v, verts[:, :, 2]], axis=2)                              # :79


# ------------------------------------------------------------------------------------------------
# compute_mask.py
# ------------------------------------------------------------------------------------------------
def compute_mask_one(pixels_with_depth):
    """compute_mask_map_over_batch + get_min_depth_vert_index_at_pixel (compute_mask.py:35-108), literal:
    one pass per pixel of the hard-coded 64x64 grid, argmax of z among the vertices that round onto it
    (first index on ties), index 1 for an empty pixel, unique, scatter 1 into a 500-filled vector."""
    img_wh = MASK_GRID
    nv = pixels_with_depth.shape[0]
    pu, pv, z = pixels_with_depth[:, 0], pixels_with_depth[:, 1], pixels_with_depth[:, 2]
    winners = []
    for r in range(img_wh):              # meshgrid 'xy': pixel_coord = (column, row), :49-54
        row_sel = pv == np.float32(r)
        for c in range(img_wh):
            at = np.nonzero(row_sel & (pu == np.float32(c)))[0]                  # :90-92
            if at.size == 0:
                winners.append(1)                                                # tf.ones([1,1,4]) -> index 1, :98-99
            else:
                winners.append(int(at[np.argmax(z[at])]))                        # :100-102 (argmax, first on ties)
    mask = np.ones(nv, np.float32) * np.float32(MASK_INVISIBLE)                  # :68
    mask[np.unique(np.asarray(winners, np.int64))] = 1.0                         # :66,:69-70
    return mask


def compute_mask_one_fast(pixels_with_depth):
    """Same result as compute_mask_one via a lexsort (used for larger oracle batches; equality is tested)."""
    img_wh = MASK_GRID
    nv = pixels_with_depth.shape[0]
    pu, pv, z = pixels_with_depth[:, 0], pixels_with_depth[:, 1], pixels_with_depth[:, 2]
    inside = (pu >= 0) & (pu <= img_wh - 1) & (pv >= 0) & (pv <= img_wh - 1)
    idx = np.nonzero(inside)[0]
    cell = (pv[idx].astype(np.int64) * img_wh + pu[idx].astype(np.int64))
    zz = z[idx] + np.float32(0.0)
    order = np.lexsort((idx, -zz.astype(np.float64), cell))      # by cell, then z descending, then index ascending
    cs = cell[order]
    first = np.ones(cs.shape[0], bool)
    first[1:] = cs[1:] != cs[:-1]
    win = idx[order][first]
    mask = np.full(nv, MASK_INVISIBLE, np.float32)
    mask[win] = 1.0
    if np.unique(cs).size < img_wh * img_wh and nv > 1:
        mask[1] = 1.0
    return mask


def compute_mask(batch_projects_with_depth, fast=True):
    """compute_mask.py:12-32.  tf.round is round-half-to-even, as is np.round (np.rint)."""
    p = np.asarray(batch_projects_with_depth, np.float32)
    batch_pixels = np.rint(p[:, :, :2])                                          # :22
    pwd = np.concatenate([batch_pixels, p[:, :, 2:3]], axis=2)                   # :23-25
    f = compute_mask_one_fast if fast else compute_mask_one
    return np.stack([f(pwd[n]) for n in range(pwd.shape[0])], 0)                 # :27-30 (map_fn, stateless intent)


# ------------------------------------------------------------------------------------------------
# projects_to_seg.py / projects_to_silhouette.py
# ------------------------------------------------------------------------------------------------
def _grid(img_wh, dt):
    t1, t2 = np.meshgrid(np.arange(img_wh), np.arange(img_wh))                   # 'xy': t1 = column, t2 = row
    return np.stack([t1, t2], axis=2).astype(dt).reshape(-1, 2)                  # projects_to_seg.py:26-31


def projects_to_seg(inputs, img_wh, vertex_sampling, part_indices, pixel_chunk=None):
    """projects_to_seg.py:9-69.  ``part_indices`` is the unpickled list of 31 lists of ORIGINAL vertex ids
    (the reference reads it from disk at :18-24)."""
    projects_with_depth, mask_vals = inputs
    dt = projects_with_depth.dtype
    projects = projects_with_depth[:, :, :2]
    N = projects.shape[0]
    reshaped_grid = _grid(img_wh, dt)                                            # (wh^2, 2)
    segs = []
    for part in range(len(part_indices)):
        indices = part_indices[part]
        if vertex_sampling is not None:
            indices = [index // vertex_sampling for index in indices]            # :36-37
        indices = np.asarray(indices, np.int64)
        part_projects = projects[:, indices, :]                                  # (N, n, 2)   :41
        part_mask_vals = mask_vals[:, indices]                                   # (N, n)      :45
        diff = part_projects[:, None, :, :] - reshaped_grid[None, :, None, :]    # (N, wh^2, n, 2) :52
        norm = np.sqrt(diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1])  # tf.norm: sqrt(sum(x*x))  :53
        norm = norm * part_mask_vals[:, None, :]                                 # :54
        exp = np.exp(-norm)                                                      # :55
        scores = exp.max(axis=2) if indices.size else np.zeros((N, img_wh * img_wh), dt)  # :56
        segs.append(scores.reshape(-1, img_wh, img_wh))                          # :57
    stacked = np.stack(segs, axis=3)                                             # :60
    sil = dt.type(1.0) - np.clip(np.sum(stacked, axis=3, dtype=dt), 0, 1)        # :61-64
    out = np.concatenate([sil[..., None], stacked], axis=3)                      # :66-67
    return out[:, ::-1].copy()                                                   # :68 flip rows


def projects_to_silhouette(projects_with_depth, img_wh, row_chunk=8):
    """projects_to_silhouette.py:14-44 (evaluated in row chunks only to bound memory; same arithmetic)."""
    dt = projects_with_depth.dtype
    projects = projects_with_depth[:, :, :2]
    N = projects.shape[0]
    grid = _grid(img_wh, dt)
    scores = np.empty((N, img_wh * img_wh), dt)
    step = row_chunk * img_wh
    for s in range(0, img_wh * img_wh, step):
        g = grid[s:s + step]
        diff = projects[:, None, :, :] - g[None, :, None, :]                     # :35
        norm = np.sqrt(diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1])  # :36
        exp = np.exp(-norm / dt.type(1.2))                                       # :37
        scores[:, s:s + step] = exp.max(axis=2)                                  # :38
    sil = scores.reshape(-1, img_wh, img_wh)                                     # :39
    out = np.stack([dt.type(1.0) - sil, sil], axis=3)                            # :40-41
    return out[:, ::-1].copy()                                                   # :42


# ------------------------------------------------------------------------------------------------
# concat_mean_param.py / set_cam_params.py
# ------------------------------------------------------------------------------------------------
def _mean86(img_wh, mean_vals, with_smpl=True):
    mean = np.zeros((1, 86))
    if with_smpl:
        mean_pose = np.array(mean_vals["pose"], np.float64)
        mean_pose[:3] = 0.0                                                      # concat_mean_param.py:14
        mean[0, 4:] = np.hstack((mean_pose, np.asarray(mean_vals["shape"], np.float64)))
    mean[0, 0] = img_wh / 2.0
    mean[0, 1] = img_wh / 2.0
    mean[0, 2] = img_wh / 2.0
    mean[0, 3] = img_wh / 1.6
    return mean.astype(np.float32)                                               # tf.constant(..., float32)


def concat_mean_param(img_features, img_wh, mean_vals):
    """concat_mean_param.py:8-31."""
    mean = np.tile(_mean86(img_wh, mean_vals), [img_features.shape[0], 1])
    return np.concatenate([img_features, mean], axis=1)


def set_cam_params(smpl, img_wh):
    """set_cam_params.py:13-26."""
    return smpl + np.tile(_mean86(img_wh, None, with_smpl=False), [smpl.shape[0], 1])


def load_mean_set_cam_params(smpl, img_wh, mean_vals):
    """set_cam_params.py:29-52."""
    return smpl + np.tile(_mean86(img_wh, mean_vals), [smpl.shape[0], 1])


# ------------------------------------------------------------------------------------------------
# whole path
# ------------------------------------------------------------------------------------------------
def decode(model, params, img_wh, vertex_sampling, part_indices, silhouette_wh=None):
    """model.py:108-118 tail: SMPLLayer -> orthographic_project -> compute_mask -> projects_to_seg
    (+ train_stage2_silhouette.py:84 silhouette branch if ``silhouette_wh``)."""
    params = np.asarray(params, np.float32)
    allv = smpl_layer_call(model, params, return_all=True)
    pwd = orthographic_project([allv["verts"], params], vertex_sampling)
    mask = compute_mask(pwd)
    seg = projects_to_seg([pwd, mask], img_wh, vertex_sampling, part_indices)
    out = dict(allv, projects=pwd, mask=mask, seg=seg)
    if silhouette_wh:
        out["silhouette"] = projects_to_silhouette(pwd, silhouette_wh)
    return out


# ---------------------------------------------------------------------------------------------------------------
# the op after the path: softmax (model.py:119-120) + categorical focal loss (focal_loss.py:10-48)
# ---------------------------------------------------------------------------------------------------------------
KERAS_EPSILON = 1e-7


def focal_class_weights(num_classes=32, dt=np.float32):
    """focal_loss.py:21-38: up-weight hands, elbows, knees, ankles, down-weight the background."""
    w = np.ones(num_classes, dt)
    w[0] = 0.3
    for c in (1, 2, 3, 4, 10, 12, 14, 15, 16, 17, 23, 25):
        if c < num_classes:        # the reference hard-codes 32 classes; narrower tables keep the same ids
            w[c] = 2.0
    return w


def softmax_last_axis(x):
    """model.py:120 Activation('softmax'): exp(x - max) / sum on the last axis (Keras / TF softmax)."""
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def categorical_focal_loss(y_true, y_pred, gamma=2.0, weight_classes=False):
    """focal_loss.py:12-46.  y_true, y_pred (N, img_wh^2, C); returns (N, img_wh^2)."""
    dt = y_pred.dtype
    y_pred = np.clip(y_pred, dt.type(KERAS_EPSILON), dt.type(1.0) - dt.type(KERAS_EPSILON))      # :16
    cross_entropy = -y_true.astype(dt) * np.log(y_pred)                                           # :17
    if weight_classes:
        cross_entropy = cross_entropy * focal_class_weights(y_pred.shape[-1], dt)                 # :19-41
    focal = np.power(dt.type(1.0) - y_pred, dt.type(gamma)) * cross_entropy                       # :44
    return focal.sum(axis=2)                                                                      # :45


# ---------------------------------------------------------------------------------------------------------------
# the op before the path: the regression module (model.py:63-105)
# ---------------------------------------------------------------------------------------------------------------
def dense(x, kernel, bias, activation="linear"):
    """keras.layers.Dense: K.dot(x, kernel) + bias, then the activation (model.py:66-68, 100-102)."""
    y = x @ kernel + bias
    return np.maximum(y, 0) if activation == "relu" else y


def ief_regressor(img_features, weights, img_wh, mean_vals, scaledown=0.005, iterations=3):
    """model.py:63-97.  weights = [(kernel, bias)] x 3 for IEF_layer_1..3, SHARED by the iterations."""
    state = concat_mean_param(img_features, img_wh, mean_vals)                    # :70-71
    param = state[:, img_features.shape[1]:]                                      # :72
    for _ in range(iterations):
        delta = dense(state, weights[0][0], weights[0][1], "relu")                # :77 / :86 / :95
        delta = dense(delta, weights[1][0], weights[1][1], "relu")
        delta = dense(delta, weights[2][0], weights[2][1], "linear")
        delta = delta * img_features.dtype.type(scaledown)                        # :80
        param = param + delta                                                     # :81
        state = np.concatenate([img_features, param], axis=1)                     # :82
    return param


def plain_regressor(img_features, weights, img_wh, mean_vals, scaledown=0.005):
    """model.py:99-105."""
    smpl = dense(img_features, weights[0][0], weights[0][1], "relu")
    smpl = dense(smpl, weights[1][0], weights[1][1], "relu")
    smpl = dense(smpl, weights[2][0], weights[2][1], "linear")
    smpl = smpl * img_features.dtype.type(scaledown)
    return load_mean_set_cam_params(smpl, img_wh, mean_vals)


# ---------------------------------------------------------------------------------------------------------------
# the consumer of `verts`: the mesh visualiser (renderer.py:23-115,146-197)
# ---------------------------------------------------------------------------------------------------------------
# renderer.py drives OpenDR (opendr.camera.ProjectPoints, opendr.lighting.LambertianPointLight,
# opendr.renderer.ColoredRenderer; unpinned, absent from the reference tree and from this image).  What follows restates
# OpenDR's published algorithm in float64 at exactly the reference's call sites; oracle/tf_shim/opendr wraps these same
# functions in OpenDR's names so that renderer.py itself can be run (oracle/make_render_vectors.py).  "parity unpinned"
# for OpenDR's GL rasterisation: the fill rule at exact edge hits, GL's fixed-point vertex snapping and the `overdraw`
# anti-aliasing of silhouette edges are not restated.
RENDER_COLORS = {"light_blue": [0.65098039, 0.74117647, 0.85882353], "light_pink": [.9, .7, .7]}   # renderer.py:16-20
# simple_renderer's three point lights, renderer.py:171-195: (position, colour)
RENDER_LIGHTS = (([-200.0, -100.0, -100.0], [1.0, 1.0, 1.0]), ([800.0, 10.0, 300.0], [1.0, 1.0, 1.0]),
                 ([-500.0, 500.0, 1000.0], [0.7, 0.7, 0.7]))


def rotate_y(points, angle):
    """renderer.py:138-143 (_rotateY)."""
    ry = np.array([[np.cos(angle), 0., np.sin(angle)], [0., 1., 0.], [-np.sin(angle), 0., np.cos(angle)]])
    return np.dot(points, ry)


def vert_normals(verts, faces):
    """opendr.geometry.VertNormals: the (v1 - v0) x (v2 - v0) of every face summed onto its three vertices, normalised."""
    v = np.asarray(verts, np.float64)
    tn = np.cross(v[faces[:, 1]] - v[faces[:, 0]], v[faces[:, 2]] - v[faces[:, 0]])
    n = np.zeros_like(v)
    for c in range(3):
        np.add.at(n, faces[:, c], tn)
    nn = np.sqrt((n * n).sum(1, keepdims=True))
    return np.where(nn > 0, n / np.where(nn > 0, nn, 1.0), 0.0)


def lambertian_point_light(verts, faces, light_pos, vc, light_color):
    """opendr.lighting.LambertianPointLight (single sided): max(n . normalise(light_pos - v), 0) * vc * light_color."""
    v = np.asarray(verts, np.float64)
    vn = vert_normals(v, faces)
    ld = np.asarray(light_pos, np.float64).reshape(1, 3) - v
    ld = ld / np.sqrt((ld * ld).sum(1, keepdims=True))
    d = np.maximum((vn * ld).sum(1), 0.0)
    return d.reshape(-1, 1) * np.asarray(vc, np.float64).reshape(-1, 3) * np.asarray(light_color, np.float64).reshape(1, 3)


def _edge(ax, ay, bx, by, px, py):
    return (bx - ax) * (py - ay) - (by - ay) * (px - ax)


def _edge_owns(ax, ay, bx, by):
    dx, dy = bx - ax, by - ay
    return (dy < 0) or (dy == 0 and dx < 0)


def rasterise(verts, faces, vc, f, c, h, w, near, far, background=None):
    """opendr.renderer.ColoredRenderer.r for camera ProjectPoints(f, rt=0, t=0, k=0, c): z-buffered triangles, vertex colours
    clamped to [0,1] and interpolated perspective-correctly, 8-bit frame buffer.  Pixel (r, c) is sampled at the projected
    position (c, r).  Returns the float image OpenDR hands back: uint8 / 255. (h, w, 3)."""
    v = np.asarray(verts, np.float64)
    col = np.clip(np.broadcast_to(np.asarray(vc, np.float64).reshape(-1, 3), v.shape), 0.0, 1.0)
    Z = v[:, 2]
    iz = np.where(Z > 0, 1.0 / np.where(Z > 0, Z, 1.0), 0.0)
    fx, fy = (float(f[0]), float(f[1])) if np.ndim(f) else (float(f), float(f))
    X = fx * v[:, 0] * iz + float(c[0])
    Y = fy * v[:, 1] * iz + float(c[1])
    iz_hi = 1.0 / near if near > 0 else np.inf
    iz_lo = 1.0 / far if far > 0 else 0.0
    best = np.full((h, w), -1.0)
    img = np.ones((h, w, 3)) if background is None else np.array(background, np.float64)
    if background is not None:
        img = np.rint(np.clip(img, 0.0, 1.0) * 255.0) / 255.0      # the 8-bit frame buffer holds the background too
    for t in range(faces.shape[0]):
        i0, i1, i2 = (int(q) for q in faces[t])
        if not (Z[i0] > 0 and Z[i1] > 0 and Z[i2] > 0):
            continue
        x0, y0, x1, y1, x2, y2 = X[i0], Y[i0], X[i1], Y[i1], X[i2], Y[i2]
        area = _edge(x0, y0, x1, y1, x2, y2)
        if area == 0:
            continue
        c_lo, c_hi = max(int(np.ceil(min(x0, x1, x2))), 0), min(int(np.floor(max(x0, x1, x2))), w - 1)
        r_lo, r_hi = max(int(np.ceil(min(y0, y1, y2))), 0), min(int(np.floor(max(y0, y1, y2))), h - 1)
        if c_lo > c_hi or r_lo > r_hi:
            continue
        px, py = np.meshgrid(np.arange(c_lo, c_hi + 1, dtype=np.float64), np.arange(r_lo, r_hi + 1, dtype=np.float64))
        w0, w1, w2 = _edge(x1, y1, x2, y2, px, py), _edge(x2, y2, x0, y0, px, py), _edge(x0, y0, x1, y1, px, py)
        if area > 0:
            inside = ((w0 > 0) | ((w0 == 0) & _edge_owns(x1, y1, x2, y2))) & \
                     ((w1 > 0) | ((w1 == 0) & _edge_owns(x2, y2, x0, y0))) & \
                     ((w2 > 0) | ((w2 == 0) & _edge_owns(x0, y0, x1, y1)))
        else:
            inside = ((w0 < 0) | ((w0 == 0) & _edge_owns(x2, y2, x1, y1))) & \
                     ((w1 < 0) | ((w1 == 0) & _edge_owns(x0, y0, x2, y2))) & \
                     ((w2 < 0) | ((w2 == 0) & _edge_owns(x1, y1, x0, y0)))
        if not inside.any():
            continue
        s = w0 + w1 + w2
        s = np.where(s == 0, 1.0, s)
        b0, b1, b2 = w0 / s, w1 / s, w2 / s
        z = b0 * iz[i0] + b1 * iz[i1] + b2 * iz[i2]
        sub = best[r_lo:r_hi + 1, c_lo:c_hi + 1]
        win = inside & (z >= iz_lo) & (z <= iz_hi) & (z > sub)          # strict: the first face drawn keeps a tie (GL_LESS)
        if not win.any():
            continue
        q0, q1, q2 = b0 * iz[i0], b1 * iz[i1], b2 * iz[i2]
        rgb = (q0[..., None] * col[i0] + q1[..., None] * col[i1] + q2[..., None] * col[i2]) / (q0 + q1 + q2)[..., None]
        rgb = np.rint(np.clip(rgb, 0.0, 1.0) * 255.0) / 255.0         # the 8-bit frame buffer, read back as k / 255.
        sub[win] = z[win]
        img[r_lo:r_hi + 1, c_lo:c_hi + 1][win] = rgb[win]
    return img


def render_mesh(verts, faces, cam=None, img=None, do_alpha=False, far=None, near=None, color="light_blue", img_size=None,
                render_seg=False, part_colors=None, default_size=224, flength=500.0, yrot=0.0):
    """SMPLRenderer.__call__ -> render_model -> simple_renderer (renderer.py:34-85, 221-256, 146-197) on one mesh.
    ``color``: the albedo's name (renderer.py:240-244 picks it from the dict by ``color_id``); ``part_colors`` (6890,3)
    uint8: the PLY's vertex colours (render_seg, :158-168).  Returns uint8 (h, w, 3 or 4)."""
    verts = np.asarray(verts, np.float64)
    if img is not None:
        h, w = img.shape[:2]                                                     # :47-48
    elif img_size is not None:
        h, w = img_size[0], img_size[1]                                          # :49-51
    else:
        h = w = default_size                                                     # :52-54
    if cam is None:
        cam = [flength, w / 2., h / 2.]                                          # :56-57
    if near is None:
        near = np.maximum(np.min(verts[:, 2]) - 25, -0.2)                        # :64-65
    if far is None:
        far = np.maximum(np.max(verts[:, 2]) + 25, 25)                           # :66-67
    bg = None
    if img is not None:
        bg = img / 255. if img.max() > 1 else img                                # :231-232
    albedo = np.asarray(RENDER_COLORS[color], np.float64)                        # :234-244, :153-154
    if render_seg:
        vc = np.asarray(part_colors, np.float64) / 255.0                         # :158-168
    else:
        vc = np.zeros((verts.shape[0], 3))
        for pos, lc in RENDER_LIGHTS:                                            # :171-195
            vc = vc + lambertian_point_light(verts, faces, rotate_y(np.array(pos), yrot), albedo, np.array(lc))
    im = rasterise(verts, faces, vc, cam[0] * np.ones(2), cam[1:3], h, w, near, far, bg)   # :57-62, :223-224
    if img is None and do_alpha:                                                 # :252-253 get_alpha
        alpha = (~np.all(im == 1., axis=2)).astype(im.dtype)
        im = np.concatenate([im, alpha[..., None]], axis=2)
    elif img is not None and do_alpha:                                           # :254-255 append_alpha
        im = np.concatenate([im, np.ones_like(im[:, :, :1])], axis=2)
    return (im * 255).astype('uint8')                                            # :85


def rodrigues(r):
    """cv2.Rodrigues(rvec)[0]: the rotation matrix of an axis-angle vector (renderer.py:99-104)."""
    r = np.asarray(r, np.float64).reshape(3)
    th = np.sqrt((r * r).sum())
    if th < 1e-300:
        return np.eye(3)
    k = r / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(k, k) + np.sin(th) * K


def rotated_verts(verts, deg, axis='y'):
    """SMPLRenderer.rotated, renderer.py:97-107: rotate about the mesh's centroid."""
    a = np.radians(deg)
    around = rodrigues([0, a, 0] if axis == 'y' else ([a, 0, 0] if axis == 'x' else [0, 0, a]))
    center = verts.mean(axis=0)
    return np.dot((verts - center), around) + center
