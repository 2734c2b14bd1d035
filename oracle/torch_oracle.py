"""ORACLE — TEST INFRASTRUCTURE ONLY.  Gradients pinned to torch autograd through the reference's own source
(tests/test_reference_pin.py; see oracle/np_oracle.py header).

torch-CPU autograd twin of oracle/np_oracle.py: the same statements written with torch ops so that
``torch.autograd`` reproduces the gradient semantics TensorFlow's autodiff gives the reference
(SURVEY.md section 3.3):
  * reduce_max -> ``torch.amax`` (gradient split evenly among exact ties, like TF's _MinOrMaxGrad),
  * tf.norm    -> explicit ``sqrt(sum(d*d))``,
  * clip_by_value -> ``torch.clamp`` (gradient passes on the closed interval),
  * compute_mask has back_prop=False -> computed under no_grad via the NumPy oracle.
It is also what bench.py times as the multi-threaded CPU baseline (torch intra-op threads).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
All file:line citations are relative to /root/reference.
"""
from __future__ import annotations

import numpy as np
import torch

from . import np_oracle


def _t(a, dt):
    return torch.as_tensor(np.asarray(a), dtype=dt)


class TorchSmplConstants:
    def __init__(self, model, dtype=torch.float32):
        self.dtype = dtype
        self.v_template = _t(model.v_template, dtype)
        self.shapedirs = _t(model.shapedirs, dtype)
        self.posedirs = _t(model.posedirs, dtype)
        self.J_regressor = _t(model.J_regressor, dtype)
        self.lbs_weights = _t(model.lbs_weights, dtype)
        self.joint_regressor = _t(model.joint_regressor, dtype)
        self.parents = [int(p) for p in model.parents]


def batch_rodrigues(theta):
    """batch_smpl.py:255-276."""
    n = theta.shape[0]
    tp = theta + 1e-8
    angle = torch.sqrt(torch.sum(tp * tp, dim=1)).unsqueeze(-1)
    r = (theta / angle).unsqueeze(-1)
    angle = angle.unsqueeze(-1)
    cos, sin = torch.cos(angle), torch.sin(angle)
    outer = torch.matmul(r, r.transpose(1, 2))
    eyes = torch.eye(3, dtype=theta.dtype).unsqueeze(0).repeat(n, 1, 1)
    v = r[:, :, 0]
    zero = torch.zeros_like(v[:, 0])
    skew = torch.stack([zero, -v[:, 2], v[:, 1], v[:, 2], zero, -v[:, 0], -v[:, 1], v[:, 0], zero], 1).reshape(n, 3, 3)
    return cos * eyes + (1 - cos) * outer + sin * skew


def batch_global_rigid_transformation(Rs, Js, parent):
    """batch_smpl.py:168-228."""
    N = Rs.shape[0]
    dt = Rs.dtype
    Js = Js.unsqueeze(-1)

    def make_A(R, t):
        R_homo = torch.nn.functional.pad(R, (0, 0, 0, 1))
        t_homo = torch.cat([t, torch.ones(N, 1, 1, dtype=dt)], 1)
        return torch.cat([R_homo, t_homo], 2)

    results = [make_A(Rs[:, 0], Js[:, 0])]
    for i in range(1, len(parent)):
        j_here = Js[:, i] - Js[:, parent[i]]
        results.append(torch.matmul(results[parent[i]], make_A(Rs[:, i], j_here)))
    results = torch.stack(results, dim=1)
    new_J = results[:, :, :3, 3]
    Js_w0 = torch.cat([Js, torch.zeros(N, 24, 1, 1, dtype=dt)], 2)
    init_bone = torch.matmul(results, Js_w0)
    init_bone = torch.nn.functional.pad(init_bone, (3, 0))
    return new_J, results - init_bone


def smpl_layer_call(C: TorchSmplConstants, x, return_all=False, joint_type="lsp"):
    """batch_smpl.py:96-153."""
    N = x.shape[0]
    V = C.v_template.shape[0]
    thetas, betas = x[:, 4:76], x[:, 76:]
    v_shaped = (betas @ C.shapedirs).reshape(-1, V, 3) + C.v_template
    J = torch.stack([v_shaped[:, :, k] @ C.J_regressor for k in range(3)], dim=2)
    Rs = batch_rodrigues(thetas.reshape(-1, 3)).reshape(-1, 24, 3, 3)
    pose_feature = (Rs[:, 1:] - torch.eye(3, dtype=x.dtype)).reshape(-1, 207)
    v_posed = (pose_feature @ C.posedirs).reshape(-1, V, 3) + v_shaped
    J_transformed, A = batch_global_rigid_transformation(Rs, J, C.parents)
    W = C.lbs_weights.repeat(N, 1).reshape(N, -1, 24)
    T = torch.matmul(W, A.reshape(N, 24, 16)).reshape(N, -1, 4, 4)
    v_posed_homo = torch.cat([v_posed, torch.ones(N, V, 1, dtype=x.dtype)], 2)
    verts = torch.matmul(T, v_posed_homo.unsqueeze(-1))[:, :, :3, 0]
    if not return_all:
        return verts
    jr = C.joint_regressor[:, :14] if joint_type == "lsp" else C.joint_regressor
    joints = torch.stack([verts[:, :, k] @ jr for k in range(3)], dim=2)
    return dict(verts=verts, J_transformed=J_transformed, A=A, v_posed=v_posed, pose_feature=pose_feature,
                joints=joints)


def orthographic_project(inputs, vertex_sampling):
    """projection.py:54-81."""
    verts, smpl = inputs
    if vertex_sampling is not None:
        verts = verts[:, ::vertex_sampling, :]
    u = smpl[:, 2:3] + verts[:, :, 0] * smpl[:, 0:1]
    v = smpl[:, 3:4] + verts[:, :, 1] * smpl[:, 1:2]
    return torch.stack([u, v, verts[:, :, 2]], dim=2)


def compute_mask(pwd):
    """compute_mask.py:12-32; back_prop=False -> no gradient."""
    with torch.no_grad():
        return torch.from_numpy(np_oracle.compute_mask(pwd.detach().to(torch.float32).numpy())).to(pwd.dtype)


def _grid(img_wh, dt):
    r, c = torch.meshgrid(torch.arange(img_wh), torch.arange(img_wh), indexing="ij")
    return torch.stack([c, r], dim=2).to(dt).reshape(-1, 2)          # (column, row) per pixel, row-major


def projects_to_seg(inputs, img_wh, vertex_sampling, part_indices):
    """projects_to_seg.py:9-69."""
    pwd, mask_vals = inputs
    projects = pwd[:, :, :2]
    grid = _grid(img_wh, pwd.dtype)
    segs = []
    for indices in part_indices:
        if vertex_sampling is not None:
            indices = [i // vertex_sampling for i in indices]
        idx = torch.as_tensor(indices, dtype=torch.long)
        diff = projects[:, idx][:, None, :, :] - grid[None, :, None, :]
        norm = torch.sqrt(diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1])
        norm = norm * mask_vals[:, idx][:, None, :]
        segs.append(torch.amax(torch.exp(-norm), dim=2).reshape(-1, img_wh, img_wh))
    stacked = torch.stack(segs, dim=3)
    sil = 1.0 - torch.clamp(stacked.sum(dim=3), 0, 1)
    out = torch.cat([sil.unsqueeze(3), stacked], dim=3)
    return torch.flip(out, dims=[1])


def projects_to_silhouette(pwd, img_wh, row_chunk=8):
    """projects_to_silhouette.py:14-44."""
    projects = pwd[:, :, :2]
    grid = _grid(img_wh, pwd.dtype)
    chunks = []
    step = row_chunk * img_wh
    for s in range(0, img_wh * img_wh, step):
        diff = projects[:, None, :, :] - grid[s:s + step][None, :, None, :]
        norm = torch.sqrt(diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1])
        chunks.append(torch.amax(torch.exp(-norm / 1.2), dim=2))
    sil = torch.cat(chunks, dim=1).reshape(-1, img_wh, img_wh)
    return torch.flip(torch.stack([1.0 - sil, sil], dim=3), dims=[1])


def decode(C: TorchSmplConstants, params, img_wh, vertex_sampling, part_indices, silhouette_wh=None):
    """model.py:108-118 tail (+ silhouette branch, train_stage2_silhouette.py:84)."""
    allv = smpl_layer_call(C, params, return_all=True)
    pwd = orthographic_project([allv["verts"], params], vertex_sampling)
    mask = compute_mask(pwd)
    seg = projects_to_seg([pwd, mask], img_wh, vertex_sampling, part_indices)
    out = dict(allv, projects=pwd, mask=mask, seg=seg)
    if silhouette_wh:
        out["silhouette"] = projects_to_silhouette(pwd, silhouette_wh)
    return out


def softmax_focal_loss(y_true, seg, gamma=2.0, weight_classes=False, from_logits=True):
    """model.py:119-120 softmax followed by focal_loss.py:12-46, with torch autograd standing in for TF's
    (torch.clamp passes the gradient on the closed interval, like tf.clip_by_value)."""
    dt = seg.dtype
    y_pred = torch.softmax(seg, dim=-1) if from_logits else seg
    y_pred = torch.clamp(y_pred, np_oracle.KERAS_EPSILON, 1.0 - np_oracle.KERAS_EPSILON)
    cross_entropy = -y_true.to(dt) * torch.log(y_pred)
    if weight_classes:
        cross_entropy = cross_entropy * torch.as_tensor(np_oracle.focal_class_weights(seg.shape[-1], np.float64), dtype=dt)
    return (torch.pow(1.0 - y_pred, gamma) * cross_entropy).sum(dim=2)
