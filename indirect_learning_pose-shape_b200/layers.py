"""Host-side mirror of the reference's keras_smpl interface, backed by the sm_100a library.

Names, argument order, list-style inputs and output layouts follow the reference (file:line relative to it):

  SMPLLayer(pkl_path, batch_size=8, dtype='float32', joint_type='lsp')   keras_smpl/batch_smpl.py:23-166
  orthographic_project([verts, smpl], vertex_sampling)                    keras_smpl/projection.py:54-81
  compute_mask(batch_projects_with_depth)                                 keras_smpl/compute_mask.py:12-32
  projects_to_seg([projects_with_depth, mask_vals], img_wh, vertex_sampling)   keras_smpl/projects_to_seg.py:9-69
  projects_to_silhouette(projects_with_depth, img_wh)                     keras_smpl/projects_to_silhouette.py:14-44
  concat_mean_param(img_features, img_wh)                                 keras_smpl/concat_mean_param.py:8-31
  set_cam_params(smpl, img_wh) / load_mean_set_cam_params(smpl, img_wh)   keras_smpl/set_cam_params.py:13-52

Tensors are torch CUDA float32; every op is a torch.autograd.Function whose forward and backward call the C ABI
(include/smpl_b200.h) on torch's current stream.  There is no CPU or eager-PyTorch fallback: a CPU tensor, a missing
library or a missing GPU raises.

Documented deviations from the reference:
  * `batch_size` is accepted but not required to match the runtime batch (the reference bakes it into the graph,
    batch_smpl.py:120,135-142).
  * compute_mask is stateless per sample (the docstring's intent, compute_mask.py:14-18), not the accidental
    cross-call K.variable state of :68-70.
  * gradients at a vertex that sits exactly on a pixel centre are 0 (TF's norm gradient gives NaN there); exact
    ties in the max take the first arg-max (TF splits evenly).  Both are measure-zero events.  (Duplicate entries of
    one vertex in a part -- `index // vertex_sampling` collisions, projects_to_seg.py:36-37 -- are exact ties by
    construction; after the gather's adjoint TF's even split and the first-arg-max rule give the vertex the same total,
    tests/test_gpu_parity.py::test_seg_duplicate_entries_tie_gradient.)
  * `categorical_focal_loss(..., from_logits=True)` / `categorical_crossentropy(..., from_logits=True)` are opt-in
    fused variants (softmax inside the kernel); the default takes probabilities, like the reference.
"""
from __future__ import annotations

import hashlib
import os
import threading
import weakref
from typing import List, Optional, Sequence

import ctypes as C
import numpy as np
import torch

from . import _lib, smpl_io

NUM_PARAMS = 86
NUM_JOINTS = 24


def _vs(vertex_sampling) -> int:
    return 1 if vertex_sampling is None else int(vertex_sampling)


def _check_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise _lib.SmplB200Error("%s is on %s: this package runs on CUDA only (no CPU fallback)" % (name, t.device))
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32 (got %s)" % (name, t.dtype))
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _ru256(nbytes: int) -> int:
    return (int(nbytes) + 255) // 256 * 256


def _vps_ld(num_sampled_verts: int) -> int:
    return (num_sampled_verts * 3 + 3) // 4 * 4            # SMPL_B200_VPS_LD


# ------------------------------------------------------------------------------------------------------------
# device handles
# ------------------------------------------------------------------------------------------------------------
class DeviceModel:
    """Immutable device copy of the SMPL constants (SMPLLayer.build, batch_smpl.py:31-94)."""

    def __init__(self, host_model: smpl_io.SmplHostModel, device):
        lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.SmplB200Error("SMPL model must live on a CUDA device (no CPU path)")
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", index)
        self.host = host_model
        struct, keep = _lib.make_host_model(host_model)
        handle = C.c_void_p()
        _lib.check(lib.smpl_b200_model_create(C.byref(struct), index, C.byref(handle)), "smpl_b200_model_create")
        del keep
        self.handle = handle
        self.V = int(lib.smpl_b200_model_num_verts(handle))
        self.lbs_width = int(lib.smpl_b200_model_lbs_width(handle))
        self.LD = (self.V + 255) // 256 * 768            # SMPL_B200_VPOSED_LD
        self.R = int(host_model.joint_regressor.shape[1])
        self._finalizer = weakref.finalize(self, lib.smpl_b200_model_destroy, handle)

    def workspace_bytes(self, op: int, N: int, img_wh: int = 0, vs: int = 1) -> int:
        return int(_lib.load().smpl_b200_workspace_bytes(self.handle, op, N, img_wh, vs))


class PartTable:
    """Device copy of a part->vertex list with indices divided by vertex_sampling (projects_to_seg.py:18-24,36-37)."""

    def __init__(self, parts: Sequence[Sequence[int]], vertex_sampling, num_sampled_verts: int, device):
        lib = _lib.load()
        self.device = torch.device(device)
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", index)
        ptr, idx = smpl_io.sampled_part_table(parts, vertex_sampling)
        self.P = len(parts)
        self.Vs = int(num_sampled_verts)
        handle = C.c_void_p()
        i32p = C.POINTER(C.c_int32)
        _lib.check(lib.smpl_b200_parts_create(index, self.P, ptr.ctypes.data_as(i32p), idx.ctypes.data_as(i32p), self.Vs,
                                              C.byref(handle)), "smpl_b200_parts_create")
        self.handle = handle
        self._finalizer = weakref.finalize(self, lib.smpl_b200_parts_destroy, handle)


_cache_lock = threading.Lock()
_model_cache = {}
_parts_cache = {}


def get_device_model(pkl_path_or_model, device) -> DeviceModel:
    """One DeviceModel per (model, device); accepts a pickle path (the reference's argument) or a SmplHostModel."""
    device = torch.device(device)
    index = device.index if device.index is not None else torch.cuda.current_device()
    key = (pkl_path_or_model if isinstance(pkl_path_or_model, str) else id(pkl_path_or_model), index)
    with _cache_lock:
        dm = _model_cache.get(key)
        if dm is None:
            host = smpl_io.load_smpl_pkl(pkl_path_or_model) if isinstance(pkl_path_or_model, str) else pkl_path_or_model
            dm = DeviceModel(host, torch.device("cuda", index))
            _model_cache[key] = dm
        return dm


def get_part_table(vertex_sampling, num_sampled_verts: int, device, part_indices_path: Optional[str] = None,
                   parts: Optional[Sequence[Sequence[int]]] = None) -> PartTable:
    """Part table for `vertex_sampling`.  Lookup order: explicit `parts`, explicit path, the reference's CWD-relative
    literal (projects_to_seg.py:18-21), the copy of the same files shipped in this package's data/."""
    device = torch.device(device)
    index = device.index if device.index is not None else torch.cuda.current_device()
    if parts is not None:
        # keyed on CONTENT: a freshly built but equal list (e.g. golden_part_vertices(5) per call) reuses the device table
        h = hashlib.sha1()
        for part in parts:
            h.update(np.asarray(part, np.int64).tobytes())
            h.update(b"|")
        src = "parts:" + h.hexdigest()
    else:
        src = "path:" + (part_indices_path or "")
    key = (src, _vs(vertex_sampling), int(num_sampled_verts), index)
    with _cache_lock:
        pt = _parts_cache.get(key)
        if pt is None:
            if parts is None:
                path = part_indices_path or smpl_io.part_vertices_filename(vertex_sampling)
                if os.path.exists(path):
                    parts = smpl_io.load_part_vertices(path)
                elif part_indices_path is not None:
                    raise IOError("part table %r not found" % part_indices_path)
                else:
                    parts = smpl_io.golden_part_vertices(vertex_sampling)
            pt = PartTable(parts, vertex_sampling, num_sampled_verts, torch.device("cuda", index))
            _parts_cache[key] = pt
        return pt


# ------------------------------------------------------------------------------------------------------------
# autograd functions
# ------------------------------------------------------------------------------------------------------------
class _DecodeFn(torch.autograd.Function):
    """params (N,86) -> (verts, J_transformed, keypoints, projects); batch_smpl.py:96-153 (+ projection.py:54-81 fused)."""

    @staticmethod
    def forward(ctx, params, dm: DeviceModel, need_verts: bool, num_keypoints: int, project_vs: int):
        lib = _lib.load()
        params = _check_cuda_f32(params, "params")
        if params.dim() != 2 or params.shape[1] != NUM_PARAMS:
            raise ValueError("params must be (N,86), got %s" % (tuple(params.shape),))
        if params.device != dm.device:
            raise _lib.SmplB200Error("params on %s but the SMPL model is on %s" % (params.device, dm.device))
        N = params.shape[0]
        dev = params.device
        with torch.cuda.device(dev):
            verts = torch.empty((N, dm.V, 3), dtype=torch.float32, device=dev) if need_verts else None
            joints = torch.empty((N, NUM_JOINTS, 3), dtype=torch.float32, device=dev)
            keyp = torch.empty((N, num_keypoints, 3), dtype=torch.float32, device=dev) if num_keypoints else None
            Vs = (dm.V + project_vs - 1) // project_vs if project_vs else 0
            proj = torch.empty((N, Vs, 3), dtype=torch.float32, device=dev) if project_vs else None
            # saved for backward: the full rest-pose mesh when a dense vertex gradient may come back (verts requested),
            # the compact copy of the sampled vertices when the gradient can only arrive through the projection
            keep_full = need_verts or not project_vs
            vp = torch.empty((N, dm.LD), dtype=torch.float32, device=dev) if keep_full else None
            # (with vertex_sampling = 1 the compact copy would duplicate the full one)
            vps = (torch.empty((N, _vps_ld(Vs)), dtype=torch.float32, device=dev)
                   if project_vs and not (keep_full and project_vs == 1) else None)
            ws_bytes = dm.workspace_bytes(_lib.OP_DECODE_FWD, N) + (0 if keep_full else _ru256(N * dm.LD * 4))
            ws = _workspace(ws_bytes, dev)
            _lib.check(lib.smpl_b200_decode_fwd(dm.handle, _ptr(params), N, _ptr(verts), _ptr(joints), _ptr(keyp),
                                                num_keypoints, _ptr(vp), _ptr(vps), _ptr(proj), max(project_vs, 1),
                                                _ptr(ws), ws.numel(), _stream()), "smpl_b200_decode_fwd")
        ctx.dm, ctx.project_vs = dm, project_vs
        ctx.have_vp, ctx.have_vps = vp is not None, vps is not None
        ctx.set_materialize_grads(False)          # unused outputs arrive as None, not as dense zero tensors
        ctx.save_for_backward(params, *[x for x in (vp, vps) if x is not None])
        if keyp is not None:
            ctx.mark_non_differentiable(keyp)     # dead code in the reference (batch_smpl.py:147-151): forward only
        return verts, joints, keyp, proj

    @staticmethod
    def backward(ctx, g_verts, g_joints, g_keyp, g_proj):
        lib = _lib.load()
        saved = list(ctx.saved_tensors)
        params = saved.pop(0)
        vp = saved.pop(0) if ctx.have_vp else None
        vps = saved.pop(0) if ctx.have_vps else None
        if g_verts is None and g_joints is None and g_proj is None:
            return None, None, None, None, None
        dm = ctx.dm
        N = params.shape[0]
        dev = params.device
        g_verts = None if g_verts is None else _check_cuda_f32(g_verts, "grad verts")
        g_joints = None if g_joints is None else _check_cuda_f32(g_joints, "grad joints")
        g_proj = None if g_proj is None else _check_cuda_f32(g_proj, "grad projects")
        vs = max(ctx.project_vs, 1)
        with torch.cuda.device(dev):
            g_params = torch.empty_like(params)
            ws_vs = 1 if (g_verts is not None or g_proj is None) else vs
            ws = _workspace(dm.workspace_bytes(_lib.OP_DECODE_BWD, N, 0, ws_vs), dev)
            _lib.check(lib.smpl_b200_decode_bwd(dm.handle, _ptr(params), N, _ptr(vp), _ptr(vps), _ptr(g_verts),
                                                _ptr(g_proj), vs, _ptr(g_joints), _ptr(g_params), _ptr(ws), ws.numel(),
                                                _stream()), "smpl_b200_decode_bwd")
        return g_params, None, None, None, None


class _ProjectFn(torch.autograd.Function):
    """projection.py:54-81, stand-alone."""

    @staticmethod
    def forward(ctx, verts, smpl, vs: int):
        lib = _lib.load()
        verts = _check_cuda_f32(verts, "verts")
        smpl = _check_cuda_f32(smpl, "smpl")
        N, V = verts.shape[0], verts.shape[1]
        if smpl.shape[0] != N or smpl.shape[1] < 4:
            raise ValueError("smpl must be (N, >=4) with the camera in columns 0..3")
        if smpl.shape[1] != NUM_PARAMS:
            raise ValueError("smpl must be (N,86) (model.py:33-35), got %s" % (tuple(smpl.shape),))
        Vs = (V + vs - 1) // vs
        with torch.cuda.device(verts.device):
            proj = torch.empty((N, Vs, 3), dtype=torch.float32, device=verts.device)
            _lib.check(lib.smpl_b200_project_fwd(_ptr(verts), _ptr(smpl), N, V, vs, _ptr(proj), _stream()),
                       "smpl_b200_project_fwd")
        ctx.vs = vs
        ctx.save_for_backward(verts, smpl)
        return proj

    @staticmethod
    def backward(ctx, g_proj):
        lib = _lib.load()
        verts, smpl = ctx.saved_tensors
        g_proj = _check_cuda_f32(g_proj, "grad projects")
        N, V = verts.shape[0], verts.shape[1]
        with torch.cuda.device(verts.device):
            g_verts = torch.empty_like(verts)
            g_smpl = torch.empty_like(smpl)
            _lib.check(lib.smpl_b200_project_bwd(_ptr(verts), _ptr(smpl), _ptr(g_proj), N, V, ctx.vs, _ptr(g_verts),
                                                 _ptr(g_smpl), _stream()), "smpl_b200_project_bwd")
        return g_verts, g_smpl, None


class _SegFn(torch.autograd.Function):
    """projects_to_seg.py:9-69."""

    @staticmethod
    def forward(ctx, pwd, mask, table: PartTable, img_wh: int):
        lib = _lib.load()
        pwd = _check_cuda_f32(pwd, "projects_with_depth")
        mask = _check_cuda_f32(mask, "mask_vals")
        N, Vs = pwd.shape[0], pwd.shape[1]
        if pwd.dim() != 3 or pwd.shape[2] != 3 or tuple(mask.shape) != (N, Vs):
            raise ValueError("expected projects (N,Vs,3) and mask (N,Vs), got %s and %s" %
                             (tuple(pwd.shape), tuple(mask.shape)))
        need_grad = bool(ctx.needs_input_grad[0])
        with torch.cuda.device(pwd.device):
            seg = torch.empty((N, img_wh, img_wh, table.P + 1), dtype=torch.float32, device=pwd.device)
            # the backward's state: clip gate + arg-min byte per (pixel, part); only produced when a backward can follow
            saved = _workspace(lib.smpl_b200_seg_saved_bytes(N, img_wh), pwd.device) if need_grad else None
            _lib.check(lib.smpl_b200_seg_fwd(table.handle, _ptr(pwd), _ptr(mask), N, Vs, img_wh, _ptr(seg), _ptr(saved),
                                             _stream()), "smpl_b200_seg_fwd")
        ctx.table, ctx.img_wh = table, img_wh
        ctx.have_state = need_grad
        if need_grad:
            ctx.save_for_backward(pwd, mask, saved)
        return seg

    @staticmethod
    def backward(ctx, g_seg):
        lib = _lib.load()
        if not ctx.have_state:                # only the mask asked for a gradient: it is a constant (back_prop=False)
            return None, None, None, None
        pwd, mask, saved = ctx.saved_tensors
        g_seg = _check_cuda_f32(g_seg, "grad seg")
        N, Vs = pwd.shape[0], pwd.shape[1]
        with torch.cuda.device(pwd.device):
            g_pwd = torch.empty_like(pwd)
            _lib.check(lib.smpl_b200_seg_bwd(ctx.table.handle, _ptr(pwd), _ptr(mask), _ptr(g_seg), _ptr(saved), N, Vs,
                                             ctx.img_wh, _ptr(g_pwd), _stream()), "smpl_b200_seg_bwd")
        return g_pwd, None, None, None       # the mask is a constant (compute_mask.py:30, back_prop=False)


class _SilFn(torch.autograd.Function):
    """projects_to_silhouette.py:14-44."""

    @staticmethod
    def forward(ctx, pwd, img_wh: int):
        lib = _lib.load()
        pwd = _check_cuda_f32(pwd, "projects_with_depth")
        N, Vs = pwd.shape[0], pwd.shape[1]
        need_grad = bool(ctx.needs_input_grad[0])
        with torch.cuda.device(pwd.device):
            sil = torch.empty((N, img_wh, img_wh, 2), dtype=torch.float32, device=pwd.device)
            # the backward's state: the arg-min vertex of every pixel (2 bytes), so it never repeats the search
            saved = _workspace(N * img_wh * img_wh * 2, pwd.device) if need_grad else None
            _lib.check(lib.smpl_b200_silhouette_fwd(_ptr(pwd), N, Vs, img_wh, _ptr(sil), _ptr(saved),
                                                    saved.numel() if saved is not None else 0, _stream()),
                       "smpl_b200_silhouette_fwd")
        ctx.img_wh = img_wh
        ctx.have_state = need_grad
        if need_grad:
            ctx.save_for_backward(pwd, saved)
        return sil

    @staticmethod
    def backward(ctx, g_sil):
        lib = _lib.load()
        if not ctx.have_state:
            return None, None
        pwd, saved = ctx.saved_tensors
        g_sil = _check_cuda_f32(g_sil, "grad silhouette")
        N, Vs = pwd.shape[0], pwd.shape[1]
        with torch.cuda.device(pwd.device):
            g_pwd = torch.empty_like(pwd)
            _lib.check(lib.smpl_b200_silhouette_bwd(_ptr(pwd), _ptr(g_sil), N, Vs, ctx.img_wh, _ptr(g_pwd), _ptr(saved),
                                                    saved.numel(), _stream()), "smpl_b200_silhouette_bwd")
        return g_pwd, None


class _FullFn(torch.autograd.Function):
    """model.py:108-118 in one C-ABI call each way (smpl_b200_full_fwd / smpl_b200_full_bwd): params (N,86) ->
    (verts, joints, projects, mask, seg).  The gradient flows from `seg` only; verts / joints / projects / mask are
    returned as non-differentiable by-products (use the modular functions to differentiate through them)."""

    @staticmethod
    def forward(ctx, params, dm: DeviceModel, table: PartTable, img_wh: int, vs: int, need_verts: bool):
        lib = _lib.load()
        params = _check_cuda_f32(params, "params")
        if params.dim() != 2 or params.shape[1] != NUM_PARAMS:
            raise ValueError("params must be (N,86), got %s" % (tuple(params.shape),))
        if params.device != dm.device:
            raise _lib.SmplB200Error("params on %s but the SMPL model is on %s" % (params.device, dm.device))
        N, dev = params.shape[0], params.device
        Vs = (dm.V + vs - 1) // vs
        need_grad = bool(ctx.needs_input_grad[0])
        with torch.cuda.device(dev):
            verts = torch.empty((N, dm.V, 3), dtype=torch.float32, device=dev) if need_verts else None
            joints = torch.empty((N, NUM_JOINTS, 3), dtype=torch.float32, device=dev)
            proj = torch.empty((N, Vs, 3), dtype=torch.float32, device=dev)
            mask = torch.empty((N, Vs), dtype=torch.float32, device=dev)
            seg = torch.empty((N, img_wh, img_wh, table.P + 1), dtype=torch.float32, device=dev)
            state = _workspace(lib.smpl_b200_full_state_bytes(dm.handle, N, img_wh, vs), dev) if need_grad else None
            ws = _workspace(dm.workspace_bytes(_lib.OP_FULL_FWD, N, img_wh, vs), dev)
            _lib.check(lib.smpl_b200_full_fwd(dm.handle, table.handle, _ptr(params), N, img_wh, vs, _ptr(verts),
                                              _ptr(joints), _ptr(proj), _ptr(mask), _ptr(seg), _ptr(state), _ptr(ws),
                                              ws.numel(), _stream()), "smpl_b200_full_fwd")
        ctx.dm, ctx.table, ctx.img_wh, ctx.vs, ctx.have_state = dm, table, img_wh, vs, need_grad
        ctx.set_materialize_grads(False)      # or autograd memsets a dense zero gradient for every unused output (1.6 GB at 16384)
        if need_grad:
            ctx.save_for_backward(params, proj, mask, state)
        ctx.mark_non_differentiable(joints, proj, mask)
        if verts is not None:
            ctx.mark_non_differentiable(verts)
        return verts, joints, proj, mask, seg

    @staticmethod
    def backward(ctx, g_verts, g_joints, g_proj, g_mask, g_seg):
        lib = _lib.load()
        if not ctx.have_state or g_seg is None:
            return None, None, None, None, None, None
        params, proj, mask, state = ctx.saved_tensors
        g_seg = _check_cuda_f32(g_seg, "grad seg")
        dm, N, dev = ctx.dm, params.shape[0], params.device
        with torch.cuda.device(dev):
            g_params = torch.empty_like(params)
            ws = _workspace(dm.workspace_bytes(_lib.OP_FULL_BWD, N, ctx.img_wh, ctx.vs), dev)
            _lib.check(lib.smpl_b200_full_bwd(dm.handle, ctx.table.handle, _ptr(params), N, ctx.img_wh, ctx.vs, _ptr(proj),
                                              _ptr(mask), _ptr(g_seg), _ptr(state), _ptr(g_params), _ptr(ws), ws.numel(),
                                              _stream()), "smpl_b200_full_bwd")
        return g_params, None, None, None, None, None


class _SegLossFn(torch.autograd.Function):
    """projects_to_seg (projects_to_seg.py:9-69) -> Reshape -> softmax (model.py:119-120) -> categorical focal loss
    (focal_loss.py:10-48) as ONE kernel each way; integer labels."""

    @staticmethod
    def forward(ctx, pwd, mask, labels, class_w, table: PartTable, img_wh: int, gamma: float, want_seg: bool):
        lib = _lib.load()
        pwd = _check_cuda_f32(pwd, "projects_with_depth")
        mask = _check_cuda_f32(mask, "mask_vals")
        N, Vs = pwd.shape[0], pwd.shape[1]
        if pwd.dim() != 3 or pwd.shape[2] != 3 or tuple(mask.shape) != (N, Vs):
            raise ValueError("expected projects (N,Vs,3) and mask (N,Vs), got %s and %s" %
                             (tuple(pwd.shape), tuple(mask.shape)))
        if labels.dtype != torch.uint8 or labels.numel() != N * img_wh * img_wh or labels.device != pwd.device:
            raise ValueError("labels must be uint8 class ids of shape (N, img_wh*img_wh) on the projections' device")
        labels = labels.contiguous()
        with torch.cuda.device(pwd.device):
            seg = torch.empty((N, img_wh, img_wh, table.P + 1), dtype=torch.float32, device=pwd.device) if want_seg else None
            loss = torch.empty((N, img_wh * img_wh), dtype=torch.float32, device=pwd.device)
            state = _workspace(lib.smpl_b200_seg_loss_state_bytes(N, img_wh), pwd.device)
            _lib.check(lib.smpl_b200_seg_loss_fwd(table.handle, _ptr(pwd), _ptr(mask), N, Vs, img_wh, _ptr(labels),
                                                  float(gamma), _ptr(class_w), _ptr(seg), _ptr(loss), _ptr(state),
                                                  _stream()), "smpl_b200_seg_loss_fwd")
        ctx.table, ctx.img_wh = table, img_wh
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(pwd, mask, state)
        if seg is not None:
            ctx.mark_non_differentiable(seg)       # a by-product here: differentiate through projects_to_seg for its own gradient
        return loss, seg

    @staticmethod
    def backward(ctx, g_loss, g_seg):
        lib = _lib.load()
        if g_loss is None:
            return None, None, None, None, None, None, None, None
        pwd, mask, state = ctx.saved_tensors
        g_loss = _check_cuda_f32(g_loss, "grad loss")
        N, Vs = pwd.shape[0], pwd.shape[1]
        with torch.cuda.device(pwd.device):
            g_pwd = torch.empty_like(pwd)
            _lib.check(lib.smpl_b200_seg_loss_bwd(ctx.table.handle, _ptr(pwd), _ptr(mask), _ptr(g_loss), _ptr(state), N, Vs,
                                                  ctx.img_wh, _ptr(g_pwd), _stream()), "smpl_b200_seg_loss_bwd")
        return g_pwd, None, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------------------
# the reference's public names
# ------------------------------------------------------------------------------------------------------------
class SMPLLayer(torch.nn.Module):
    """Drop-in for keras_smpl.batch_smpl.SMPLLayer (batch_smpl.py:23-166).

    `pkl_path` may also be a smpl_io.SmplHostModel (the real pickle is a licensed download).  `call(x)` maps
    (N,86) params to (N,6890,3) vertices and leaves the posed joints in `self.J_transformed` (batch_smpl.py:131).
    `joints(x)` additionally evaluates the cocoplus/LSP keypoint regression the reference keeps commented out
    (batch_smpl.py:147-151).
    """

    def __init__(self, pkl_path, batch_size=8, dtype="float32", joint_type="lsp", device=None, **kwargs):
        super().__init__()
        if str(dtype) not in ("float32", "torch.float32"):
            raise TypeError("only float32 is supported (reference default, batch_smpl.py:24)")
        if joint_type not in ("lsp", "cocoplus"):
            raise ValueError("joint_type must be 'lsp' or 'cocoplus'")
        self.pkl_path = pkl_path
        self.dtype_name = "float32"
        self.joint_type = joint_type
        self.batch_size = batch_size
        self.num_cam = 4                                    # batch_smpl.py:90
        self._device = None if device is None else torch.device(device)
        self._dm: Optional[DeviceModel] = None
        self.J_transformed = None
        self.trainable = False                              # batch_smpl.py:92-94: no trainable weights

    def build(self, input_shape=None, device=None):
        """batch_smpl.py:31-94.  Raises IOError (FileNotFoundError) if the pickle is missing, like the reference."""
        dev = torch.device(device) if device is not None else self._device
        if dev is None or dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if self._dm is None or self._dm.device != dev:
            self._dm = get_device_model(self.pkl_path, dev)
        self.size = [self._dm.V, 3]
        self.num_betas = 10
        self.num_joints = NUM_JOINTS
        self.num_thetas = NUM_JOINTS * 3
        return self._dm

    @property
    def num_keypoints(self) -> int:
        return 14 if self.joint_type == "lsp" else 19       # batch_smpl.py:86-87

    def _model_for(self, x: torch.Tensor) -> DeviceModel:
        if not x.is_cuda:
            raise _lib.SmplB200Error("SMPLLayer input is on %s: CUDA only, no CPU fallback" % x.device)
        if self._dm is None or self._dm.device != x.device:
            self.build(device=x.device)
        return self._dm

    def call(self, x: torch.Tensor) -> torch.Tensor:
        dm = self._model_for(x)
        verts, joints, _, _ = _DecodeFn.apply(x, dm, True, 0, 0)
        self.J_transformed = joints
        return verts

    forward = call

    def joints(self, x: torch.Tensor):
        """(verts, keypoints (N,14|19,3)): the commented-out regression of batch_smpl.py:147-151 (forward only)."""
        dm = self._model_for(x)
        verts, joints, keyp, _ = _DecodeFn.apply(x, dm, True, min(self.num_keypoints, dm.R), 0)
        self.J_transformed = joints
        return verts, keyp

    def compute_output_shape(self, input_shape):
        V = self._dm.V if self._dm is not None else smpl_io.NUM_VERTS
        return (input_shape[0], V, 3)                       # batch_smpl.py:155-159

    def get_config(self):
        return {"pkl_path": self.pkl_path, "batch_size": self.batch_size, "dtype": self.dtype_name}   # :161-166


def orthographic_project(inputs, vertex_sampling):
    """projection.py:54-81: (verts (N,V,3), smpl (N,86)) -> (N, ceil(V/vs), 3) = (u, v, z)."""
    verts, smpl = inputs
    return _ProjectFn.apply(verts, smpl, _vs(vertex_sampling))


def compute_mask(batch_projects_with_depth: torch.Tensor) -> torch.Tensor:
    """compute_mask.py:12-32: (N,Vs,3) -> (N,Vs) in {1,500}; no gradient (back_prop=False)."""
    lib = _lib.load()
    pwd = _check_cuda_f32(batch_projects_with_depth.detach(), "batch_projects_with_depth")
    if pwd.dim() != 3 or pwd.shape[2] != 3:
        raise ValueError("expected (N,Vs,3), got %s" % (tuple(pwd.shape),))
    N, Vs = pwd.shape[0], pwd.shape[1]
    with torch.cuda.device(pwd.device):
        mask = torch.empty((N, Vs), dtype=torch.float32, device=pwd.device)
        _lib.check(lib.smpl_b200_mask_fwd(_ptr(pwd), N, Vs, _ptr(mask), _stream()), "smpl_b200_mask_fwd")
    return mask


def projects_to_seg(input, img_wh, vertex_sampling, part_indices_path: Optional[str] = None, parts=None):
    """projects_to_seg.py:9-69: ([projects (N,Vs,3), mask (N,Vs)]) -> (N, wh, wh, 32), channel 0 = background,
    rows flipped.  The part table is read from the reference's CWD-relative file unless given explicitly."""
    projects_with_depth, mask_vals = input
    table = get_part_table(vertex_sampling, projects_with_depth.shape[1], projects_with_depth.device,
                           part_indices_path, parts)
    return _SegFn.apply(projects_with_depth, mask_vals, table, int(img_wh))


def projects_to_silhouette(projects_with_depth, img_wh):
    """projects_to_silhouette.py:14-44: (N,Vs,3) -> (N, wh, wh, 2) = [1-s, s], rows flipped."""
    return _SilFn.apply(projects_with_depth, int(img_wh))


_mean_cache = {}


def _mean_tensor(img_wh, device, with_smpl: bool, mean_params_path: Optional[str] = None) -> torch.Tensor:
    key = (float(img_wh), str(device), with_smpl, mean_params_path)
    t = _mean_cache.get(key)
    if t is None:
        if with_smpl:
            m = smpl_io.mean_param_vector(img_wh, smpl_io.load_mean_params(mean_params_path))
        else:
            m = np.zeros((1, NUM_PARAMS))
            m[0, :4] = [img_wh / 2.0, img_wh / 2.0, img_wh / 2.0, img_wh / 1.6]
        t = torch.as_tensor(m.astype(np.float32), device=device)       # tf.constant(..., dtype='float32')
        _mean_cache[key] = t
    return t


def concat_mean_param(img_features, img_wh, mean_params_path: Optional[str] = None):
    """concat_mean_param.py:8-31: (N,F) -> (N,F+86) = [features | mean params]."""
    mean = _mean_tensor(img_wh, img_features.device, True, mean_params_path)
    return torch.cat([img_features, mean.expand(img_features.shape[0], -1)], dim=1)


def set_cam_params(smpl, img_wh):
    """set_cam_params.py:13-26: adds the camera initialisation [wh/2, wh/2, wh/2, wh/1.6] to columns 0..3."""
    return smpl + _mean_tensor(img_wh, smpl.device, False)


def load_mean_set_cam_params(smpl, img_wh, mean_params_path: Optional[str] = None):
    """set_cam_params.py:29-52: adds the mean pose (global rotation zeroed), mean shape and camera initialisation."""
    return smpl + _mean_tensor(img_wh, smpl.device, True, mean_params_path)


# ------------------------------------------------------------------------------------------------------------
# the op right after the path: softmax (model.py:119-120) + categorical focal loss
# ------------------------------------------------------------------------------------------------------------
def focal_class_weights(num_classes: int = 32) -> np.ndarray:
    """focal_loss.py:21-38."""
    w = np.ones(num_classes, np.float32)
    w[0] = 0.3
    for c in (1, 2, 3, 4, 10, 12, 14, 15, 16, 17, 23, 25):
        if c < num_classes:
            w[c] = 2.0
    return w


class _FocalFn(torch.autograd.Function):
    """seg scores (N, P, C) + labels -> per-pixel focal loss (N, P); focal_loss.py:12-46 (+ softmax, model.py:120)."""

    @staticmethod
    def forward(ctx, seg, y_true, labels, gamma: float, class_w, from_logits: bool):
        lib = _lib.load()
        seg = _check_cuda_f32(seg, "y_pred")
        C_ = seg.shape[-1]
        npix = seg.numel() // C_
        with torch.cuda.device(seg.device):
            loss = torch.empty(seg.shape[:-1], dtype=torch.float32, device=seg.device)
            _lib.check(lib.smpl_b200_focal_loss_fwd(_ptr(seg), _ptr(y_true), _ptr(labels), npix, C_, float(gamma),
                                                    _ptr(class_w), int(from_logits), _ptr(loss), _stream()),
                       "smpl_b200_focal_loss_fwd")
        ctx.gamma, ctx.from_logits = float(gamma), bool(from_logits)
        ctx.save_for_backward(seg, y_true if y_true is not None else labels, class_w if class_w is not None else seg.new_empty(0))
        ctx.soft = y_true is not None
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        lib = _lib.load()
        seg, lab, class_w = ctx.saved_tensors
        g_loss = _check_cuda_f32(g_loss, "grad loss")
        C_ = seg.shape[-1]
        npix = seg.numel() // C_
        with torch.cuda.device(seg.device):
            g_seg = torch.empty_like(seg)
            _lib.check(lib.smpl_b200_focal_loss_bwd(_ptr(seg), _ptr(lab) if ctx.soft else None,
                                                    None if ctx.soft else _ptr(lab), _ptr(g_loss), npix, C_, ctx.gamma,
                                                    _ptr(class_w) if class_w.numel() else None, int(ctx.from_logits),
                                                    _ptr(g_seg), _stream()), "smpl_b200_focal_loss_bwd")
        return g_seg, None, None, None, None, None


def categorical_focal_loss(gamma=2.0, weight_classes=False, from_logits=False):
    """focal_loss.py:10-48, same factory signature and the same meaning of ``y_pred``: probabilities, the output of
    ``Activation('softmax')`` (model.py:119-120).  The returned ``loss(y_true, y_pred)`` gives the per-pixel loss
    (N, img_wh^2) like the reference.  ``from_logits=True`` is the opt-in fused variant: it takes the rasteriser's raw
    scores and applies the softmax of model.py:120 inside the kernel.  ``y_true`` is the reference's one-hot / soft
    (N, img_wh^2, C) float tensor, or integer class ids (N, img_wh^2) to save the 128-byte label row per pixel."""

    def categorical_focal_loss_fixed(y_true, y_pred):
        seg = _check_cuda_f32(y_pred, "y_pred")
        if seg.dim() == 4:                                       # (N, wh, wh, C): the Reshape of model.py:119
            seg = seg.reshape(seg.shape[0], -1, seg.shape[-1])
        C_ = seg.shape[-1]
        if y_true.dtype.is_floating_point:
            yt = _check_cuda_f32(y_true.reshape(seg.shape), "y_true")
            lab = None
        else:
            yt = None
            lab = y_true.reshape(seg.shape[:-1]).to(device=seg.device, dtype=torch.uint8).contiguous()
        cw = torch.as_tensor(focal_class_weights(C_), device=seg.device) if weight_classes else None
        return _FocalFn.apply(seg, yt, lab, gamma, cw, from_logits)

    return categorical_focal_loss_fixed


def projects_to_seg_focal_loss(input, labels, img_wh, vertex_sampling, gamma=2.0, weight_classes=False,
                               part_indices_path: Optional[str] = None, parts=None, return_seg: bool = False):
    """The tail of every training graph of the reference as one op: ``projects_to_seg`` (projects_to_seg.py:9-69) ->
    ``Reshape((img_wh*img_wh, 32))`` -> ``Activation('softmax')`` (model.py:119-120) -> ``categorical_focal_loss(gamma,
    weight_classes)`` (focal_loss.py:10-48), for integer class ids ``labels`` (N, img_wh*img_wh) [or (N, img_wh, img_wh)]
    in the order of ``y_true``'s pixels.  Returns the per-pixel loss (N, img_wh*img_wh) like the reference's loss
    function (and the segmentation, detached, if ``return_seg``).  Equal to
    ``categorical_focal_loss(gamma, weight_classes, from_logits=True)(labels, projects_to_seg(input, ...))`` without ever
    writing the (N, wh, wh, 32) scores or reading an upstream gradient of that shape."""
    projects_with_depth, mask_vals = input
    table = get_part_table(vertex_sampling, projects_with_depth.shape[1], projects_with_depth.device,
                           part_indices_path, parts)
    N = projects_with_depth.shape[0]
    lab = labels.reshape(N, -1).to(device=projects_with_depth.device, dtype=torch.uint8)
    cw = torch.as_tensor(focal_class_weights(table.P + 1), device=projects_with_depth.device) if weight_classes else None
    loss, seg = _SegLossFn.apply(projects_with_depth, mask_vals, lab, cw, table, int(img_wh), float(gamma), bool(return_seg))
    return (loss, seg) if return_seg else loss


def categorical_crossentropy(y_true, y_pred, from_logits=False):
    """The silhouette branch's loss (train_stage2_silhouette.py:226-228, Keras 'categorical_crossentropy' on the
    softmax(2) of the silhouette): -sum_c y_c log clip(p_c, eps, 1-eps) per pixel, i.e. the focal kernel with gamma = 0
    and no class weights.  ``y_pred`` holds probabilities like Keras'; ``from_logits=True`` (opt-in) fuses the softmax
    of train_stage2_silhouette.py:85-86."""
    return categorical_focal_loss(gamma=0.0, weight_classes=False, from_logits=from_logits)(y_true, y_pred)


# ------------------------------------------------------------------------------------------------------------
# the whole path as one module (model.py:108-118 tail; + train_stage2_silhouette.py:84 silhouette branch)
# ------------------------------------------------------------------------------------------------------------
class SmplDecoder(torch.nn.Module):
    """params (N,86) -> dict(seg, [silhouette], projects, mask, [verts], joints): the decoder tail of
    model.py:108-118 with the projection fused into the skinning kernel."""

    def __init__(self, pkl_path, img_wh: int, vertex_sampling=None, silhouette_wh: Optional[int] = None,
                 need_verts: bool = True, parts=None, part_indices_path: Optional[str] = None, device=None,
                 fused: bool = False):
        """fused=True runs the seg path as ONE C-ABI call each way (smpl_b200_full_fwd / _bwd): the same kernels with
        less host work and a 16.5 KB instead of 82.9 KB per-sample backward state; the gradient then flows from
        out["seg"] only (verts / joints / projects / mask come back detached)."""
        super().__init__()
        self.fused = bool(fused)
        self.smpl = SMPLLayer(pkl_path, device=device)
        self.img_wh = int(img_wh)
        self.vertex_sampling = vertex_sampling
        self.silhouette_wh = silhouette_wh
        self.need_verts = need_verts
        self._parts, self._parts_path = parts, part_indices_path

    def forward(self, params: torch.Tensor, seg: bool = True):
        dm = self.smpl._model_for(params)
        vs = _vs(self.vertex_sampling)
        if self.fused and seg and not self.silhouette_wh:
            Vs = (dm.V + vs - 1) // vs
            table = get_part_table(self.vertex_sampling, Vs, params.device, self._parts_path, self._parts)
            verts, joints, proj, mask, segm = _FullFn.apply(params, dm, table, self.img_wh, vs, self.need_verts)
            self.smpl.J_transformed = joints
            return {"verts": verts, "joints": joints, "projects": proj, "mask": mask, "seg": segm}
        verts, joints, _, proj = _DecodeFn.apply(params, dm, self.need_verts, 0, vs)
        self.smpl.J_transformed = joints
        out = {"verts": verts, "joints": joints, "projects": proj}
        if seg:
            mask = compute_mask(proj)
            table = get_part_table(self.vertex_sampling, proj.shape[1], proj.device, self._parts_path, self._parts)
            out["mask"] = mask
            out["seg"] = _SegFn.apply(proj, mask, table, self.img_wh)
        if self.silhouette_wh:
            out["silhouette"] = _SilFn.apply(proj, int(self.silhouette_wh))
        return out

    def focal_loss(self, params: torch.Tensor, labels: torch.Tensor, gamma: float = 2.0, weight_classes: bool = False):
        """The decoder's training objective in one call: params (N,86) -> per-pixel categorical focal loss (N, wh*wh)
        against integer labels (model.py:108-120 + focal_loss.py), with the segmentation -> softmax -> loss tail fused
        into the rasteriser (projects_to_seg_focal_loss).  Returns dict(loss, projects, mask, joints)."""
        dm = self.smpl._model_for(params)
        vs = _vs(self.vertex_sampling)
        _, joints, _, proj = _DecodeFn.apply(params, dm, False, 0, vs)
        self.smpl.J_transformed = joints
        mask = compute_mask(proj)
        loss = projects_to_seg_focal_loss([proj, mask], labels, self.img_wh, self.vertex_sampling, gamma, weight_classes,
                                          self._parts_path, self._parts)
        return {"loss": loss, "projects": proj, "mask": mask, "joints": joints}


class GraphedDecoderStep:
    """One training-shaped step of a SmplDecoder -- forward, then backward from an upstream gradient of the segmentation's
    shape -- at a FIXED batch, captured once as a CUDA graph and replayed: one graph launch per step instead of a dozen
    kernel launches, eight tensor-map encodes, the allocator and autograd bookkeeping.  That host work is invisible behind
    7 ms of kernels at 16384 samples but not behind the ~1 ms of a 2048-sample shard (one 16384 batch over 8 GPUs).

        step = GraphedDecoderStep(decoder, batch)
        g_params = step(params, g_seg)          # copies into the static inputs, replays, returns the static gradient
        step.replay()                           # inputs already in step.params / step.g_seg

    The arithmetic is the eager path's own kernels in the same order; outputs of the last replay are in ``step.out``.
    """

    def __init__(self, decoder: "SmplDecoder", batch: int, device=None, warmup: int = 3, micro_batches: int = 1,
                 g_seg: Optional[torch.Tensor] = None):
        """micro_batches > 1: the batch is cut into that many contiguous slices and each slice's forward + backward is
        captured on its own stream (fork / join inside the one graph).  Samples are independent, so the arithmetic of a
        slice is unchanged; what changes is that the last, partly filled wave of blocks of one slice's kernel runs beside
        the other slices' kernels instead of beside idle SMs -- which matters for the ~1 ms step of a 2048-sample shard
        (one-sample blocks: 3.5 / 4.6 waves of the two seg kernels) and not for 16384 samples."""
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.decoder, self.batch, self.device = decoder, int(batch), dev
        self.micro_batches = max(1, min(int(micro_batches), self.batch))
        self._bounds = [(self.batch * i // self.micro_batches, self.batch * (i + 1) // self.micro_batches)
                        for i in range(self.micro_batches)]
        dm = decoder.smpl.build(device=dev)
        wh = decoder.img_wh
        with torch.cuda.device(dev):
            self.params = torch.zeros((self.batch, NUM_PARAMS), dtype=torch.float32, device=dev)
            self.params[:, :4] = torch.tensor([wh / 2.0, wh / 2.0, wh / 2.0, wh / 1.6], device=dev)
            # g_seg: an existing (batch, wh, wh, 32) buffer to use as the static upstream gradient (shared between steps)
            self.g_seg = g_seg if g_seg is not None else torch.zeros((self.batch, wh, wh, 32), dtype=torch.float32, device=dev)
            self.g_params = torch.zeros((self.batch, NUM_PARAMS), dtype=torch.float32, device=dev)
            self._streams = [torch.cuda.Stream(device=dev) for _ in range(self.micro_batches - 1)]
            _lib.profile_enable(False)                               # event pairs cannot be recorded into a capture
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(max(warmup, 1)):                      # allocator / lazy-init warm-up, as torch's graph recipe asks
                    self._eager()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            n0 = _lib.launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.outs = self._eager()
            self.launches_per_step = _lib.launch_count() - n0       # kernels of this library inside one replay
        del dm

    def _slice_step(self, a: int, b: int):
        x = self.params[a:b].detach().requires_grad_(True)
        out = self.decoder(x)
        out["seg"].backward(self.g_seg[a:b])                        # autograd runs a node's backward on its forward's stream
        self.g_params[a:b].copy_(x.grad)
        return {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}

    def _eager(self):
        """One step: slice 0 on the current stream, the others on their own streams between a fork and a join."""
        cur = torch.cuda.current_stream(self.device)
        for st in self._streams:
            st.wait_stream(cur)
        outs = [None] * self.micro_batches
        for i, st in enumerate(self._streams, start=1):
            with torch.cuda.stream(st):
                outs[i] = self._slice_step(*self._bounds[i])
        outs[0] = self._slice_step(*self._bounds[0])
        for st in self._streams:
            cur.wait_stream(st)
        return outs

    @property
    def out(self) -> dict:
        """Outputs of the last replay (tensors of the slices concatenated on demand; not part of the step)."""
        if self.micro_batches == 1:
            return self.outs[0]
        return {k: (torch.cat([o[k] for o in self.outs]) if isinstance(v, torch.Tensor) else v) for k, v in self.outs[0].items()}

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.g_params

    def __call__(self, params: Optional[torch.Tensor] = None, g_seg: Optional[torch.Tensor] = None) -> torch.Tensor:
        if params is not None:
            self.params.copy_(params, non_blocking=True)
        if g_seg is not None:
            self.g_seg.copy_(g_seg, non_blocking=True)
        return self.replay()


class PipelinedDecoderSteps:
    """Host-fed training steps with the copies off the critical path: two GraphedDecoderSteps (sharing one upstream-gradient
    buffer) alternate, an upload stream copies step k+1's parameters from pinned host memory while step k's graph runs, and a
    download stream copies step k's parameter gradient back while step k+1 runs.  Every step still does its own H2D and D2H.

        pipe = PipelinedDecoderSteps(decoder, batch, micro_batches=2)
        for params_host, grad_host in batches:          # pinned host tensors (batch, 86)
            pipe.step(params_host, grad_host)            # returns at once; grad_host is valid after pipe.synchronize()
        pipe.synchronize()

    A caller that reuses ONE grad_host must consume it between steps (the bench only measures the traffic)."""

    def __init__(self, decoder: "SmplDecoder", batch: int, device=None, micro_batches: int = 1,
                 g_seg: Optional[torch.Tensor] = None, first: Optional["GraphedDecoderStep"] = None):
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        a = first if first is not None else GraphedDecoderStep(decoder, batch, device=dev, micro_batches=micro_batches, g_seg=g_seg)
        b = GraphedDecoderStep(decoder, batch, device=dev, micro_batches=micro_batches, g_seg=a.g_seg)
        self.steps = [a, b]
        self.g_seg = a.g_seg
        with torch.cuda.device(dev):
            self._up, self._down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            self._uploaded = [torch.cuda.Event() for _ in range(2)]
            self._computed = [torch.cuda.Event() for _ in range(2)]
            self._downloaded = [torch.cuda.Event() for _ in range(2)]
            cur = torch.cuda.current_stream(dev)
            for e in self._computed + self._downloaded:
                e.record(cur)
        self._k = 0

    def step(self, params_host: torch.Tensor, grad_host: torch.Tensor) -> None:
        i = self._k & 1
        self._k += 1
        s = self.steps[i]
        cur = torch.cuda.current_stream(self.device)
        # upload: the graph that last read s.params (two steps ago) and the download of ITS gradient must be done
        self._up.wait_event(self._computed[i])
        with torch.cuda.stream(self._up):
            s.params.copy_(params_host, non_blocking=True)
            self._uploaded[i].record(self._up)
        cur.wait_event(self._uploaded[i])
        cur.wait_event(self._downloaded[i])              # s.g_params is about to be rewritten
        s.replay()
        self._computed[i].record(cur)
        self._down.wait_event(self._computed[i])
        with torch.cuda.stream(self._down):
            grad_host.copy_(s.g_params, non_blocking=True)
            self._downloaded[i].record(self._down)

    def join(self) -> None:
        """Make the current stream wait for the copies in flight (no host synchronisation)."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self._up)
        cur.wait_stream(self._down)

    def synchronize(self) -> None:
        self.join()
        torch.cuda.current_stream(self.device).synchronize()
