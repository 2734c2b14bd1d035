"""The regression module that feeds the decoder: model.py:63-105 of the reference.

  IEFRegressor    use_IEF=True  (model.py:63-97): three passes through the SHARED layers Dense(1024, relu) -> Dense(1024, relu)
                  -> Dense(86, linear) on the state [img_features (2048) | params (86)], params += scaledown * delta
  PlainRegressor  use_IEF=False (model.py:99-105): Dense(2048, relu) -> Dense(1024, relu) -> Dense(86) -> * scaledown ->
                  load_mean_set_cam_params

Every Dense product (forward, and both backward products) is a 3xTF32 tcgen05 GEMM of libsmpl_b200 (fp32 accuracy, like
the reference's fp32 Keras layers); torch only concatenates / adds the 86-wide parameter rows between them.  Weights use
the Keras layout: kernel (in, out), bias (out,), glorot-uniform / zeros initialisation (Keras defaults), so a trained
.hdf5's arrays load as they are.  CUDA only, no fallback.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from .layers import _check_cuda_f32, _ptr, _stream, _workspace, concat_mean_param, load_mean_set_cam_params


def _rows(t: torch.Tensor, name: str) -> torch.Tensor:
    """2-D fp32 CUDA tensor whose rows are contiguous (a column block of a wider buffer is fine)."""
    if not t.is_cuda:
        raise _lib.SmplB200Error("%s is on %s: CUDA only, no CPU fallback" % (name, t.device))
    if t.dtype != torch.float32 or t.dim() != 2:
        raise TypeError("%s must be a 2-D float32 tensor" % name)
    return t if t.stride(1) == 1 and t.stride(0) >= t.shape[1] else t.contiguous()


class _DenseFn(torch.autograd.Function):
    """Keras Dense: y = act(x @ kernel + bias)."""

    @staticmethod
    def forward(ctx, x, kernel, bias, relu: bool):
        lib = _lib.load()
        x = _rows(x, "x")
        kernel = _check_cuda_f32(kernel, "kernel")
        bias = None if bias is None else _check_cuda_f32(bias, "bias")
        M, fin = x.shape
        if kernel.shape[0] != fin:
            raise ValueError("kernel is %s but the input has %d features" % (tuple(kernel.shape), fin))
        fout = kernel.shape[1]
        with torch.cuda.device(x.device):
            y = torch.empty((M, fout), dtype=torch.float32, device=x.device)
            ws = _workspace(lib.smpl_b200_dense_workspace_bytes(M, fin, fout), x.device)
            _lib.check(lib.smpl_b200_dense_fwd(_ptr(x), x.stride(0), _ptr(kernel), _ptr(bias), M, fin, fout, int(relu), _ptr(y),
                                               fout, _ptr(ws), ws.numel(), _stream()), "smpl_b200_dense_fwd")
        ctx.relu = bool(relu)
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, kernel, y if relu else x.new_empty(0))
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x, kernel, y = ctx.saved_tensors
        gy = _rows(gy, "grad y")
        M, fin = x.shape
        fout = kernel.shape[1]
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        with torch.cuda.device(x.device):
            gx = torch.empty((M, fin), dtype=torch.float32, device=x.device) if need_x else None
            gw = torch.empty_like(kernel) if need_w else None
            gb = torch.empty((fout,), dtype=torch.float32, device=x.device) if need_b else None
            ws = _workspace(lib.smpl_b200_dense_workspace_bytes(M, fin, fout), x.device)
            _lib.check(lib.smpl_b200_dense_bwd(_ptr(x), x.stride(0), _ptr(kernel), _ptr(y) if ctx.relu else None, fout, _ptr(gy),
                                               gy.stride(0), M, fin, fout, int(ctx.relu), _ptr(gx), fin, _ptr(gw), _ptr(gb), 0,
                                               _ptr(ws), ws.numel(), _stream()), "smpl_b200_dense_bwd")
        return gx, gw, gb, None


class Dense(torch.nn.Module):
    """keras.layers.Dense(units, activation) on the tensor cores; `kernel` is (in, out) as Keras stores it."""

    def __init__(self, in_features: int, units: int, activation: str = "linear", device=None):
        super().__init__()
        if activation not in ("relu", "linear", None):
            raise ValueError("activation must be 'relu' or 'linear'")
        self.activation = activation or "linear"
        limit = math.sqrt(6.0 / (in_features + units))              # glorot_uniform, the Keras default
        self.kernel = torch.nn.Parameter((torch.rand(in_features, units, device=device) * 2 - 1) * limit)
        self.bias = torch.nn.Parameter(torch.zeros(units, device=device))

    def forward(self, x):
        return _DenseFn.apply(x, self.kernel, self.bias, self.activation == "relu")


class IEFRegressor(torch.nn.Module):
    """model.py:63-97: iterative error feedback from 2048 image features to the 86 decoder parameters."""

    def __init__(self, img_wh, num_features: int = 2048, scaledown: float = 0.005, iterations: int = 3, device=None,
                 mean_params_path=None):
        super().__init__()
        self.img_wh, self.scaledown, self.iterations = img_wh, float(scaledown), int(iterations)
        self.mean_params_path = mean_params_path
        self.IEF_layer_1 = Dense(num_features + 86, 1024, "relu", device)          # :66
        self.IEF_layer_2 = Dense(1024, 1024, "relu", device)                       # :67
        self.IEF_layer_3 = Dense(1024, 86, "linear", device)                       # :68

    def forward(self, img_features):
        feat = _check_cuda_f32(img_features, "img_features")
        state = concat_mean_param(feat, self.img_wh, self.mean_params_path)         # :70-71
        param = state[:, feat.shape[1]:]                                            # :72
        for _ in range(self.iterations):                                            # :77-97 (shared layers)
            delta = self.IEF_layer_3(self.IEF_layer_2(self.IEF_layer_1(state)))
            param = param + self.scaledown * delta                                  # Lambda x*d, Add
            state = torch.cat([feat, param], dim=1)                                 # Concatenate
        return param


class PlainRegressor(torch.nn.Module):
    """model.py:99-105: Dense(2048, relu) -> Dense(1024, relu) -> Dense(86) -> * scaledown -> + mean / camera init."""

    def __init__(self, img_wh, num_features: int = 2048, scaledown: float = 0.005, device=None, mean_params_path=None):
        super().__init__()
        self.img_wh, self.scaledown, self.mean_params_path = img_wh, float(scaledown), mean_params_path
        self.dense_1 = Dense(num_features, 2048, "relu", device)
        self.dense_2 = Dense(2048, 1024, "relu", device)
        self.dense_3 = Dense(1024, 86, "linear", device)

    def forward(self, img_features):
        smpl = self.dense_3(self.dense_2(self.dense_1(_check_cuda_f32(img_features, "img_features")))) * self.scaledown
        return load_mean_set_cam_params(smpl, self.img_wh, self.mean_params_path)
