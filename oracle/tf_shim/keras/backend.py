"""keras.backend (TensorFlow backend) symbols used by the reference, on the shim's Tensor."""
import numpy as np
import torch

import tensorflow as tf
from tensorflow import Tensor, _T, _dtype, _ints


def epsilon():
    return 1e-7


def variable(value, dtype=None, name=None):
    a = np.asarray(value)
    d = _dtype(dtype) or torch.float32                   # K.floatx() == 'float32'
    return Tensor(torch.from_numpy(np.ascontiguousarray(a)).to(d))


def eye(size, dtype=None, name=None):
    return Tensor(torch.eye(int(size), dtype=_dtype(dtype) or torch.float32))


def reshape(x, shape):
    return tf.reshape(x, shape)


def dot(x, y):
    tx, ty = _T(x), _T(y)
    assert tx.dim() == 2 and ty.dim() == 2               # the reference only takes 2-D x 2-D products
    return Tensor(torch.matmul(tx, ty))


def stack(x, axis=0):
    return tf.stack(x, axis=axis)


def expand_dims(x, axis=-1):
    return tf.expand_dims(x, axis)


def cos(x):
    return Tensor(torch.cos(_T(x)))


def sin(x):
    return Tensor(torch.sin(_T(x)))


def shape(x):
    return [int(s) for s in _T(x).shape]


def clip(x, min_value, max_value):
    return tf.clip_by_value(x, min_value, max_value)


def log(x):
    return Tensor(torch.log(_T(x)))


def pow(x, a):  # noqa: A001
    return Tensor(torch.pow(_T(x), a))


def sum(x, axis=None, keepdims=False):  # noqa: A001
    return tf.reduce_sum(x, axis=axis, keep_dims=keepdims)
