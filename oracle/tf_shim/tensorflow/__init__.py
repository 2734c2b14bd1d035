"""Stand-in for the `tensorflow` 1.x symbols the reference's decoder path uses, on torch-CPU tensors.
TEST INFRASTRUCTURE ONLY (see ../README.md).  Semantics follow the TF 1.x API documentation."""
import contextlib

import numpy as np
import torch

float32 = "float32"
int32 = "int32"

_DT = {"float32": torch.float32, "float64": torch.float64, "int32": torch.int32, "int64": torch.int64,
       None: None}


def _dtype(d):
    if isinstance(d, torch.dtype):
        return d
    return _DT[str(d) if d is not None else None]


class Dimension(object):
    def __init__(self, v):
        self.value = int(v)

    def __int__(self):
        return self.value

    __index__ = __int__

    def __eq__(self, o):
        return self.value == int(o)

    def __hash__(self):
        return hash(self.value)

    def __repr__(self):
        return "Dimension(%d)" % self.value


class TensorShape(object):
    def __init__(self, dims):
        self.dims = [Dimension(d) for d in dims]

    def __getitem__(self, i):
        return self.dims[i]

    def __len__(self):
        return len(self.dims)

    def __iter__(self):
        return iter(self.dims)

    def as_list(self):
        return [d.value for d in self.dims]

    def __repr__(self):
        return "TensorShape(%r)" % (self.as_list(),)


def _raw(x, like=None):
    """torch view of a Tensor / numpy / python value (python floats become fp32 scalars, like TF constants)."""
    if isinstance(x, Tensor):
        return x.t
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, Dimension):
        return int(x)
    if isinstance(x, (np.ndarray, np.matrix, list, tuple)):
        a = np.asarray(x)
        t = torch.from_numpy(np.ascontiguousarray(a))
        if like is not None and t.dtype.is_floating_point:
            t = t.to(like.dtype)
        return t
    return x


class Tensor(object):
    __array_priority__ = 1000

    def __init__(self, t):
        assert isinstance(t, torch.Tensor), type(t)
        self.t = t

    # -- static shape API -------------------------------------------------------------------------
    @property
    def shape(self):
        return TensorShape(self.t.shape)

    def get_shape(self):
        return self.shape

    @property
    def dtype(self):
        return str(self.t.dtype).replace("torch.", "")

    # -- indexing / arithmetic ----------------------------------------------------------------------
    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        idx = tuple(int(i) if isinstance(i, (Dimension, np.integer)) else i for i in idx)
        return Tensor(self.t[idx])

    def _bin(self, o, f, rev=False):
        a, b = self.t, _raw(o, self.t)
        if isinstance(b, torch.Tensor) and b.dtype != a.dtype and a.dtype.is_floating_point:
            b = b.to(a.dtype)
        return Tensor(f(b, a) if rev else f(a, b))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, lambda x, y: torch.sub(torch.as_tensor(x, dtype=y.dtype), y), True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __neg__(self): return Tensor(-self.t)

    def numpy(self):
        return self.t.detach().numpy()


def _T(x, like=None):
    r = _raw(x, like)
    if not isinstance(r, torch.Tensor):
        r = torch.as_tensor(r)
    return r


def _ints(seq):
    return [int(_raw(s)) if not isinstance(s, (list, tuple)) else _ints(s) for s in seq]


# ---- construction ---------------------------------------------------------------------------------
def constant(value, dtype=None, shape=None, name=None):
    a = np.asarray(value)
    d = _dtype(dtype)
    if d is None:
        d = torch.float32 if a.dtype.kind == "f" else torch.int32
    return Tensor(torch.as_tensor(a).to(d))


def ones(shape, dtype="float32", name=None):
    return Tensor(torch.ones(_ints(shape), dtype=_dtype(dtype)))


def zeros(shape, dtype="float32", name=None):
    return Tensor(torch.zeros(_ints(shape), dtype=_dtype(dtype)))


def ones_like(x, dtype=None):
    return Tensor(torch.ones_like(_T(x), dtype=_dtype(dtype)))


def eye(n, dtype="float32"):
    return Tensor(torch.eye(int(n), dtype=_dtype(dtype)))


def range(start, limit=None, delta=1, dtype=None, name=None):  # noqa: A001  (tf.range)
    if limit is None:
        start, limit = 0, start
    d = _dtype(dtype) or torch.int32
    return Tensor(torch.arange(int(start), int(limit), int(delta)).to(d))


def meshgrid(*args, **kw):
    assert kw.get("indexing", "xy") == "xy"
    a, b = [_T(x) for x in args]
    g = torch.meshgrid(a, b, indexing="xy")           # TF default is 'xy' (Cartesian) indexing, like numpy
    return [Tensor(x) for x in g]


def cast(x, dtype, name=None):
    d = _dtype(dtype)
    t = _T(x)
    if t.dtype.is_floating_point and not d.is_floating_point:
        t = torch.trunc(t)                               # float -> int casts truncate toward zero
    return Tensor(t.to(d))


# ---- shape manipulation ----------------------------------------------------------------------------
def reshape(x, shape, name=None):
    return Tensor(_T(x).reshape(_ints(shape)))


def expand_dims(x, axis=None, name=None, dim=None):
    return Tensor(_T(x).unsqueeze(int(axis if axis is not None else dim)))


def squeeze(x, axis=None, name=None):
    t = _T(x)
    return Tensor(t.squeeze() if axis is None else t.squeeze(axis))


def tile(x, multiples, name=None):
    return Tensor(_T(x).repeat(*_ints(multiples)))


def stack(values, axis=0, name=None):
    return Tensor(torch.stack([_T(v) for v in values], dim=axis))


def concat(values, axis, name=None):
    ts = [_T(v) for v in values]
    return Tensor(torch.cat(ts, dim=axis))


def pad(x, paddings, mode="CONSTANT", name=None):
    t = _T(x)
    flat = []
    for lo, hi in reversed(paddings):                    # torch pads from the last dimension backwards
        flat += [int(lo), int(hi)]
    return Tensor(torch.nn.functional.pad(t, flat))


def reverse(x, axis, name=None):
    return Tensor(torch.flip(_T(x), dims=list(axis)))


def gather(params, indices, axis=0, name=None):
    p, i = _T(params), _T(indices).long()
    out = torch.index_select(p, axis, i.reshape(-1))
    return Tensor(out.reshape(list(p.shape[:axis]) + list(i.shape) + list(p.shape[axis + 1:])))


def where(condition, x=None, y=None, name=None):
    assert x is None and y is None
    return Tensor(torch.nonzero(_T(condition)))          # (?, rank) int64 coordinates, row-major order


def size(x, name=None):
    return Tensor(torch.tensor(_T(x).numel(), dtype=torch.int32))


def unique(x, name=None):
    t = _T(x)
    vals, first = np.unique(t.numpy(), return_index=True)
    order = np.argsort(first)                            # TF keeps first-occurrence order
    y = torch.from_numpy(vals[order]).to(t.dtype)
    return Tensor(y), None


def scatter_nd(indices, updates, shape, name=None):
    idx, upd = _T(indices).long(), _T(updates)
    assert idx.shape[-1] == 1 and len(shape) == 1
    out = torch.zeros(_ints(shape), dtype=upd.dtype)
    return Tensor(out.index_add(0, idx.reshape(-1), upd))   # duplicate indices accumulate


def scatter_update(ref, indices, updates, name=None):
    out = _T(ref).clone()
    out[_T(indices).long()] = _T(updates).to(out.dtype)
    return Tensor(out)


# ---- math ----------------------------------------------------------------------------------------
def add(a, b, name=None):
    if isinstance(a, Tensor):
        return a + b
    return Tensor(torch.as_tensor(a, dtype=_T(b).dtype)) + b


def subtract(a, b, name=None):
    if isinstance(a, Tensor):
        return a - b
    return Tensor(torch.as_tensor(a, dtype=_T(b).dtype)) - b


def multiply(a, b, name=None):
    return Tensor(_T(a)) * b


def div(a, b, name=None):
    return Tensor(_T(a)) / b


def negative(x, name=None):
    return Tensor(-_T(x))


def scalar_mul(s, x):
    return Tensor(_T(x)) * s


def exp(x, name=None):
    return Tensor(torch.exp(_T(x)))


def round(x, name=None):  # noqa: A001  (tf.round: half to even)
    return Tensor(torch.round(_T(x)))


def norm(x, ord="euclidean", axis=None, keep_dims=False, name=None):  # noqa: A002
    assert ord == "euclidean"
    t = _T(x)
    return Tensor(torch.sqrt(torch.sum(t * t, dim=axis, keepdim=keep_dims)))   # tf.norm: sqrt(reduce_sum(x * conj(x)))


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    ta, tb = _T(a), _T(b)
    if transpose_a:
        ta = ta.transpose(-1, -2)
    if transpose_b:
        tb = tb.transpose(-1, -2)
    return Tensor(torch.matmul(ta, tb))


def reduce_max(x, axis=None, keep_dims=False, name=None):
    t = _T(x)
    return Tensor(torch.amax(t, dim=axis, keepdim=keep_dims))     # amax's gradient is split evenly among ties, like TF's


def reduce_sum(x, axis=None, keep_dims=False, name=None):
    t = _T(x)
    return Tensor(torch.sum(t, dim=axis, keepdim=keep_dims))


def reduce_all(x, axis=None, keep_dims=False, name=None):
    return Tensor(torch.all(_T(x), dim=axis, keepdim=keep_dims))


def equal(a, b, name=None):
    ta = _T(a)
    tb = _T(b, ta)
    if isinstance(tb, torch.Tensor) and tb.dtype != ta.dtype:
        tb = tb.to(ta.dtype)
    return Tensor(torch.eq(ta, tb))


def argmax(x, axis=None, name=None, dimension=None):
    t = _T(x)
    a = axis if axis is not None else dimension
    # TF returns the smallest index among equal maxima
    mx = torch.amax(t, dim=a, keepdim=True)
    first = torch.argmax((t == mx).to(torch.uint8), dim=a)
    return Tensor(first)


def clip_by_value(x, clip_value_min, clip_value_max, name=None):
    return Tensor(torch.clamp(_T(x), float(clip_value_min), float(clip_value_max)))   # gradient passes on the closed interval


# ---- control flow ----------------------------------------------------------------------------------
def cond(pred, true_fn=None, false_fn=None, name=None, fn1=None, fn2=None):
    return (true_fn or fn1)() if bool(_T(pred)) else (false_fn or fn2)()


def map_fn(fn, elems, dtype=None, parallel_iterations=10, back_prop=True, swap_memory=False, infer_shape=True,
           name=None):
    if isinstance(elems, (list, tuple)):
        n = _T(elems[0]).shape[0]
        outs = [fn([Tensor(_T(e)[i]) for e in elems]) for i in builtins_range(n)]
    else:
        n = _T(elems).shape[0]
        outs = [fn(Tensor(_T(elems)[i])) for i in builtins_range(n)]
    out = torch.stack([_T(o) for o in outs], dim=0)
    if not back_prop:
        out = out.detach()
    if dtype is not None:
        out = out.to(_dtype(dtype))
    return Tensor(out)


@contextlib.contextmanager
def name_scope(name, default_name=None, values=None):
    yield


import builtins as _b  # noqa: E402

builtins_range = _b.range
