"""`provider.np` — the duck-typed NumPy-like module the KlongPy verbs dereference.

An explicit allow-list: exactly the attribute set `dyads.py` / `monads.py` / `adverbs.py` / `sys_fn.py` touch
through `backend.np.*`, each routed to libkgb200 kernels via `klongpy_b200.ops`.  Unlike the torch shim
(`klongpy/backends/torch_backend.py:144-154`) unknown names are NOT forwarded to torch or numpy: no library
compute op can sneak onto the device path.  Host object arrays (strings, dicts, jagged lists — host data in
KlongPy by construction) keep NumPy semantics.
"""
import numpy as np
import torch

from . import ops
from .array import DeviceArray


def _is_obj(x):
    return isinstance(x, np.ndarray) and x.dtype == object


def _is_str_arr(x):
    return isinstance(x, np.ndarray) and x.dtype.kind in "US"


class _Ufunc:
    """callable + .reduce + .accumulate, like the numpy ufuncs the adverbs use (adverbs.py:232-239, 325-332)."""

    def __init__(self, shim, op, numpy_ufunc):
        self._shim = shim
        self._op = op
        self._nu = numpy_ufunc
        self.__name__ = numpy_ufunc.__name__

    def __call__(self, a, b):
        if _is_obj(a) or _is_obj(b) or isinstance(a, str) or isinstance(b, str):
            return self._nu(self._shim._host(a), self._shim._host(b))
        return ops.binary(self._op, a, b)

    def reduce(self, a, axis=None):
        if _is_obj(a):
            return self._nu.reduce(a, axis=0 if axis is None else axis)
        a = self._shim.asarray(a)
        op = self._op
        if op in ("add", "mul", "max", "min"):
            return ops.reduce(op, a)
        if a._t.dim() != 1:
            raise NotImplementedError(f"{self.__name__}.reduce on N-d device arrays")
        if a._t.numel() == 0:
            raise ValueError("zero-size array to reduction operation which has no identity")
        first = DeviceArray(a._t[0])
        if a._t.numel() == 1:
            return first
        rest = DeviceArray(a._t[1:])
        # -/a = a0 - (a1 + ... + an),  %/a = a0 / (a1 * ... * an)
        if op == "sub":
            return ops.binary("sub", first, ops.reduce("add", rest))
        if op == "div":
            return ops.binary("div", first, ops.reduce("mul", rest))
        raise NotImplementedError(f"{self.__name__}.reduce")

    def accumulate(self, a, axis=0):
        if _is_obj(a):
            return self._nu.accumulate(a, axis=axis)
        a = self._shim.asarray(a)
        op = self._op
        if op in ("add", "mul", "max", "min"):
            return ops.scan(op, a)
        if a._t.dim() != 1:
            raise NotImplementedError(f"{self.__name__}.accumulate on N-d device arrays")
        if a._t.numel() <= 1:
            return ops.copy(a) if op == "sub" else ops.astype(a, np.float64)
        # x - y == x + (-y) exactly, so the running difference is the running sum of [a0, -a1, -a2, ...]
        if op == "sub":
            b = ops.unary("neg", a)
            b._t[0:1].copy_(a._t[0:1])  # 8-byte D2D memcpy
            return ops.scan("add", b)
        if op == "div":
            b = ops.unary("recip", a)
            b._t[0:1].copy_(ops.astype(DeviceArray(a._t[0:1]), np.float64)._t)
            return ops.scan("mul", b)
        raise NotImplementedError(f"{self.__name__}.accumulate")


def _unary(name, op, numpy_fn):
    def f(self, x):
        if _is_obj(x):
            return numpy_fn(x)
        return ops.unary(op, x if not isinstance(x, (list, tuple)) else self.asarray(x))
    f.__name__ = name
    return f


def _binary(name, op, numpy_fn):
    def f(self, a, b):
        if _is_obj(a) or _is_obj(b):
            return numpy_fn(self._host(a), self._host(b))
        return ops.binary(op, a, b)
    f.__name__ = name
    return f


class B200Random:
    """`backend.np.random` — device-resident draws (torch precedent: torch_backend.py:24-78).  Generation is
    input synthesis, not a Klong primitive, so torch's device RNG fills the buffers."""

    def __init__(self, shim):
        self._shim = shim
        self._gen = None

    def _g(self):
        if self._gen is None:
            self._gen = torch.Generator(device=self._shim.device)
        return self._gen

    def seed(self, seed):
        self._g().manual_seed(int(seed))

    def random(self, size=None):
        shape = () if size is None else (size,) if isinstance(size, (int, np.integer)) else tuple(size)
        return DeviceArray(torch.rand(shape, dtype=torch.float64, device=self._shim.device, generator=self._g()))

    def rand(self, *shape):
        return self.random(shape)

    def randn(self, *shape):
        return DeviceArray(torch.randn(shape, dtype=torch.float64, device=self._shim.device, generator=self._g()))

    def randint(self, low, high=None, size=None):
        if high is None:
            low, high = 0, low
        shape = () if size is None else (size,) if isinstance(size, (int, np.integer)) else tuple(size)
        return DeviceArray(torch.randint(int(low), int(high), shape, dtype=torch.int64, device=self._shim.device,
                                         generator=self._g()))


class B200Numpy:
    """Module-like object returned by `B200BackendProvider.np`."""

    ndarray = DeviceArray          # compiler.py:56 isinstance check
    integer = np.integer
    floating = np.floating
    inf = np.inf
    nan = np.nan
    pi = np.pi
    e = np.e
    float64 = np.float64
    int64 = np.int64
    float32 = np.float32
    int32 = np.int32
    bool_ = np.bool_

    def __init__(self, device):
        self.device = torch.device(device)
        self.random = B200Random(self)
        self.add = _Ufunc(self, "add", np.add)
        self.subtract = _Ufunc(self, "sub", np.subtract)
        self.multiply = _Ufunc(self, "mul", np.multiply)
        self.divide = _Ufunc(self, "div", np.divide)

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _host(x):
        return x.to_numpy() if isinstance(x, DeviceArray) else x

    def seterr(self, **kwargs):
        return {}

    def isarray(self, x):
        return isinstance(x, (np.ndarray, DeviceArray))

    def asarray(self, a, dtype=None):
        if dtype is not None and (dtype is object or np.dtype(dtype) == object):
            return np.asarray(self._host(a), dtype=object)
        if isinstance(a, DeviceArray):
            return a if dtype is None else ops.astype(a, dtype)
        if isinstance(a, np.ndarray) and (a.dtype == object or a.dtype.kind in "US"):
            return a
        if isinstance(a, (list, tuple)):
            if len(a) and all(isinstance(x, DeviceArray) for x in a):
                return self.stack(a)
            arr = np.asarray([self._host(x) for x in a], dtype=dtype)
            if arr.dtype == object or arr.dtype.kind in "US":
                return arr
            return ops.asarray(arr, device=self.device)
        return ops.asarray(a, device=self.device, dtype=dtype)

    def array(self, a, dtype=None):
        r = self.asarray(a, dtype=dtype)
        if r is a:
            return ops.copy(r) if isinstance(r, DeviceArray) else np.array(r)
        return r

    def copy(self, a):
        return ops.copy(a) if isinstance(a, DeviceArray) else np.copy(a)

    # ------------------------------------------------------------------ constructors
    def _filled(self, shape, value, dtype=None):
        return ops.full(shape, value, dtype, self.device)

    def full(self, shape, fill_value, dtype=None):
        if isinstance(fill_value, (str, np.ndarray)) or _is_obj(fill_value):
            return np.full(shape, fill_value, dtype=dtype)
        r = self._filled(shape, fill_value)
        return r if dtype is None else ops.astype(r, dtype)

    def zeros(self, shape, dtype=None):
        return self._filled(shape, 0.0 if dtype in (None, float, np.float64) else 0)

    def ones(self, shape, dtype=None):
        return self._filled(shape, 1.0 if dtype in (None, float, np.float64) else 1)

    def arange(self, *args, **kwargs):
        a = [int(x) for x in args]
        if len(a) == 1:
            start, stop, step = 0, a[0], 1
        elif len(a) == 2:
            start, stop, step = a[0], a[1], 1
        else:
            start, stop, step = a
        n = max(0, -(-(stop - start) // step)) if step > 0 else max(0, -(-(start - stop) // -step))
        return ops.iota(n, self.device, start, step)

    # ------------------------------------------------------------------ shape ops
    def concatenate(self, arrays, axis=0):
        if isinstance(arrays, DeviceArray):
            if arrays._t.dim() < 2:
                return arrays
            return arrays.reshape((-1,) + arrays.shape[2:])
        arrays = list(arrays)
        if any(_is_obj(a) or _is_str_arr(a) for a in arrays):
            return np.concatenate([self._host(a) for a in arrays], axis=axis)
        arrs = [self.asarray(a) for a in arrays]
        arrs = [a if a._t.dim() > 0 else a.reshape(1) for a in arrs]
        if axis != 0:
            raise NotImplementedError("concatenate along axis != 0 on device arrays")
        kinds = {a.dtype for a in arrs}
        if len(kinds) > 1:
            d = np.result_type(*kinds)
            arrs = [ops.astype(a, d) if a.dtype != d else a for a in arrs]
        total = sum(a.shape[0] for a in arrs)
        out = torch.empty((total,) + tuple(arrs[0].shape[1:]), dtype=arrs[0]._t.dtype, device=self.device)
        o = 0
        for a in arrs:  # D2D memcpys into the result
            k = a.shape[0]
            if k:
                out[o:o + k].copy_(a._t)
            o += k
        return DeviceArray(out)

    def stack(self, arrays, axis=0):
        arrs = [self.asarray(a) for a in arrays]
        if axis != 0:
            raise NotImplementedError("stack along axis != 0 on device arrays")
        kinds = {a.dtype for a in arrs}
        if len(kinds) > 1:
            d = np.result_type(*kinds)
            arrs = [ops.astype(a, d) if a.dtype != d else a for a in arrs]
        out = torch.empty((len(arrs),) + tuple(arrs[0].shape), dtype=arrs[0]._t.dtype, device=self.device)
        for i, a in enumerate(arrs):
            out[i].copy_(a._t)
        return DeviceArray(out)

    def tile(self, a, reps):
        if not isinstance(a, DeviceArray):
            return np.tile(a, self._host(reps))
        if isinstance(reps, DeviceArray):
            reps = reps.tolist()
        if isinstance(reps, (list, tuple, np.ndarray)):
            reps = [int(r) for r in reps]
            if len(reps) != 1 and not all(r == 1 for r in reps[1:]):
                raise NotImplementedError("tile with multi-axis repetitions on device arrays")
            reps = reps[0]
        reps = int(reps)
        n = a.shape[0] if a.ndim else 1
        if reps <= 0 or n == 0:
            return DeviceArray(a._t[:0])
        idx = ops.binary("fmod", ops.iota(n * reps, self.device), n)
        return ops.gather(a if a.ndim else a.reshape(1), idx)

    def resize(self, a, new_shape):
        if not isinstance(a, DeviceArray):
            return np.resize(a, new_shape)
        new_shape = (int(new_shape),) if isinstance(new_shape, (int, np.integer)) else \
            tuple(int(s) for s in (new_shape.tolist() if isinstance(new_shape, DeviceArray) else new_shape))
        total = 1
        for s in new_shape:
            total *= s
        flat = a.flatten()
        n = flat.shape[0]
        if total == 0 or n == 0:
            return DeviceArray(torch.zeros(new_shape, dtype=a._t.dtype, device=self.device))
        idx = ops.binary("fmod", ops.iota(total, self.device), n)
        return ops.gather(flat, idx).reshape(new_shape)

    def flip(self, a, axis=None, dims=None):
        if not isinstance(a, DeviceArray):
            return np.flip(a, axis=axis)
        if dims is not None:
            raise TypeError("flip() takes axis=, not dims=")  # monads.py:335-339 then retries numpy style
        if axis not in (None, 0) and not (axis is None and a.ndim == 1):
            raise NotImplementedError("flip along axis != 0 on device arrays")
        return ops.reverse(a)

    def where(self, cond, x=None, y=None):
        if x is None and y is None:
            if isinstance(cond, DeviceArray):
                return (ops.nonzero(cond),)
            return np.where(cond)
        if isinstance(cond, DeviceArray) or isinstance(x, DeviceArray) or isinstance(y, DeviceArray):
            return ops.where3(cond, x, y)
        return np.where(cond, x, y)

    def repeat(self, a, repeats):
        if isinstance(repeats, DeviceArray) or isinstance(a, DeviceArray):
            a = self.asarray(a)
            reps = self.asarray(repeats)
            if reps.ndim == 0:
                reps = self._filled(a.shape[0], reps.item(), np.int64)
            return ops.gather(a, ops.expand(reps))
        return np.repeat(a, repeats)

    def unique(self, a, return_index=False, return_inverse=False):
        raise NotImplementedError("np.unique on device arrays: use the group / range verbs (ops.group, ops.unique_first)")

    def argsort(self, a, kind=None):
        return ops.argsort(self.asarray(a))

    def take(self, a, indices, axis=None):
        return ops.gather(self.asarray(a), self.asarray(indices))

    def transpose(self, a, axes=None):
        if isinstance(a, DeviceArray):
            return ops.contiguous(DeviceArray(a._t.T if axes is None else a._t.permute(*axes)))
        return np.transpose(a, axes)

    # ------------------------------------------------------------------ folds
    def _fold(self, name, a, axis):
        # numpy semantics: axis=None folds every element, axis=0 the leading axis (what `ufunc.reduce` does by default)
        a = self.asarray(a)
        if axis is None:
            return ops.reduce(name, a.flatten() if a.ndim > 1 else a)
        if axis in (0, -a.ndim) and a.ndim >= 1:
            return ops.reduce(name, a)
        raise NotImplementedError(f"np.{name}: axis={axis!r} on a device array (only None and 0 have kernels)")

    def sum(self, a, axis=None):
        return self._fold("add", a, axis) if not _is_obj(a) else np.sum(a, axis=axis)

    def prod(self, a, axis=None):
        if _is_obj(a):
            return np.prod(a, axis=axis)
        if isinstance(a, (list, tuple)) or (isinstance(a, np.ndarray)):
            return np.prod(a)  # shapes are host data
        return ops.reduce("mul", self.asarray(a))

    def max(self, a, axis=None):
        return self._fold("max", a, axis) if not _is_obj(a) else np.max(a, axis=axis)

    def min(self, a, axis=None):
        return self._fold("min", a, axis) if not _is_obj(a) else np.min(a, axis=axis)

    def cumsum(self, a, axis=0):
        return ops.scan("add", self.asarray(a))

    def cumprod(self, a, axis=0):
        return ops.scan("mul", self.asarray(a))

    def array_equal(self, a, b):
        if isinstance(a, DeviceArray) and isinstance(b, DeviceArray):
            return ops.array_equal(a, b)
        return bool(np.array_equal(self._host(a), self._host(b)))

    def all(self, a):
        return ops.all_(a) if isinstance(a, DeviceArray) else np.all(a)

    def any(self, a):
        return ops.any_(a) if isinstance(a, DeviceArray) else np.any(a)

    # ------------------------------------------------------------------ elementwise (dyads.py / monads.py / .bkf)
    maximum = _binary("maximum", "max", np.maximum)
    minimum = _binary("minimum", "min", np.minimum)
    fmod = _binary("fmod", "fmod", np.fmod)
    less = _binary("less", "lt_b", np.less)
    greater = _binary("greater", "gt_b", np.greater)
    less_equal = _binary("less_equal", "le_b", np.less_equal)
    greater_equal = _binary("greater_equal", "ge_b", np.greater_equal)
    equal = _binary("equal", "eq_b", np.equal)
    not_equal = _binary("not_equal", "ne_b", np.not_equal)
    power = _binary("power", "pow", np.power)

    def isclose(self, a, b, rtol=1e-05, atol=1e-08):
        if (rtol, atol) != (1e-05, 1e-08) or _is_obj(a) or _is_obj(b):
            return np.isclose(self._host(a), self._host(b), rtol=rtol, atol=atol)
        return ops.binary("isclose", a, b)

    negative = _unary("negative", "neg", np.negative)
    abs = _unary("abs", "abs", np.abs)
    absolute = abs
    trunc = _unary("trunc", "trunc", np.trunc)
    floor = _unary("floor", "floor", np.floor)
    ceil = _unary("ceil", "ceil", np.ceil)
    sign = _unary("sign", "sign", np.sign)
    sqrt = _unary("sqrt", "sqrt", np.sqrt)
    exp = _unary("exp", "exp", np.exp)
    log = _unary("log", "log", np.log)
    log2 = _unary("log2", "log2", np.log2)
    log10 = _unary("log10", "log10", np.log10)
    log1p = _unary("log1p", "log1p", np.log1p)
    expm1 = _unary("expm1", "expm1", np.expm1)
    sin = _unary("sin", "sin", np.sin)
    cos = _unary("cos", "cos", np.cos)
    tan = _unary("tan", "tan", np.tan)
    tanh = _unary("tanh", "tanh", np.tanh)
    sinh = _unary("sinh", "sinh", np.sinh)
    cosh = _unary("cosh", "cosh", np.cosh)
    arcsin = _unary("arcsin", "asin", np.arcsin)
    arccos = _unary("arccos", "acos", np.arccos)
    arctan = _unary("arctan", "atan", np.arctan)
    square = _unary("square", "square", np.square)
    reciprocal = _unary("reciprocal", "recip", np.reciprocal)
    logical_not = _unary("logical_not", "not_b", np.logical_not)
    isnan = _unary("isnan", "isnan_b", np.isnan)
    isinf = _unary("isinf", "isinf_b", np.isinf)
