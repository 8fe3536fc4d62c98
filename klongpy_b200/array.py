"""DeviceArray — the array type of the B200 backend (what `backend.np.ndarray` names).

A thin wrapper over a device-resident `torch.Tensor` used purely as the memory container (allocation, views,
H2D/D2H copies, autograd interop).  Every arithmetic / comparison / indexing method the KlongPy verbs invoke
directly on array objects (`x*1` in kg_truth types.py:339-340, `a - 1` adverbs.py:206,
`.trunc()`, `==`, `.all()` dyads.py:754-755, slicing dyads.py:261, 1029-1031, `.astype(int)` base.py:203 ...)
is routed to the CUDA library through `klongpy_b200.ops`; no torch compute op sits on that path.
"""
import numpy as np
import torch

from . import _lib

_NP2T = {
    np.dtype("float64"): torch.float64,
    np.dtype("int64"): torch.int64,
    np.dtype("float32"): torch.float32,
    np.dtype("int32"): torch.int32,
    np.dtype("bool"): torch.bool,
}
_T2NP = {v: k for k, v in _NP2T.items()}
_T2K = {torch.float64: _lib.F64, torch.int64: _lib.I64, torch.float32: _lib.F32, torch.int32: _lib.I32,
        torch.bool: _lib.B8}
_K2T = {v: k for k, v in _T2K.items()}


def torch_dtype_of(np_dtype):
    d = np.dtype(np_dtype)
    if d in _NP2T:
        return _NP2T[d]
    if d.kind in "iu":
        return torch.int64
    if d.kind == "f":
        return torch.float64
    if d.kind == "b":
        return torch.bool
    raise TypeError(f"dtype {d} cannot live on the device")


def kgb_dtype(t):
    try:
        return _T2K[t.dtype]
    except KeyError:
        raise TypeError(f"tensor dtype {t.dtype} is not supported by libkgb200") from None


# ops / lazy import this module, so they are bound on first use — ONCE: an `from . import ops` statement inside every
# operator method costs ~1 us of import machinery per call, a tenth of a small-vector launch
_ops = None
_lazy = None


def _bind_ops():
    global _ops
    from . import ops
    _ops = ops
    return ops


def _bind_lazy():
    global _lazy
    from . import lazy
    _lazy = lazy
    return lazy


class DeviceArray:
    """N-d array living in B200 HBM.  Immutable from the verbs' point of view (results are fresh arrays)."""

    __slots__ = ("_tensor", "_node", "__weakref__")
    __array_priority__ = 1000

    def __init__(self, tensor):
        if not isinstance(tensor, torch.Tensor):
            raise TypeError("DeviceArray wraps a torch.Tensor")
        self._tensor = tensor
        self._node = None

    @classmethod
    def _deferred(cls, node):
        """An elementwise expression that has not been launched yet (klongpy_b200.lazy, opt-in)."""
        obj = cls.__new__(cls)
        obj._tensor = None
        obj._node = node
        return obj

    @property
    def _t(self):
        """The backing tensor.  Touching it materialises a deferred expression (one fused kernel)."""
        if self._tensor is None:
            self._tensor = (_lazy or _bind_lazy()).materialize(self._node)
            self._node = None
        return self._tensor

    # ------------------------------------------------------------------ metadata (no device work)
    @property
    def shape(self):
        return self._node.shape if self._tensor is None else tuple(self._tensor.shape)

    @property
    def ndim(self):
        return len(self._node.shape) if self._tensor is None else self._tensor.dim()

    @property
    def size(self):
        if self._tensor is None:
            n = 1
            for d in self._node.shape:
                n *= d
            return n
        return self._tensor.numel()

    @property
    def dtype(self):
        return _T2NP[_K2T[self._node.kdt]] if self._tensor is None else _T2NP[self._tensor.dtype]

    @property
    def device(self):
        return self._node.device if self._tensor is None else self._tensor.device

    @property
    def tensor(self):
        """The underlying torch.Tensor (autograd / DLPack interop)."""
        return self._t

    @property
    def requires_grad(self):
        if self._tensor is None:
            return any(isinstance(a, DeviceArray) and a.requires_grad for a in self._node.args)
        return self._tensor.requires_grad

    @property
    def T(self):
        return DeviceArray(self._t.T)

    def __len__(self):
        shape = self.shape   # metadata only: never launches a deferred expression or gathers a sharded vector
        if len(shape) == 0:
            raise TypeError("len() of unsized object")
        return shape[0]

    def __repr__(self):
        return f"DeviceArray({self.to_numpy()!r}, device='{self._t.device}')"

    # ------------------------------------------------------------------ host transfer (explicit syncs)
    def to_numpy(self):
        if self._tensor is None and type(self._node).__name__ == "ShardNode":
            from . import sharded
            return sharded.to_numpy(self)   # block by block, no gather on a device
        return self._t.detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.to_numpy()
        return a.astype(dtype) if dtype is not None else a

    def tolist(self):
        return self.to_numpy().tolist()

    def item(self):
        if self._t.numel() != 1:
            raise ValueError("can only convert an array of size 1 to a Python scalar")
        return self._t.item()

    def __bool__(self):
        if self._t.numel() != 1:
            raise ValueError("The truth value of an array with more than one element is ambiguous. "
                             "Use a.any() or a.all()")
        return bool(self._t.item())

    def __int__(self):
        return int(self.item())

    def __float__(self):
        return float(self.item())

    def __index__(self):
        if self._t.dtype not in (torch.int64, torch.int32):
            raise TypeError("only integer scalar arrays can be converted to a scalar index")
        return int(self.item())

    def __iter__(self):
        if self._t.dim() == 0:
            raise TypeError("iteration over a 0-d array")
        if self._t.dim() == 1:
            # one D2H copy, then numpy scalars like ndarray iteration
            return iter(self.to_numpy())
        return (DeviceArray(self._t[i]) for i in range(self._t.shape[0]))

    # ------------------------------------------------------------------ views (metadata only)
    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        ops = _ops or _bind_ops()
        return DeviceArray(ops.contiguous(self)._t.reshape(*shape))

    def flatten(self):
        ops = _ops or _bind_ops()
        return DeviceArray(ops.contiguous(self)._t.reshape(-1))

    ravel = flatten

    def copy(self):
        ops = _ops or _bind_ops()
        return ops.copy(self)

    def astype(self, dtype, copy=True):
        ops = _ops or _bind_ops()
        return ops.astype(self, dtype)

    def __getitem__(self, key):
        ops = _ops or _bind_ops()
        return ops.getitem(self, key)

    # ------------------------------------------------------------------ arithmetic -> fused kernels
    def _bin(self, other, op, reflected=False):
        ops = _ops or _bind_ops()
        if isinstance(other, (list, tuple)):
            other = ops.asarray(other, device=self.device)
        if not ops.is_operand(other):
            return NotImplemented
        return ops.binary(op, other, self) if reflected else ops.binary(op, self, other)

    def __add__(self, o): return self._bin(o, "add")
    def __radd__(self, o): return self._bin(o, "add", True)
    def __sub__(self, o): return self._bin(o, "sub")
    def __rsub__(self, o): return self._bin(o, "sub", True)
    def __mul__(self, o): return self._bin(o, "mul")
    def __rmul__(self, o): return self._bin(o, "mul", True)
    def __truediv__(self, o): return self._bin(o, "div")
    def __rtruediv__(self, o): return self._bin(o, "div", True)
    def __pow__(self, o): return self._bin(o, "pow")
    def __rpow__(self, o): return self._bin(o, "pow", True)
    def __mod__(self, o): return self._bin(o, "mod")          # numpy's floored remainder (np.fmod is backend.np.fmod)
    def __rmod__(self, o): return self._bin(o, "mod", True)
    def __floordiv__(self, o): return self._bin(o, "floordiv")  # integer arrays stay exact integers
    def __rfloordiv__(self, o): return self._bin(o, "floordiv", True)
    def __lt__(self, o): return self._bin(o, "lt_b")
    def __gt__(self, o): return self._bin(o, "gt_b")
    def __le__(self, o): return self._bin(o, "le_b")
    def __ge__(self, o): return self._bin(o, "ge_b")
    def __eq__(self, o): return self._bin(o, "eq_b")
    def __ne__(self, o): return self._bin(o, "ne_b")

    def __neg__(self):
        ops = _ops or _bind_ops()
        return ops.unary("neg", self)

    def __pos__(self):
        return self

    def __abs__(self):
        ops = _ops or _bind_ops()
        return ops.unary("abs", self)

    def __invert__(self):
        ops = _ops or _bind_ops()
        return ops.unary("not_b", self)

    def trunc(self):
        ops = _ops or _bind_ops()
        return ops.unary("trunc", self)

    def all(self):
        ops = _ops or _bind_ops()
        return ops.all_(self)

    def any(self):
        ops = _ops or _bind_ops()
        return ops.any_(self)

    # numpy ufuncs applied to device arrays (np.float64(2) * a, np.equal(a, b), np.add.reduce(a) ...) run on the device
    # when the ufunc has a kernel opcode.  Anything else RAISES: there is no CPU fallback on this path (north_star); a
    # caller who wants host arithmetic says so with `.to_numpy()`.
    _UFUNC = {"add": "add", "subtract": "sub", "multiply": "mul", "divide": "div", "true_divide": "div", "power": "pow",
              "maximum": "max", "minimum": "min", "fmod": "fmod", "equal": "eq_b", "not_equal": "ne_b", "less": "lt_b",
              "greater": "gt_b", "less_equal": "le_b", "greater_equal": "ge_b", "negative": "neg", "absolute": "abs",
              "floor": "floor", "ceil": "ceil", "trunc": "trunc", "reciprocal": "recip", "logical_not": "not_b",
              "sqrt": "sqrt", "exp": "exp", "log": "log", "sin": "sin", "cos": "cos", "tan": "tan", "tanh": "tanh",
              "sign": "sign", "isnan": "isnan_b", "isinf": "isinf_b", "log2": "log2", "log10": "log10", "square": "square",
              "expm1": "expm1", "log1p": "log1p", "sinh": "sinh", "cosh": "cosh", "arcsin": "asin", "arccos": "acos",
              "arctan": "atan", "floor_divide": "floordiv", "remainder": "mod", "mod": "mod"}

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        ops = _ops or _bind_ops()
        name = self._UFUNC.get(ufunc.__name__)
        if name is not None and not kwargs:
            if method == "__call__" and all(ops.is_operand(x) for x in inputs):
                if len(inputs) == 2:
                    return ops.binary(name, inputs[0], inputs[1])
                if len(inputs) == 1:
                    return ops.unary(name, inputs[0])
            if method in ("reduce", "accumulate") and name in ("add", "mul", "max", "min") and len(inputs) == 1:
                return ops.reduce(name, inputs[0]) if method == "reduce" else ops.scan(name, inputs[0])
        raise TypeError(f"numpy.{ufunc.__name__}.{method} has no device kernel for these operands"
                        f"{' / keyword arguments ' + repr(sorted(kwargs)) if kwargs else ''}; "
                        "call .to_numpy() explicitly to compute on the host")

    def __hash__(self):
        # 0-d device values appear as dictionary keys (`:{[k v]}` builds `a[b[0]] = b[1]`, dyads dictionary code):
        # hash like the host scalar they stand for; arrays stay unhashable like numpy's
        if self.ndim == 0:
            return hash(self.item())
        raise TypeError("unhashable type: 'DeviceArray'")

    def _fold(self, name, axis):
        # ndarray.sum/max/min: axis=None folds EVERY element, axis=0 the leading axis; other axes are not on the path
        ops = _ops or _bind_ops()
        if axis is None:
            return ops.reduce(name, self.flatten() if self.ndim > 1 else self)
        if axis in (0, -self.ndim) and self.ndim >= 1:
            return ops.reduce(name, self)
        raise NotImplementedError(f"DeviceArray.{name}: axis={axis!r} (only None and 0 have kernels)")

    def sum(self, axis=None, out=None, **_kw):
        return self._fold("add", axis)

    def max(self, axis=None, out=None, **_kw):
        return self._fold("max", axis)

    def min(self, axis=None, out=None, **_kw):
        return self._fold("min", axis)

    def argsort(self, kind=None):
        ops = _ops or _bind_ops()
        return ops.argsort(self)
