"""Host-side bridge from Python operands to libkgb200 launches.

Everything here is argument marshalling: classify operands (device array / 0-d device scalar / host scalar),
build or look up the postfix program, allocate the result with torch (memory container only) and launch on
torch's current stream.  Host scalars combined with host scalars stay on the host (they are interpreter
control flow) — no device value ever falls back to NumPy.
"""
import ctypes
import struct
import threading

import numpy as np
import torch

from . import _lib
from .array import DeviceArray, _K2T, _T2K as _T2K_FAST, kgb_dtype, torch_dtype_of

# ---------------------------------------------------------------------------------------- opcodes
OP_ARG, OP_LIT_I64, OP_LIT_F64, OP_SCANVAL = 1, 2, 3, 4
OPC = {
    "add": 16, "sub": 17, "mul": 18, "div": 19, "pow": 20, "eq": 21, "gt": 22, "lt": 23,
    "max": 24, "min": 25, "fmod": 26, "eq_b": 27, "gt_b": 28, "lt_b": 29, "ne_b": 30, "ge_b": 31, "le_b": 32,
    "isclose": 33, "floordiv": 34, "mod": 35,
    "neg": 48, "abs": 49, "floor": 50, "floor_i64": 51, "trunc": 52, "recip": 53, "not_b": 54,
    "to_f64": 55, "to_i64": 56, "to_f32": 57, "to_b8": 58, "to_i32": 59,
    "sqrt": 64, "exp": 65, "log": 66, "sin": 67, "cos": 68, "tan": 69, "tanh": 70, "ceil": 71, "sign": 72,
    "isnan_b": 73, "isinf_b": 74, "log2": 75, "log10": 76, "sinh": 77, "cosh": 78, "asin": 79, "acos": 80,
    "atan": 81, "expm1": 82, "log1p": 83, "square": 84,
    "where": 96,
}
RED = {"add": _lib.RED_ADD, "mul": _lib.RED_MUL, "max": _lib.RED_MAX, "min": _lib.RED_MIN,
       "+": _lib.RED_ADD, "*": _lib.RED_MUL, "|": _lib.RED_MAX, "&": _lib.RED_MIN}

# numpy equivalents for host-scalar (control-flow) arithmetic
_HOST = {
    "add": np.add, "sub": np.subtract, "mul": np.multiply, "div": np.divide, "pow": np.power,
    "max": np.maximum, "min": np.minimum, "fmod": np.fmod, "eq_b": np.equal, "gt_b": np.greater,
    "lt_b": np.less, "ne_b": np.not_equal, "ge_b": np.greater_equal, "le_b": np.less_equal,
    "isclose": np.isclose, "floordiv": np.floor_divide, "mod": np.remainder,
    "neg": np.negative, "abs": np.abs, "floor": np.floor, "trunc": np.trunc, "ceil": np.ceil,
    "recip": lambda x: np.reciprocal(np.asarray(x, dtype=float))[()], "not_b": np.logical_not,
    "sqrt": np.sqrt, "exp": np.exp, "log": np.log, "sin": np.sin, "cos": np.cos, "tan": np.tan,
    "tanh": np.tanh, "sign": np.sign, "isnan_b": np.isnan, "isinf_b": np.isinf, "log2": np.log2,
    "log10": np.log10, "sinh": np.sinh, "cosh": np.cosh, "asin": np.arcsin, "acos": np.arccos,
    "atan": np.arctan, "expm1": np.expm1, "log1p": np.log1p, "square": np.square,
    "floor_i64": lambda x: int(np.floor(x)),
}


def lit(v):
    """Inline literal words for a postfix program."""
    if isinstance(v, (bool, np.bool_)):
        v = int(v)
    if isinstance(v, (int, np.integer)):
        lo, hi = struct.unpack("<ii", struct.pack("<q", int(v)))
        return [OP_LIT_I64, lo, hi]
    lo, hi = struct.unpack("<ii", struct.pack("<d", float(v)))
    return [OP_LIT_F64, lo, hi]


# ---------------------------------------------------------------------------------------- device state
_default_device = None


def set_default_device(device):
    global _default_device
    _default_device = torch.device(device)


def default_device():
    global _default_device
    if _default_device is None:
        if not torch.cuda.is_available():
            raise _lib.KgbError("klongpy_b200 needs an NVIDIA B200 (sm_100) GPU; there is no CPU fallback")
        _default_device = torch.device("cuda", torch.cuda.current_device())
    return _default_device


_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)  # torch's own fast accessor of the current stream


def _lib_for(device):
    """Library bound to `device` on this thread (and the raw stream to launch on)."""
    if device.type != "cuda":
        lib = _lib.load_for_host_device(device)  # only a test double registers itself here
        return lib, None
    idx = device.index if device.index is not None else torch.cuda.current_device()
    lib = _lib.bind_device(idx)
    if _RAW_STREAM is not None:
        return lib, ctypes.c_void_p(_RAW_STREAM(idx))
    return lib, ctypes.c_void_p(torch.cuda.current_stream(idx).cuda_stream)


# ---------------------------------------------------------------------------------------- operands
_SCALARS = (int, float, bool, np.integer, np.floating, np.bool_)


def is_host_scalar(x):
    return isinstance(x, _SCALARS) or (isinstance(x, np.ndarray) and x.ndim == 0 and x.dtype.kind in "ifub")


def is_operand(x):
    return isinstance(x, DeviceArray) or is_host_scalar(x) or \
        (isinstance(x, np.ndarray) and x.dtype.kind in "ifub")


def asarray(a, device=None, dtype=None):
    """Host data -> DeviceArray (one H2D copy).  DeviceArrays pass through unchanged
    (same-object contract, tests/test_torch_backend.py:112-118)."""
    if isinstance(a, DeviceArray):
        return a if dtype is None else astype(a, dtype)
    device = torch.device(device) if device is not None else default_device()
    if isinstance(a, torch.Tensor):
        t = a if a.device == device else a.to(device)
        return DeviceArray(t)
    arr = np.asarray(a, dtype=dtype)
    if arr.dtype == object or arr.dtype.kind not in "ifub":
        raise TypeError(f"cannot place dtype {arr.dtype} on the device")
    td = torch_dtype_of(arr.dtype)
    if arr.ndim > 0:  # (np.ascontiguousarray would turn a 0-d array into shape (1,))
        arr = np.ascontiguousarray(arr)
    t = torch.from_numpy(arr) if arr.dtype in (np.float64, np.int64, np.float32, np.int32, np.bool_) \
        else torch.from_numpy(arr.astype(np.int64 if arr.dtype.kind in "iu" else np.float64))
    return DeviceArray(t.to(device=device, dtype=td))


def contiguous(x):
    return x if x._t.is_contiguous() else DeviceArray(x._t.contiguous())


def empty(shape, kdt, device):
    return DeviceArray(torch.empty(shape, dtype=_K2T[kdt], device=device))


# ---------------------------------------------------------------------------------------- program runner
_handles = {}


def _classify(x, device):
    """-> (kind, kgb dtype, keepalive object, raw pointer)"""
    if isinstance(x, DeviceArray):
        t = x._t
        if t.device != device:
            raise ValueError(f"operand on {t.device}, expected {device}")
        if not t.is_contiguous():
            t = t.contiguous()
            x = DeviceArray(t)
        return (_lib.ARG_DEVSCALAR if t.dim() == 0 else _lib.ARG_ARRAY), kgb_dtype(t), x, t.data_ptr()
    if isinstance(x, (bool, np.bool_)):
        x = int(x)
    if isinstance(x, np.ndarray) and x.ndim == 0:
        x = x[()]
    if isinstance(x, (int, np.integer)):
        c = ctypes.c_int64(int(x))
        return _lib.ARG_SCALAR, _lib.I64, c, ctypes.addressof(c)
    if isinstance(x, (float, np.floating)):
        c = ctypes.c_double(float(x))
        return _lib.ARG_SCALAR, _lib.F64, c, ctypes.addressof(c)
    raise TypeError(f"unsupported operand type {type(x).__name__}")


def _classify_meta(x):
    """(kind, kgb dtype) of an operand from metadata only (never materialises a deferred array)."""
    if isinstance(x, DeviceArray):
        t = x._tensor
        if t is None:
            nd = x._node
            return (_lib.ARG_DEVSCALAR if len(nd.shape) == 0 else _lib.ARG_ARRAY), nd.kdt
        return (_lib.ARG_DEVSCALAR if t.dim() == 0 else _lib.ARG_ARRAY), kgb_dtype(t)
    if isinstance(x, (bool, np.bool_, int, np.integer)):
        return _lib.ARG_SCALAR, _lib.I64
    if isinstance(x, np.ndarray) and x.ndim == 0:
        return _lib.ARG_SCALAR, (_lib.I64 if x.dtype.kind in "iub" else _lib.F64)
    if isinstance(x, (float, np.floating)):
        return _lib.ARG_SCALAR, _lib.F64
    raise TypeError(f"unsupported operand type {type(x).__name__}")


from . import sharded as _sharded  # noqa: E402 - imports only array.py; ops is imported lazily inside its functions

_autograd_mod = None


def _bind_autograd():
    global _autograd_mod
    from . import autograd
    _autograd_mod = autograd
    return autograd


_ARGV = [ctypes.c_void_p * max(k, 1) for k in range(17)]  # argument-pointer array types, by argument count
_lazy = None  # klongpy_b200.lazy, imported on first use (it imports this module)


def _compile_handle(lib, device, code_t, code2_t, kinds, dts, mode, red, red2):
    """(handle, kgb output dtype, torch output dtype) of one specialised program: NVRTC on the first use, a dictionary
    hit afterwards."""
    key = (device.index, code_t, code2_t, kinds, dts, mode, red, red2)
    ent = _handles.get(key)
    if ent is None:
        n_args = len(kinds)
        c_code = (ctypes.c_int32 * len(code_t))(*code_t)
        c_code2 = (ctypes.c_int32 * len(code2_t))(*code2_t) if code2_t else None
        c_k = (ctypes.c_int32 * max(n_args, 1))(*kinds)
        c_d = (ctypes.c_int32 * max(n_args, 1))(*dts)
        h = ctypes.c_void_p()
        od = ctypes.c_int32()
        _lib.check(lib.kgb_expr_compile(c_code, len(code_t), c_code2, len(code2_t), c_k, c_d, n_args, mode, red,
                                        red2, ctypes.byref(h), ctypes.byref(od)), "kgb_expr_compile")
        ent = (h, od.value, _K2T[od.value])
        _handles[key] = ent
    return ent


def run(code, args, mode=_lib.MODE_MAP, red=0, code2=None, red2=0, out_shape=None, device=None, _no_defer=False):
    """Compile (cached) and launch one expression program.  Returns a DeviceArray."""
    global _lazy
    if _lazy is None:
        from . import lazy as _lazy_mod
        _lazy = _lazy_mod
    if _sharded.ACTIVE and any(_sharded.is_sharded(a) for a in args):
        # a vector block-sharded over several devices of this process (klongpy_b200.sharded): same program per shard
        r = _sharded.run(code, args, mode, red, code2, red2)
        if r is not NotImplemented:
            return r
    if _lazy.ENABLED and not _no_defer:
        # deferred operands are inlined into this program (elementwise chains, reduce / scan prologues fuse);
        # a MAP result is itself deferred until something needs the data.  Small vectors launch at once (lazy.MIN_SIZE):
        # a cheap look at the operands keeps the deferral machinery off their path entirely.
        has_lazy, n_el = False, 0
        for a in args:
            if isinstance(a, DeviceArray):
                t = a._tensor
                if t is None:
                    has_lazy = True
                    break
                if t.dim() > 0:
                    n_el = t.numel()
        if has_lazy or n_el == 0 or n_el >= _lazy.MIN_SIZE:
            code, args = _lazy.inline(code, args)
            if mode == _lib.MODE_MAP and out_shape is None and not code2:
                r = _lazy.defer(code, args)
                if r is not None:
                    return r
    # ---- one pass over the operands: shape agreement, autograd, and the launch description (kind, dtype, pointer)
    shape = None
    mixed = False
    grad = False
    dev_param = device
    kinds, dts, ptrs, keep = [], [], [], []
    for a in args:
        if isinstance(a, DeviceArray):
            t = a._t
            device = t.device
            if t.requires_grad:
                grad = True
            nd = t.dim()
            if nd > 0:
                if shape is None:
                    shape = tuple(t.shape)
                elif tuple(t.shape) != shape:
                    mixed = True
                if not t.is_contiguous():
                    t = t.contiguous()
                    keep.append(t)
                kinds.append(_lib.ARG_ARRAY)
            else:
                kinds.append(_lib.ARG_DEVSCALAR)
            dts.append(kgb_dtype(t))
            ptrs.append(t.data_ptr())
        else:
            if isinstance(a, (bool, np.bool_)):
                a = int(a)
            elif isinstance(a, np.ndarray) and a.ndim == 0:
                a = a[()]
            if isinstance(a, (int, np.integer)):
                c = ctypes.c_int64(int(a))
                dts.append(_lib.I64)
            elif isinstance(a, (float, np.floating)):
                c = ctypes.c_double(float(a))
                dts.append(_lib.F64)
            else:
                raise TypeError(f"unsupported operand type {type(a).__name__}")
            keep.append(c)
            kinds.append(_lib.ARG_SCALAR)
            ptrs.append(ctypes.addressof(c))
    if grad and torch.is_grad_enabled():
        autograd = _autograd_mod or _bind_autograd()  # operands on the autograd tape: same launch, recorded as one node
        return autograd.run_differentiable(code, args, mode, red, code2, red2, out_shape, dev_param)
    if mixed:
        # NumPy broadcasting between arrays of different shapes (`[[1]]&[2]`, tests/kgtests/language/test_broadcasting.kg):
        # rare and small in Klong programs, so the operands are expanded to the common shape first (one copy each)
        shapes = [tuple(a._t.shape) for a in args if isinstance(a, DeviceArray) and a._t.dim() > 0]
        try:
            shape = tuple(np.broadcast_shapes(*shapes))
        except ValueError:
            raise ValueError("operands could not be broadcast together with shapes " + " ".join(map(str, shapes))) from None
        args = [DeviceArray(a._t.expand(shape).contiguous()) if isinstance(a, DeviceArray) and a._t.dim() > 0 and
                tuple(a._t.shape) != shape else a for a in args]
        return run(code, args, mode=mode, red=red, code2=code2, red2=red2, out_shape=out_shape, device=dev_param, _no_defer=True)
    if device is None:
        raise TypeError("run() needs at least one device operand")
    for a in args:
        if isinstance(a, DeviceArray) and a._t.device != device:
            raise ValueError(f"operand on {a._t.device}, expected {device}")
    if shape is None and out_shape is not None:
        shape = tuple(out_shape)
    lib, stream = _lib_for(device)
    kinds = tuple(kinds)
    dts = tuple(dts)
    code_t = tuple(code)
    code2_t = tuple(code2) if code2 else ()
    h, od, out_td = _compile_handle(lib, device, code_t, code2_t, kinds, dts, mode, red, red2)
    n = 1
    if shape is not None:
        for s in shape:
            n *= s
    if mode in (_lib.MODE_REDUCE, _lib.MODE_SCANRED):
        oshape = ()
        if n == 0:
            raise ValueError("zero-size array to reduction operation which has no identity")
    else:
        oshape = shape if shape is not None else ()
    if out_shape is not None:
        oshape = out_shape
    out = torch.empty(oshape, dtype=out_td, device=device)
    ws_ptr = None
    ws_bytes = 0
    if mode != _lib.MODE_MAP:
        wsb = ctypes.c_size_t()
        _lib.check(lib.kgb_expr_workspace_bytes(h, n, ctypes.byref(wsb)), "kgb_expr_workspace_bytes")
        ws_bytes = wsb.value
        ws_ptr = _scratch(device, stream, ws_bytes)
    if n == 0:
        return DeviceArray(out)  # empty MAP / SCAN: nothing to launch
    na = len(ptrs)
    argv = _ARGV[na](*ptrs) if na < len(_ARGV) else (ctypes.c_void_p * na)(*ptrs)
    rc = lib.kgb_expr_launch_ex(h, argv, n, out.data_ptr(), ws_ptr, ws_bytes, stream, _lib.LAUNCH_WS_CLEAN)
    if rc:
        _lib.check(rc, "kgb_expr_launch")
    return DeviceArray(out)


class FastStage:
    """Pre-resolved launch state of ONE compiled stage for the common call shape of small-vector programs
    (`examples/bench_compiler.kg`): every operand a materialised, contiguous 1-d device array of one length on one device,
    a 0-d device value (an earlier stage's result) or a Python int / float, nothing on the autograd tape, nothing to
    defer.  Then a call is: one signature tuple, one dictionary hit, one output allocation, one C-ABI launch.  Anything
    else returns NotImplemented and the caller goes through `run` / `reduce2` (same kernels, same handle cache: this is
    a shortcut through the host bookkeeping, not another path).  With `code2` the stage is the two-output fold
    (`(+/p*s)%+/s`: both sums in one pass) and the call returns two 0-d values."""
    __slots__ = ("code", "code2", "code3", "mode", "red", "red2", "cache", "ws_bytes")

    def __init__(self, code, mode, red, code2=None, red2=0, code3=None):
        self.code, self.code2 = tuple(code), tuple(code2) if code2 else ()
        self.code3 = tuple(code3) if code3 else ()   # epilogue over the two finished folds: ONE result, one launch
        self.mode, self.red, self.red2 = mode, red, red2
        self.cache = {}      # (device index, operand signature) -> (handle, output torch dtype[, second output's])
        self.ws_bytes = {}   # (handle value, n) -> scratch bytes of a fold / scan launch

    def __call__(self, args):
        global _lazy
        dev = None
        n = -1
        sig, ptrs, keep = [], [], None
        for a in args:
            ta = type(a)
            if ta is DeviceArray:
                t = a._tensor
                if t is None or t.requires_grad:
                    return NotImplemented
                if t.dim() == 1:
                    if not t.is_contiguous():
                        return NotImplemented
                    m = t.numel()
                    if n < 0:
                        n = m
                    elif m != n:
                        return NotImplemented
                    sig.append(t.dtype)
                elif t.dim() == 0:
                    sig.append((t.dtype,))   # 0-d device value
                else:
                    return NotImplemented
                if dev is None:
                    dev = t.device
                elif t.device != dev:
                    return NotImplemented
                ptrs.append(t.data_ptr())
            elif ta is float:
                c = ctypes.c_double(a)
                keep = (keep, c)
                sig.append(float)
                ptrs.append(ctypes.addressof(c))
            elif ta is int:
                c = ctypes.c_int64(a)
                keep = (keep, c)
                sig.append(int)
                ptrs.append(ctypes.addressof(c))
            else:
                return NotImplemented
        mode = self.mode
        if dev is None or n == 0:
            return NotImplemented
        if n < 0:
            if mode != _lib.MODE_MAP:
                return NotImplemented   # a fold over scalars: interpreter semantics (the caller raises)
            scalar_out, n = True, 1
        else:
            scalar_out = mode == _lib.MODE_REDUCE
            if mode == _lib.MODE_MAP:
                if _lazy is None:
                    from . import lazy as _lazy_mod
                    _lazy = _lazy_mod
                if _lazy.ENABLED and n >= _lazy.MIN_SIZE:
                    return NotImplemented   # long vectors: the result is deferred (run)
        lib, stream = _lib_for(dev)
        key = (dev.index, tuple(sig))
        ent = self.cache.get(key)
        if ent is None:
            kinds, dts = [], []
            for x in sig:
                if isinstance(x, torch.dtype) or isinstance(x, tuple):
                    td = x[0] if isinstance(x, tuple) else x
                    if td not in _T2K_FAST:
                        return NotImplemented   # (run raises the TypeError)
                    kinds.append(_lib.ARG_DEVSCALAR if isinstance(x, tuple) else _lib.ARG_ARRAY)
                    dts.append(_T2K_FAST[td])
                else:
                    kinds.append(_lib.ARG_SCALAR)
                    dts.append(_lib.F64 if x is float else _lib.I64)
            if self.code3:
                ent = _compile_epi(lib, dev, self.code, self.code2, self.code3, tuple(kinds), tuple(dts), self.red, self.red2)
            elif self.code2:
                ent = _compile_pair(lib, dev, self.code, self.code2, tuple(kinds), tuple(dts), self.red, self.red2)
            else:
                h, od, out_td = _compile_handle(lib, dev, self.code, (), tuple(kinds), tuple(dts), mode, self.red, self.red2)
                ent = (h, out_td)
            self.cache[key] = ent
        h, out_td = ent[0], ent[1]
        if self.code3:
            out = torch.empty((), dtype=out_td, device=dev)
        elif self.code2:
            out = torch.empty(2, dtype=torch.int64, device=dev)   # two 8-byte results side by side
        else:
            out = torch.empty(() if scalar_out else (n,), dtype=out_td, device=dev)
        if mode == _lib.MODE_MAP:
            ws_ptr, wsb = None, 0
        else:
            wkey = (h.value, n)
            wsb = self.ws_bytes.get(wkey)
            if wsb is None:
                q = ctypes.c_size_t()
                _lib.check(lib.kgb_expr_workspace_bytes(h, n, ctypes.byref(q)), "kgb_expr_workspace_bytes")
                if len(self.ws_bytes) > 64:
                    self.ws_bytes.clear()
                wsb = self.ws_bytes[wkey] = q.value
            ws_ptr = _scratch(dev, stream, wsb)
        na = len(ptrs)
        argv = _ARGV[na](*ptrs) if na < len(_ARGV) else (ctypes.c_void_p * na)(*ptrs)
        rc = lib.kgb_expr_launch_ex(h, argv, n, out.data_ptr(), ws_ptr, wsb, stream, _lib.LAUNCH_WS_CLEAN)
        if rc:
            _lib.check(rc, "kgb_expr_launch")
        if self.code2 and not self.code3:
            a0, a1 = out.unbind(0)
            return (DeviceArray(a0 if out_td is torch.int64 else a0.view(out_td)),
                    DeviceArray(a1 if ent[2] is torch.int64 else a1.view(ent[2])))
        return DeviceArray(out)


def _compile_epi(lib, device, code_t, code2_t, code3_t, kinds, dts, red, red2):
    """(handle, torch dtype of the single result) of a two-output fold with an epilogue over both folds."""
    key = (device.index, code_t, code2_t, code3_t, kinds, dts, _lib.MODE_REDUCE, red, red2, "epi")
    ent = _handles.get(key)
    if ent is None:
        n_args = len(kinds)
        h = ctypes.c_void_p()
        od = ctypes.c_int32()
        _lib.check(lib.kgb_expr_compile_epi((ctypes.c_int32 * len(code_t))(*code_t), len(code_t),
                                            (ctypes.c_int32 * len(code2_t))(*code2_t), len(code2_t),
                                            (ctypes.c_int32 * len(code3_t))(*code3_t), len(code3_t),
                                            (ctypes.c_int32 * max(n_args, 1))(*kinds), (ctypes.c_int32 * max(n_args, 1))(*dts),
                                            n_args, _lib.MODE_REDUCE, red, red2, ctypes.byref(h), ctypes.byref(od)),
                   "kgb_expr_compile_epi")
        ent = (h, _K2T[od.value])
        _handles[key] = ent
    return ent


def _compile_pair(lib, device, code_t, code2_t, kinds, dts, red, red2):
    """(handle, torch dtype of output 1, of output 2) of a two-output fold (both accumulators 8 bytes wide)."""
    key = (device.index, code_t, code2_t, kinds, dts, _lib.MODE_REDUCE, red, red2, "reduce2")
    ent = _handles.get(key)
    if ent is None:
        n_args = len(kinds)
        h = ctypes.c_void_p()
        od = ctypes.c_int32()
        _lib.check(lib.kgb_expr_compile((ctypes.c_int32 * len(code_t))(*code_t), len(code_t),
                                        (ctypes.c_int32 * len(code2_t))(*code2_t), len(code2_t),
                                        (ctypes.c_int32 * max(n_args, 1))(*kinds), (ctypes.c_int32 * max(n_args, 1))(*dts),
                                        n_args, _lib.MODE_REDUCE, red, red2, ctypes.byref(h), ctypes.byref(od)),
                   "kgb_expr_compile")
        od2 = lib.kgb_expr_out_dtype2(h)
        if od2 < 0:
            raise _lib.KgbError("kgb_expr_compile did not produce a two-output reduce")
        ent = (h, od.value, od2)
        _handles[key] = ent
    return ent[0], _K2T[ent[1]], _K2T[ent[2]]


_scratch_cache = {}


def _scratch(device, stream, nbytes, owner="expr"):
    """Fold / scan scratch (partials, tile states) for one launch: ONE growing buffer per (device, stream, owner), reused
    by every launch on that stream — launches on a stream are ordered, so consecutive kernels can share it (a torch.empty
    per call was a third of the host time of a small fold).  The expression kernels' buffer is zero-filled ONCE and never
    lent to anything else: folds keep their ticket word zero between launches (KGB_LAUNCH_WS_CLEAN), scans zero what
    they use themselves.  Other kernels (stream compaction) get their own buffer."""
    key = (device.index, stream.value if stream is not None else None, owner)
    buf = _scratch_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)
        _scratch_cache[key] = buf
    return buf.data_ptr()


def reduce2(code, red, code2, red2, args):
    """Two folds over the same arguments in ONE pass (kgb200.h: KGB_MODE_REDUCE with a second program).
    Returns two 0-d DeviceArrays.  Both accumulators must be 8 bytes wide (float64 / int64)."""
    device = None
    n = None
    for a in args:
        if isinstance(a, DeviceArray):
            device = a.device
            if a.ndim > 0:
                m = a.size
                if a.ndim != 1 or (n is not None and m != n):
                    raise ValueError("reduce2: operands must be 1-d arrays of one length")
                n = m
    if device is None or n is None or n == 0:
        raise ValueError("reduce2 needs a non-empty device array operand")
    red = RED[red] if not isinstance(red, int) else red
    red2 = RED[red2] if not isinstance(red2, int) else red2
    if _sharded.ACTIVE and any(_sharded.is_sharded(a) for a in args):
        r = _sharded.reduce2(code, red, code2, red2, args)
        if r is not NotImplemented:
            return r
    if torch.is_grad_enabled() and any(isinstance(a, DeviceArray) and a._t.requires_grad for a in args):
        autograd = _autograd_mod or _bind_autograd()  # one two-output node on the tape (raises NotDifferentiable for folds other than +/)
        return autograd.run_differentiable2(code, red, code2, red2, args, device)
    lib, stream = _lib_for(device)
    cls = [_classify(a, device) for a in args]
    kinds, dts = tuple(c[0] for c in cls), tuple(c[1] for c in cls)
    code_t, code2_t = tuple(code), tuple(code2)
    key = (device.index, code_t, code2_t, kinds, dts, _lib.MODE_REDUCE, red, red2, "reduce2")
    ent = _handles.get(key)
    if ent is None:
        n_args = len(args)
        h = ctypes.c_void_p()
        od = ctypes.c_int32()
        _lib.check(lib.kgb_expr_compile((ctypes.c_int32 * len(code_t))(*code_t), len(code_t),
                                        (ctypes.c_int32 * len(code2_t))(*code2_t), len(code2_t),
                                        (ctypes.c_int32 * max(n_args, 1))(*kinds), (ctypes.c_int32 * max(n_args, 1))(*dts),
                                        n_args, _lib.MODE_REDUCE, red, red2, ctypes.byref(h), ctypes.byref(od)),
                   "kgb_expr_compile")
        od2 = lib.kgb_expr_out_dtype2(h)
        if od2 < 0:
            raise _lib.KgbError("kgb_expr_compile did not produce a two-output reduce")
        ent = (h, od.value, od2)
        _handles[key] = ent
    h, od, od2 = ent
    buf = torch.empty(16, dtype=torch.uint8, device=device)
    wsb = ctypes.c_size_t()
    _lib.check(lib.kgb_expr_workspace_bytes(h, n, ctypes.byref(wsb)), "kgb_expr_workspace_bytes")
    argv = (ctypes.c_void_p * max(len(args), 1))(*[c[3] for c in cls])
    _lib.check(lib.kgb_expr_launch_ex(h, argv, n, buf.data_ptr(), _scratch(device, stream, wsb.value), wsb.value, stream,
                                      _lib.LAUNCH_WS_CLEAN), "kgb_expr_launch")
    return DeviceArray(buf[:8].view(_K2T[od]).reshape(())), DeviceArray(buf[8:].view(_K2T[od2]).reshape(()))


def empty_fold(code, args, red):
    """Identity of an add / multiply fold over zero elements, in the dtype the expression would produce."""
    m = run(code, args)  # zero-length MAP launch: compiles, allocates, launches nothing
    t = m._t
    dt = torch.int64 if t.dtype in (torch.bool, torch.int32) else t.dtype
    fill = torch.zeros if red == _lib.RED_ADD else torch.ones
    return DeviceArray(fill((), dtype=dt, device=t.device))


# ---------------------------------------------------------------------------------------- eager ops
def _device_of(*xs):
    for x in xs:
        if isinstance(x, DeviceArray):
            return x.device  # metadata only: never materialises a deferred operand
    return None


def _lift(x, device):
    """Host numeric ndarrays (ndim>0) become device arrays; everything else is passed through."""
    if isinstance(x, np.ndarray) and x.ndim > 0:
        return asarray(x, device=device)
    if isinstance(x, (list, tuple)):
        return asarray(x, device=device)
    return x


_fast_dyads, _fast_monads, _fast_folds, _fast_scans = {}, {}, {}, {}


def binary(op, a, b):
    """Eager dyad: one-node program (dyads.py:6-21, 214-232, 588-803 call sites)."""
    if op != "pow":   # (an integer power may need its operands converted first)
        fs = _fast_dyads.get(op)
        if fs is None:
            fs = _fast_dyads[op] = FastStage((OP_ARG, 0, OP_ARG, 1, OPC[op]), _lib.MODE_MAP, 0)
        r = fs((a, b))   # short 1-d device vectors / Python numbers: the pre-resolved launcher
        if r is not NotImplemented:
            return r
    dev = _device_of(a, b)
    if dev is None:
        if is_host_scalar(a) and is_host_scalar(b):
            return _HOST[op](a, b)
        dev = default_device()
    a, b = _lift(a, dev), _lift(b, dev)
    if op == "pow":
        a, b = _pow_operands(a, b)
    return run([OP_ARG, 0, OP_ARG, 1, OPC[op]], [a, b])


def _pow_operands(a, b):
    # int ** negative int is a ValueError in numpy; like the reference's other non-numpy backends
    # (tests/custom_backend.py:121-138, torch_backend.py:667-682) compute it in floating point.
    def _is_int(x):
        return (isinstance(x, DeviceArray) and x.dtype.kind in "iub") or isinstance(x, (int, np.integer))
    if _is_int(a) and _is_int(b):
        neg = (b < 0) if not isinstance(b, DeviceArray) else bool(any_(binary("lt_b", b, 0)))
        if neg:
            a = astype(a, np.float64) if isinstance(a, DeviceArray) else float(a)
    return a, b


def unary(op, a):
    fs = _fast_monads.get(op)
    if fs is None:
        fs = _fast_monads[op] = FastStage((OP_ARG, 0, OPC[op]), _lib.MODE_MAP, 0)
    r = fs((a,))
    if r is not NotImplemented:
        return r
    dev = _device_of(a)
    if dev is None:
        if is_host_scalar(a):
            return _HOST[op](a)
        a = _lift(a, default_device())
    return run([OP_ARG, 0, OPC[op]], [a])


def full(shape, value, dtype=None, device=None):
    """np.full on the device: a MAP program over scalars only."""
    dev = torch.device(device) if device is not None else default_device()
    shape = (int(shape),) if isinstance(shape, (int, np.integer)) else tuple(int(s) for s in shape)
    if isinstance(value, DeviceArray):
        value = value.item()
    if dtype is None:
        dtype = np.int64 if isinstance(value, (int, np.integer, bool, np.bool_)) else np.float64
    d = np.dtype(dtype)
    if isinstance(value, (bool, np.bool_)):
        value = int(value)
    return run([OP_ARG, 0, OPC[_CAST[d]]], [value], out_shape=shape, device=dev)


def where3(c, a, b):
    dev = _device_of(c, a, b) or default_device()
    c, a, b = _lift(c, dev), _lift(a, dev), _lift(b, dev)
    return run([OP_ARG, 0, OP_ARG, 1, OP_ARG, 2, OPC["where"]], [c, a, b])


_CAST = {np.dtype("float64"): "to_f64", np.dtype("int64"): "to_i64", np.dtype("float32"): "to_f32",
         np.dtype("int32"): "to_i32", np.dtype("bool"): "to_b8"}


def astype(a, dtype):
    if dtype in (int, "int"):
        dtype = np.int64
    elif dtype in (float, "float"):
        dtype = np.float64
    elif dtype in (bool, "bool"):
        dtype = np.bool_
    d = np.dtype(dtype)
    if d not in _CAST:
        d = np.dtype(np.int64) if d.kind in "iu" else np.dtype(np.float64) if d.kind == "f" else d
    if d not in _CAST:
        raise TypeError(f"cannot cast device array to {dtype}")
    if not isinstance(a, DeviceArray):
        return np.asarray(a).astype(d)[()]
    if a.dtype == d:
        return copy(a)
    return run([OP_ARG, 0, OPC[_CAST[d]]], [a])


def copy(a):
    if a.size == 0:
        return DeviceArray(a._t.clone())
    return run([OP_ARG, 0], [a])


def reduce(op, a, code=None, args=None):
    """op-fold over a 1-d device array (adverbs.py:232-243); N-d folds along axis 0 like numpy ufunc.reduce."""
    fs = _fast_folds.get(op)
    if fs is None:
        fs = _fast_folds[op] = FastStage((OP_ARG, 0), _lib.MODE_REDUCE, RED[op])
    r = fs((a,))
    if r is not NotImplemented:
        return r
    if a.ndim == 0:
        return a
    if a.ndim > 1:
        return reduce_axis0([OP_ARG, 0], [a], RED[op])
    if a.size == 0:
        if op == "add":
            return DeviceArray(torch.zeros((), dtype=a._t.dtype if a._t.dtype != torch.bool else torch.int64,
                                           device=a.device))
        if op == "mul":
            return DeviceArray(torch.ones((), dtype=a._t.dtype if a._t.dtype != torch.bool else torch.int64,
                                          device=a._t.device))
        raise ValueError("zero-size array to reduction operation which has no identity")
    return run([OP_ARG, 0], [a], mode=_lib.MODE_REDUCE, red=RED[op])


_RED_OPNAME = {_lib.RED_ADD: "add", _lib.RED_MUL: "mul", _lib.RED_MAX: "max", _lib.RED_MIN: "min"}
_AXIS0_ARGS = 12  # KGB_MAX_ARGS of the expression kernels


def reduce_axis0(code, args, red):
    """Fold of the elementwise program `code` over N-d operands of one shape along axis 0 — what `ufunc.reduce` does
    (numpy_backend.py:163; adverbs.py:232-243) — as FUSED kernels over row views: the rows of up to 11 operands' worth
    of arguments are separate arguments of one MAP program  f(row 0) op f(row 1) op ...  so every element is read once
    and only the running result is re-read between launches (the round-1 path was one launch and one temporary per row)."""
    nd = [i for i, a in enumerate(args) if isinstance(a, DeviceArray) and a.ndim > 1]
    if not nd:
        raise ValueError("reduce_axis0 needs an N-d operand")
    shape = args[nd[0]].shape
    for i in nd:
        if args[i].shape != shape:
            raise NotImplementedError("axis-0 fold over operands of different shapes")
    for a in args:
        if isinstance(a, DeviceArray) and a.ndim == 1:
            raise NotImplementedError("axis-0 fold mixing 1-d and N-d operands")
    rows_total = shape[0]
    if rows_total == 0:
        raise NotImplementedError("axis-0 fold of an empty leading axis")
    opc = OPC[_RED_OPNAME[red]]
    shared = [i for i in range(len(args)) if i not in nd]
    per = max(1, (_AXIS0_ARGS - len(shared) - 1) // len(nd))
    tensors = {i: contiguous(args[i])._t for i in nd}
    code = list(code)
    acc = None
    r = 0
    while r < rows_total:
        take = min(per, rows_total - r)
        largs = [args[i] for i in shared]
        slot_of_shared = {i: k for k, i in enumerate(shared)}
        prog = []
        for j in range(take):
            base = len(largs)
            for i in nd:
                largs.append(DeviceArray(tensors[i][r + j]))
            slot_of_nd = {i: base + k for k, i in enumerate(nd)}
            w = 0
            while w < len(code):
                c = code[w]
                if c == OP_ARG:
                    i = code[w + 1]
                    prog += [OP_ARG, slot_of_nd[i] if i in slot_of_nd else slot_of_shared[i]]
                    w += 2
                elif c in (OP_LIT_I64, OP_LIT_F64):
                    prog += code[w:w + 3]
                    w += 3
                else:
                    prog.append(c)
                    w += 1
            if j > 0:
                prog.append(opc)
        if acc is not None:
            largs.append(acc)
            prog = [OP_ARG, len(largs) - 1] + prog + [opc]
        acc = run(prog, largs, _no_defer=True)
        r += take
    if red in (_lib.RED_ADD, _lib.RED_MUL) and acc._t.dtype in (torch.bool, torch.int32):
        acc = astype(acc, np.int64)  # np.add.reduce / np.multiply.reduce promote bool and small ints
    return acc


def scan(op, a):
    """inclusive op-scan (adverbs.py:325-332); N-d scans along axis 0."""
    fs = _fast_scans.get(op)
    if fs is None:
        fs = _fast_scans[op] = FastStage((OP_ARG, 0), _lib.MODE_SCAN, RED[op])
    r = fs((a,))
    if r is not NotImplemented:
        return r
    if a.ndim == 0:
        return a
    if a.ndim > 1:
        rows = [DeviceArray(a._t[0])]
        for i in range(1, a._t.shape[0]):
            rows.append(binary(op, rows[-1], DeviceArray(a._t[i])))
        out = torch.empty_like(a._t if a._t.dtype == rows[-1]._t.dtype else a._t.to(rows[-1]._t.dtype))
        for i, r in enumerate(rows):
            out[i].copy_(r._t)
        return DeviceArray(out)
    if a.size == 0:
        return DeviceArray(a._t.clone())
    return run([OP_ARG, 0], [a], mode=_lib.MODE_SCAN, red=RED[op])


def all_(a):
    if a.size == 0:
        return True
    flat = a.flatten()
    return run([OP_ARG, 0, OPC["to_b8"]], [flat], mode=_lib.MODE_REDUCE, red=_lib.RED_MIN)


def any_(a):
    if a.size == 0:
        return False
    flat = a.flatten()
    return run([OP_ARG, 0, OPC["to_b8"]], [flat], mode=_lib.MODE_REDUCE, red=_lib.RED_MAX)


# ---------------------------------------------------------------------------------------- AOT kernels
def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 64), dtype=torch.uint8, device=device)


def argsort(a, descending=False, return_sorted=False):
    """Stable grade (writer.py:97-122 -> numpy_backend.py:75-80)."""
    a = contiguous(a)
    if a._t.dim() != 1:
        raise NotImplementedError("argsort: 1-d arrays only")
    dev = a._t.device
    lib, stream = _lib_for(dev)
    if a._t.dtype == torch.bool:
        a = astype(a, np.int64)
    kdt = kgb_dtype(a._t)
    n = a._t.numel()
    idx = torch.empty(n, dtype=torch.int64, device=dev)
    ks = torch.empty_like(a._t) if return_sorted else None
    if n:
        wsb = ctypes.c_size_t()
        _lib.check(lib.kgb_argsort_workspace_bytes(kdt, n, ctypes.byref(wsb)))
        ws = _ws(wsb.value, dev)
        _lib.check(lib.kgb_argsort(a._t.data_ptr(), kdt, n, 1 if descending else 0, idx.data_ptr(),
                                   ks.data_ptr() if ks is not None else None, ws.data_ptr(), wsb.value, stream),
                   "kgb_argsort")
    if return_sorted:
        return DeviceArray(idx), DeviceArray(ks)
    return DeviceArray(idx)


def sort_values(a, descending=False):
    """`a@<a` / `a@>a`: the sorted keys, written by the last radix pass itself (no gather through the permutation)."""
    return argsort(a, descending=descending, return_sorted=True)[1]


def grade_deferred(a, descending=False):
    """`<a` / `>a` as a deferred permutation (lazy.GradeNode): materialised by any use except `a@<a`."""
    from . import lazy
    if a._t.dim() != 1 or a._t.numel() == 0 or a._t.dtype == torch.bool:
        return argsort(a, descending=descending)
    return DeviceArray._deferred(lazy.GradeNode(a, descending))


def grade_source(b):
    """(source array, descending) when `b` is a still-deferred grade, else None."""
    from . import lazy
    if isinstance(b, DeviceArray) and b._tensor is None and isinstance(b._node, lazy.GradeNode):
        return b._node.source, b._node.descending
    return None


def group(a):
    """`=a` (monads.py:182-204): -> (order int64[n], starts int64[G+1]) with one D2H read of G."""
    a = contiguous(a)
    dev = a._t.device
    lib, stream = _lib_for(dev)
    if a._t.dtype == torch.bool:
        a = astype(a, np.int64)
    kdt = kgb_dtype(a._t)
    n = a._t.numel()
    order = torch.empty(n, dtype=torch.int64, device=dev)
    starts = torch.empty(n + 1, dtype=torch.int64, device=dev)
    ng = torch.zeros(1, dtype=torch.int64, device=dev)
    wsb = ctypes.c_size_t()
    _lib.check(lib.kgb_group_workspace_bytes(kdt, n, ctypes.byref(wsb)))
    ws = _ws(wsb.value, dev)
    _lib.check(lib.kgb_group(a._t.data_ptr(), kdt, n, order.data_ptr(), starts.data_ptr(), ng.data_ptr(),
                             ws.data_ptr(), wsb.value, stream), "kgb_group")
    g = int(ng.item())
    return DeviceArray(order), DeviceArray(starts[: g + 1]), g


_slots = threading.local()


def _count_slot(dev):
    """One 8-byte result slot per (thread, device): pinned host memory on a GPU, plain memory on the test double."""
    d = getattr(_slots, "d", None)
    if d is None:
        d = _slots.d = {}
    t = d.get(dev)
    if t is None:
        t = d[dev] = torch.zeros(1, dtype=torch.int64, pin_memory=(dev.type == "cuda"))
    return t


def _select(fn_name, a, needle=None):
    a = contiguous(a)
    dev = a._t.device
    lib, stream = _lib_for(dev)
    kdt = kgb_dtype(a._t)
    n = a._t.numel()
    if not n:
        return DeviceArray(torch.empty(0, dtype=torch.int64, device=dev))
    pos = torch.empty(n, dtype=torch.int64, device=dev)
    # the match count is the one value the host needs: the offsets kernel writes it straight into pinned (device-visible)
    # host memory and the host POLLS that word — it does not wait for the emit kernel that follows on the stream (the
    # positions are consumed in stream order anyway), and no D2H copy is involved
    cnt = _count_slot(dev)
    cnt_np = cnt.numpy()
    cnt_np[0] = -1
    wsb = ctypes.c_size_t()
    _lib.check(lib.kgb_select_workspace_bytes(n, ctypes.byref(wsb)))
    ws_ptr = _scratch(dev, stream, wsb.value, "select")
    if fn_name == "find":
        _lib.check(lib.kgb_find_equal(a._t.data_ptr(), kdt, n, ctypes.addressof(needle), pos.data_ptr(),
                                      cnt.data_ptr(), ws_ptr, wsb.value, stream), "kgb_find_equal")
    else:
        _lib.check(lib.kgb_nonzero(a._t.data_ptr(), kdt, n, pos.data_ptr(), cnt.data_ptr(), ws_ptr, wsb.value, stream),
                   "kgb_nonzero")
    if dev.type == "cuda":
        spins = 0
        while cnt_np[0] < 0:
            spins += 1
            if spins > 2_000_000:  # (seconds) the kernel must have failed: let the synchronise raise the CUDA error
                torch.cuda.current_stream(dev).synchronize()
                break
    c = int(cnt_np[0])
    if c < 0:
        raise _lib.KgbError("select: the match count never arrived")
    res = pos[:c]
    return DeviceArray(res.clone() if 2 * c < n else res)  # do not pin n words of storage under a short result


def find_equal(a, b):
    """`a?b` for an atom b (dyads.py:342): ascending positions where a == b."""
    a = a.flatten() if a._t.dim() != 1 else a
    if a._t.dtype == torch.bool:
        a = astype(a, np.int64)
    if isinstance(b, DeviceArray):
        b = b.item()
    k = a.dtype.kind
    if k == "f":
        if a._t.dtype == torch.float32:
            needle = ctypes.c_float(float(b))
        else:
            needle = ctypes.c_double(float(b))
    else:
        # np.where(a == b)[0] is empty when an integer array cannot hold b: fractional, non-finite or out of range
        empty = DeviceArray(torch.empty(0, dtype=torch.int64, device=a._t.device))
        if isinstance(b, (float, np.floating)):
            if not np.isfinite(b) or float(b) != int(b):
                return empty
        info = np.iinfo(np.int64 if a._t.dtype == torch.int64 else np.int32)
        if not (info.min <= int(b) <= info.max):
            return empty
        needle = ctypes.c_int64(int(b)) if a._t.dtype == torch.int64 else ctypes.c_int32(int(b))
    return _select("find", a, needle)


def nonzero(a):
    """np.where(flags)[0]"""
    a = a.flatten() if a._t.dim() != 1 else a
    return _select("nonzero", a)


def unique_first(a):
    """`?a` (monads.py:260-295): unique values in first-appearance order."""
    a = contiguous(a)
    dev = a._t.device
    lib, stream = _lib_for(dev)
    was_bool = a._t.dtype == torch.bool
    if was_bool:
        a = astype(a, np.int64)
    kdt = kgb_dtype(a._t)
    n = a._t.numel()
    if n == 0:
        return a
    out = torch.empty_like(a._t)
    first = torch.empty(n, dtype=torch.int64, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    wsb = ctypes.c_size_t()
    _lib.check(lib.kgb_unique_first_workspace_bytes(kdt, n, ctypes.byref(wsb)))
    ws = _ws(wsb.value, dev)
    _lib.check(lib.kgb_unique_first(a._t.data_ptr(), kdt, n, out.data_ptr(), first.data_ptr(), cnt.data_ptr(),
                                    ws.data_ptr(), wsb.value, stream), "kgb_unique_first")
    c = int(cnt.item())
    return DeviceArray(out[:c])


def iota(n, device=None, start=0, step=1):
    """`!n` (monads.py:52)."""
    dev = torch.device(device) if device is not None else default_device()
    lib, stream = _lib_for(dev)
    out = torch.empty(int(n), dtype=torch.int64, device=dev)
    if n:
        _lib.check(lib.kgb_iota_i64(out.data_ptr(), int(n), int(start), int(step), stream), "kgb_iota_i64")
    return DeviceArray(out)


def gather(a, idx):
    """`a@b` for an integer index list (dyads.py:176-181)."""
    if torch.is_grad_enabled() and a._t.requires_grad:
        # autograd interop only: the backward of a gather is a scatter-add torch already knows
        return DeviceArray(a._t[idx._t.to(torch.int64)])
    a = contiguous(a)
    idx = contiguous(idx)
    dev = a._t.device
    lib, stream = _lib_for(dev)
    if idx._t.dtype != torch.int64:
        idx = astype(idx, np.int64)
    n_idx = idx._t.numel()
    if a._t.dim() == 0:
        raise IndexError("too many indices for array")
    row = 1
    for s in a._t.shape[1:]:
        row *= s
    out = torch.empty(tuple(idx._t.shape) + tuple(a._t.shape[1:]), dtype=a._t.dtype, device=dev)
    if n_idx and row:
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(lib.kgb_gather(a._t.data_ptr(), a._t.element_size() * row, a._t.shape[0], idx._t.data_ptr(),
                                  n_idx, out.data_ptr(), err.data_ptr(), stream), "kgb_gather")
        if int(err.item()):
            raise IndexError(f"index out of bounds for axis 0 with size {a._t.shape[0]}")
    return DeviceArray(out)


def reverse(a):
    a = contiguous(a)
    dev = a._t.device
    lib, stream = _lib_for(dev)
    out = torch.empty_like(a._t)
    n = a._t.shape[0] if a._t.dim() else 1
    row = a._t.numel() // n if n else 0
    if n and row:
        _lib.check(lib.kgb_reverse(a._t.data_ptr(), a._t.element_size() * row, n, out.data_ptr(), stream),
                   "kgb_reverse")
    return DeviceArray(out)


def expand(counts):
    """`&a` (monads.py:79): np.repeat(np.arange(len(a)), a)."""
    counts = contiguous(counts)
    if counts._t.dtype != torch.int64:
        counts = astype(counts, np.int64)
    dev = counts._t.device
    n = counts._t.numel()
    if n == 0:
        return DeviceArray(torch.empty(0, dtype=torch.int64, device=dev))
    if bool(any_(binary("lt_b", counts, 0))):
        raise ValueError("repeats may not contain negative values.")
    incl = scan("add", counts)
    total = int(DeviceArray(incl._t[-1]).item())
    lib, stream = _lib_for(dev)
    out = torch.empty(total, dtype=torch.int64, device=dev)
    if total:
        _lib.check(lib.kgb_expand(counts._t.data_ptr(), incl._t.data_ptr(), n, total, out.data_ptr(), stream),
                   "kgb_expand")
    return DeviceArray(out)


def array_equal(a, b):
    """backend.array_equal (base.py:222-224): exact elementwise equality, one 4-byte D2H read."""
    if tuple(a._t.shape) != tuple(b._t.shape):
        return False
    if a._t.numel() == 0:
        return True
    if a._t.dtype != b._t.dtype:
        d = np.result_type(a.dtype, b.dtype)
        a, b = astype(a, d), astype(b, d)
    a, b = contiguous(a), contiguous(b)
    dev = a._t.device
    lib, stream = _lib_for(dev)
    flag = torch.empty(1, dtype=torch.int32, device=dev)
    _lib.check(lib.kgb_array_equal(a._t.data_ptr(), b._t.data_ptr(), kgb_dtype(a._t), a._t.numel(),
                                   flag.data_ptr(), stream), "kgb_array_equal")
    return bool(flag.item())


def groupagg(keys, vals, n_groups, tables=None, accumulate=None, defer_check=False):
    """1brc-style grouped (min, max, sum, count) (examples/1brc/worker.kg:3-5).  `tables`: four preallocated result
    arrays; they are merged into (the `merge` step) unless accumulate=False.  defer_check=True does not wait for the
    "key outside [0, n_groups)" flag: -> (tables, check) where check() collects it later (and raises) — the caller can
    enqueue more work behind the aggregation first (dist.sharded_groupagg: the exchange)."""
    if _sharded.ACTIVE and (_sharded.is_sharded(keys) or _sharded.is_sharded(vals)) and tables is None:
        r = _sharded.groupagg(keys, vals, n_groups)
        if r is not NotImplemented:
            return r
    keys, vals = contiguous(keys), contiguous(vals)
    dev = keys._t.device
    lib, stream = _lib_for(dev)
    if keys._t.dtype != torch.int64:
        keys = astype(keys, np.int64)
    if vals._t.dtype != torch.float64:
        vals = astype(vals, np.float64)
    n = keys._t.numel()
    if vals._t.numel() != n:
        raise ValueError("groupagg: keys and values differ in length")
    G = int(n_groups)
    acc = (tables is not None) if accumulate is None else bool(accumulate)
    if tables is None:
        tables = (torch.empty(G, dtype=torch.float64, device=dev), torch.empty(G, dtype=torch.float64, device=dev),
                  torch.empty(G, dtype=torch.float64, device=dev), torch.empty(G, dtype=torch.int64, device=dev))
    else:
        tables = tuple(t._t if isinstance(t, DeviceArray) else t for t in tables)
    wsb = ctypes.c_size_t()
    _lib.check(lib.kgb_groupagg_workspace_bytes(n, G, ctypes.byref(wsb)))
    ws = _ws(wsb.value, dev)
    if defer_check:
        _lib.check(lib.kgb_groupagg_f64_async(keys._t.data_ptr(), vals._t.data_ptr(), n, G, tables[0].data_ptr(),
                                              tables[1].data_ptr(), tables[2].data_ptr(), tables[3].data_ptr(),
                                              1 if acc else 0, ws.data_ptr(), wsb.value, stream), "kgb_groupagg_f64_async")

        def check(ws=ws, keep=(keys, vals)):
            _lib.check(lib.kgb_groupagg_status(ws.data_ptr(), G, stream), "kgb_groupagg_f64")
        return tuple(DeviceArray(t) for t in tables), check
    _lib.check(lib.kgb_groupagg_f64(keys._t.data_ptr(), vals._t.data_ptr(), n, G, tables[0].data_ptr(),
                                    tables[1].data_ptr(), tables[2].data_ptr(), tables[3].data_ptr(),
                                    1 if acc else 0, ws.data_ptr(), wsb.value, stream), "kgb_groupagg_f64")
    return tuple(DeviceArray(t) for t in tables)


# ---------------------------------------------------------------------------------------- indexing
def getitem(a, key):
    t = a._t
    if isinstance(key, DeviceArray):
        if key._t.dim() == 0:
            key = key.item()
        elif key._t.dtype == torch.bool:
            return gather(a, nonzero(key))
        else:
            return gather(a, key)
    if isinstance(key, (bool, np.bool_)):
        raise IndexError("boolean scalar index")
    if isinstance(key, (int, np.integer)):
        return DeviceArray(t[int(key)])
    if isinstance(key, float) and float(key).is_integer():
        return DeviceArray(t[int(key)])
    if isinstance(key, slice):
        if key.step is not None and key.step < 0:
            n = t.shape[0]
            start, stop, step = key.indices(n)
            if step == -1 and key.start is None and key.stop is None:
                return reverse(a)
            count = len(range(start, stop, step))
            return gather(a, iota(count, t.device, start, step))
        return DeviceArray(t[key])
    if isinstance(key, (list, np.ndarray)):
        k = np.asarray(key)
        if k.dtype == bool:
            return gather(a, nonzero(asarray(k, device=t.device)))
        if k.size == 0:
            return DeviceArray(t[:0])
        return gather(a, asarray(k.astype(np.int64), device=t.device))
    if isinstance(key, tuple):
        if all(isinstance(k, (int, np.integer, slice)) or k is None or k is Ellipsis for k in key) and \
                not any(isinstance(k, slice) and k.step is not None and k.step < 0 for k in key):
            return DeviceArray(t[tuple(int(k) if isinstance(k, np.integer) else k for k in key)])
        raise NotImplementedError("advanced tuple indexing on device arrays")
    if key is None or key is Ellipsis:
        return DeviceArray(t[key])
    raise IndexError(f"unsupported index type {type(key).__name__}")
