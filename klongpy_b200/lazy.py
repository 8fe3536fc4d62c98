"""Deferred elementwise expressions — SURVEY.md §8f rank 2 ("compiler coverage").  ON by default (KGB200_LAZY=0 or
`lazy.enable(False)` switch it off).

The reference's compiler IR covers `+ - * % ^ = < > -x +/ +\\` only (`compiler.py:77-90`); everything else — `& | !
_`, `%x`, and the `.bkf` unary math the autograd examples live on (`exp tanh sqrt ...`, `sys_fn.py:715-723`) — reaches
the backend one operator at a time, so `sigmoid((w1*X)+b1)` is five kernels and four temporaries.  With the lazy
layer an eager elementwise op returns a `DeviceArray` that only *describes* its result (postfix program + leaf
operands + inferred shape/dtype via `kgb_expr_infer`); consumers that are themselves elementwise, reductions or
scans inline that description into their own program, and anything that touches the data (`to_numpy`, indexing,
sort, an operand used a second time) materialises it with ONE fused kernel.  Under autograd a fused program is a
single tape node whose backward is derived symbolically (autodiff.py), so the backward fuses too.

Rules that keep it safe:
  * shapes are checked and dtypes inferred when the node is built — errors are raised at the operator, not later
    (elementwise kernels have no data-dependent errors: x%0 is inf/nan like NumPy, never an exception);
  * a node inlines into at most ONE consumer; a second consumer materialises it first (no exponential recompute);
  * programs are capped (arguments <= KGB_MAX_ARGS, code <= 96 words); beyond that operands are materialised.
"""
import ctypes
import os

import torch

from . import _lib

ENABLED = os.environ.get("KGB200_LAZY", "1") != "0"   # on by default since round 2 (KGB200_LAZY=0 switches it off)
MAX_ARGS = 12
MAX_CODE = 96
# Deferring costs ~8 us of host bookkeeping per operator; an extra elementwise pass over n float64 elements costs the GPU
# 24 n bytes / 6.5 TB/s, i.e. the same 8 us at n ~ 2 M.  Below that an immediate launch is cheaper than the chance to fuse
# (bench_compiler.kg's 100 K-element rows: `a+b` 11 us launched at once, 19 us deferred and then materialised), so small
# vectors launch eagerly; 0-d results stay deferred (a chain of scalar operators becomes one launch instead of one each).
MIN_SIZE = int(os.environ.get("KGB200_LAZY_MIN", str(1 << 21)))
_infer_cache = {}


def enable(flag=True):
    global ENABLED
    ENABLED = bool(flag)


class Node:
    __slots__ = ("code", "args", "shape", "kdt", "device", "uses")

    def __init__(self, code, args, shape, kdt, device):
        self.code, self.args, self.shape, self.kdt, self.device = tuple(code), list(args), tuple(shape), kdt, device
        self.uses = 0


class GradeNode:
    """`<a` / `>a` that has not been launched yet (always on, independent of ENABLED).  Klong sorts with `a@<a`
    (monads.py:143-179: "To sort a list a, use a@<a"): when the permutation's only consumer is that `@` on the same
    array, the sorted keys come straight out of the last radix pass (kgb_argsort's keys_sorted_out) and the 24 B/elem
    random gather never runs.  Any other use materialises the permutation with the ordinary grade."""
    __slots__ = ("source", "descending", "shape", "kdt", "device", "uses", "args", "code")

    def __init__(self, source, descending):
        self.source, self.descending = source, bool(descending)
        self.shape, self.kdt, self.device = tuple(source.shape), _lib.I64, source.device
        self.uses = 1   # never inlined into an elementwise program: consumers materialise it
        self.args, self.code = [], ()


_DA = None   # klongpy_b200.array.DeviceArray / klongpy_b200.ops, bound on first use (they import this module's users)
_OPS = None


def _bind():
    global _DA, _OPS
    from . import ops
    from .array import DeviceArray
    _DA, _OPS = DeviceArray, ops
    return ops


def _is_lazy(a):
    if _DA is None:
        _bind()
    return isinstance(a, _DA) and a._tensor is None


def inline(code, args):
    """Expand deferred operands into `code` (postfix over `args`).  Returns (code, args) over leaf operands only."""
    if _DA is None:
        _bind()
    DA = _DA
    lz = [isinstance(a, DA) and a._tensor is None for a in args]   # which operands are deferred (looked at once)
    if True not in lz:
        return list(code), list(args)
    ops = _OPS
    OP_ARG, LIT_I, LIT_F = ops.OP_ARG, ops.OP_LIT_I64, ops.OP_LIT_F64
    new_args, index_of = [], {}

    def leaf(a):
        k = id(a)
        i = index_of.get(k)
        if i is None:
            i = index_of[k] = len(new_args)
            new_args.append(a)
        return i

    # decide which deferred operands can be inlined (first consumer, and the result must stay within the caps)
    budget_args = MAX_ARGS - lz.count(False)
    budget_code = MAX_CODE - len(code)
    plan = []
    for a, z in zip(args, lz):
        if z:
            nd = a._node
            ok = nd.uses == 0 and len(nd.args) <= budget_args and len(nd.code) <= budget_code
            if ok:
                budget_args -= len(nd.args)
                budget_code -= len(nd.code)
            plan.append(ok)
        else:
            plan.append(False)
    out, i = [], 0
    code = list(code)
    ncode = len(code)
    while i < ncode:
        w = code[i]
        if w == OP_ARG:
            ai = code[i + 1]
            a = args[ai]
            if plan[ai]:
                nd = a._node
                nd.uses += 1
                sub, j = nd.code, 0
                nsub = len(sub)
                nargs = nd.args
                while j < nsub:
                    v = sub[j]
                    if v == OP_ARG:
                        out.append(OP_ARG)
                        out.append(leaf(nargs[sub[j + 1]]))
                        j += 2
                    elif v == LIT_I or v == LIT_F:
                        out += sub[j:j + 3]
                        j += 3
                    else:
                        out.append(v)
                        j += 1
            else:
                if lz[ai]:
                    a._t  # noqa: B018 - materialise: used twice, or the fused program would get too large
                out.append(OP_ARG)
                out.append(leaf(a))
            i += 2
        elif w == LIT_I or w == LIT_F:
            out += code[i:i + 3]
            i += 3
        else:
            out.append(w)
            i += 1
    return out, new_args


def _infer(code, args, device):
    ops = _OPS or _bind()
    lib, _ = ops._lib_for(device)
    cls = [ops._classify_meta(a) for a in args]
    key = (tuple(code), tuple(cls))
    kdt = _infer_cache.get(key)
    if kdt is None:
        n = len(args)
        od = ctypes.c_int32()
        _lib.check(lib.kgb_expr_infer((ctypes.c_int32 * len(code))(*code), len(code),
                                      (ctypes.c_int32 * max(n, 1))(*[c[0] for c in cls]),
                                      (ctypes.c_int32 * max(n, 1))(*[c[1] for c in cls]), n, ctypes.byref(od)),
                   "kgb_expr_infer")
        kdt = od.value
        _infer_cache[key] = kdt
    return kdt


def defer(code, args):
    """A deferred DeviceArray for the MAP program `code` over leaf operands `args`, or None when it cannot be
    deferred (no device operand, too many arguments)."""
    if _DA is None:
        _bind()
    DeviceArray = _DA
    device, shape = None, None
    for a in args:
        if isinstance(a, DeviceArray):
            device = a.device
            if a.ndim > 0:
                if shape is None:
                    shape = a.shape
                elif a.shape != shape:
                    return None  # NumPy broadcasting between different shapes: the eager path expands (or raises)
    if device is None or len(args) > MAX_ARGS or len(code) > MAX_CODE:
        return None
    if shape is not None:
        size = 1
        for d in shape:
            size *= d
        if 1 < size < MIN_SIZE or size == 0:
            return None
    kdt = _infer(code, args, device)
    return DeviceArray._deferred(Node(code, args, shape if shape is not None else (), kdt, device))


def materialize(node):
    """Launch the deferred program: one fused kernel (and one autograd node when an operand is on the tape)."""
    from . import ops
    if isinstance(node, GradeNode):
        return ops.argsort(node.source, descending=node.descending)._t
    from . import sharded
    if isinstance(node, sharded.ShardNode):
        return sharded.gather(node)   # no sharded kernel for this consumer: all blocks onto the first device
    return ops.run(node.code, node.args, _no_defer=True)._t
