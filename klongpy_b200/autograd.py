"""Autograd interop (BASELINE config 5).

The reference's torch backend differentiates by running the interpreter on `requires_grad` tensors and letting
ATen record one node per operator (`torch_backend.py:692-782`; inputs are cast to float32, `:692-699`).  Here
torch's tape is only the *bookkeeping*: every libkgb200 expression launch whose operands require grad is recorded
as ONE `torch.autograd.Function` node whose backward launches the fused gradient programs derived symbolically
by `autodiff.py` (one fused kernel per differentiated operand, recomputing from the inputs — no intermediate
is saved).  Compiled IR trees therefore stay fused under `:>`: `loss::{+/(sigmoid((w1*X)+b1)-Y)^2}` is one
reduce kernel forward and one reduce kernel per scalar parameter backward.  float64 throughout.

Provider hooks mirror `BackendProvider.create_grad_tensor / compute_autograd / compute_multi_autograd /
compute_jacobian` (`base.py:226-258`) with the reference's error types (`autograd.py:10-32`).
"""
import numpy as np
import torch

from . import _lib, autodiff, ops
from .array import DeviceArray

try:
    from klongpy.autograd import AutogradChainBrokenError, NonScalarLossError
except Exception:  # pragma: no cover - GPU box without klongpy
    class AutogradChainBrokenError(Exception):
        def __init__(self, context, expected, actual, suggestion=None):
            msg = f"Autograd chain broken at {context}: expected {expected}, got {actual}."
            if suggestion:
                msg += f" {suggestion}"
            super().__init__(msg)

    class NonScalarLossError(Exception):
        def __init__(self, shape):
            self.shape = shape
            super().__init__(f"Loss function must return a scalar, got shape {shape}. "
                             "Use sum (+/) or mean (%#) to reduce to a scalar.")

NotDifferentiable = autodiff.NotDifferentiable


def wants_grad(args):
    return torch.is_grad_enabled() and any(isinstance(a, DeviceArray) and a._t.requires_grad for a in args)


class _Spec:
    """Static description of one launch: program, mode and where the tensor operands sit among the arguments."""
    __slots__ = ("code", "mode", "red", "template", "out_shape", "device")

    def __init__(self, code, mode, red, template, out_shape, device):
        self.code, self.mode, self.red = tuple(code), mode, red
        self.template, self.out_shape, self.device = template, out_shape, device

    def rebuild(self, tensors):
        """argument list with detached DeviceArrays in the tensor slots"""
        it = iter(tensors)
        return [DeviceArray(next(it).detach()) if slot is _TENSOR else slot for slot in self.template]


_TENSOR = object()


def run_differentiable(code, args, mode, red, code2, red2, out_shape, device):
    """ops.run for operands that require grad: same launch, recorded as one autograd node."""
    if code2 or mode == _lib.MODE_SCANRED:
        raise NotDifferentiable("fused scan+reduce programs take the eager path under autograd")
    if mode == _lib.MODE_REDUCE and red == _lib.RED_MUL:
        raise NotDifferentiable("*/ has no gradient rule here")
    if mode == _lib.MODE_SCAN and red != _lib.RED_ADD:
        raise NotDifferentiable("only +\\ is differentiable")
    autodiff.parse(code)  # unknown operators fail before anything runs
    tensors = [a._t for a in args if isinstance(a, DeviceArray)]
    template = [_TENSOR if isinstance(a, DeviceArray) else a for a in args]
    out = _Launch.apply(_Spec(code, mode, red, template, out_shape, device), *tensors)
    return DeviceArray(out)


def run_differentiable2(code, red, code2, red2, args, device):
    """ops.reduce2 (two folds in one pass) for operands that require grad: ONE autograd node with two outputs; its
    backward is one fused kernel per differentiated operand for BOTH incoming gradients."""
    if red != _lib.RED_ADD or red2 != _lib.RED_ADD:
        raise NotDifferentiable("paired reductions are differentiated for +/ only")
    autodiff.parse(code)
    autodiff.parse(code2)
    tensors = [a._t for a in args if isinstance(a, DeviceArray)]
    template = [_TENSOR if isinstance(a, DeviceArray) else a for a in args]
    spec = _Spec(code, _lib.MODE_REDUCE, red, template, None, device)
    o1, o2 = _Launch2.apply(spec, tuple(code2), *tensors)
    return DeviceArray(o1), DeviceArray(o2)


def _sum_to(x, shape):
    """gradient of a broadcast scalar operand: fold the elementwise gradient"""
    if x.ndim == 0:
        return x._t.reshape(shape)
    size = 1
    for d in shape:
        size *= int(d)
    if size != 1:  # (1, n) against (m, n) and the like: no kernel folds a subset of the axes
        raise autodiff.NotDifferentiable(f"gradient of an operand of shape {tuple(shape)} broadcast to {tuple(x.shape)}")
    return ops.reduce("add", x.flatten())._t.reshape(shape)


_prog_cache = {}


def _grad_prog(code, pos, n_args, gated):
    """Postfix program of  g * d f/d(arg pos)  [* the arg-extreme gate]  over (args..., g[, out, count]); cached —
    the symbolic work happens once per (program, operand), not once per backward pass."""
    key = (code, pos, n_args, gated)
    if key not in _prog_cache:
        tan = autodiff.tangent(autodiff.parse(code), pos)
        if autodiff._is(tan, 0.0):
            _prog_cache[key] = None
        else:
            tree = autodiff.op("mul", ("arg", n_args), tan)
            if gated:
                hit = autodiff.op("eq_b", autodiff.parse(code), ("arg", n_args + 1))
                tree = autodiff.op("mul", tree, autodiff.op("where", hit, autodiff.op("div", autodiff.ONE, ("arg", n_args + 2)),
                                                            autodiff.ZERO))
            _prog_cache[key] = tuple(autodiff.emit(("op", "to_f64", tree)))
    return _prog_cache[key]


def _grad_prog2(code, code2, pos, n_args):
    """g1 * d f1/d(arg pos) + g2 * d f2/d(arg pos) over (args..., g1, g2) as one postfix program (None if both vanish)."""
    key = (code, code2, pos, n_args, "pair")
    if key not in _prog_cache:
        t1 = autodiff.tangent(autodiff.parse(code), pos)
        t2 = autodiff.tangent(autodiff.parse(code2), pos)
        terms = []
        if not autodiff._is(t1, 0.0):
            terms.append(autodiff.op("mul", ("arg", n_args), t1))
        if not autodiff._is(t2, 0.0):
            terms.append(autodiff.op("mul", ("arg", n_args + 1), t2))
        if not terms:
            _prog_cache[key] = None
        else:
            tree = terms[0] if len(terms) == 1 else autodiff.op("add", terms[0], terms[1])
            _prog_cache[key] = tuple(autodiff.emit(("op", "to_f64", tree)))
    return _prog_cache[key]


def _fold_scalars(pending, grads):
    """Gradients of 0-d operands are folds of an elementwise program over the same arguments: two at a time share one
    pass (two-output reduce), e.g. dL/dw and dL/db of a dense layer."""
    k = 0
    while k < len(pending):
        if k + 1 < len(pending) and pending[k][2] is pending[k + 1][2]:
            (ia, pa, full, ta), (ib, pb, _, tb) = pending[k], pending[k + 1]
            try:
                ra, rb = ops.reduce2(pa, _lib.RED_ADD, pb, _lib.RED_ADD, full)
                grads[ia] = ra._t.reshape(ta.shape).to(ta.dtype)
                grads[ib] = rb._t.reshape(tb.shape).to(tb.dtype)
                k += 2
                continue
            except (ValueError, _lib.KgbError):
                pass
        i, prog, full, t = pending[k]
        r = ops.run(prog, full, mode=_lib.MODE_REDUCE, red=_lib.RED_ADD)
        grads[i] = r._t.reshape(t.shape).to(t.dtype)
        k += 1


class _Launch2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec, code2, *tensors):
        with torch.no_grad():
            o1, o2 = ops.reduce2(spec.code, spec.red, code2, _lib.RED_ADD, spec.rebuild(tensors))
        ctx.spec, ctx.code2 = spec, code2
        ctx.save_for_backward(*tensors)
        for o in (o1, o2):
            if not o._t.dtype.is_floating_point:
                ctx.mark_non_differentiable(o._t)
        return o1._t, o2._t

    @staticmethod
    def backward(ctx, g1, g2):
        spec = ctx.spec
        tensors = ctx.saved_tensors
        with torch.no_grad():
            args = spec.rebuild(tensors)
            n_args = len(args)
            dev = tensors[0].device
            z = None

            def scalar(g):
                nonlocal z
                if g is None:
                    if z is None:
                        z = torch.zeros((), dtype=torch.float64, device=dev)
                    return DeviceArray(z)
                return DeviceArray(g.contiguous().to(torch.float64))

            full = args + [scalar(g1), scalar(g2)]
            grads = [None, None] + [None] * len(tensors)
            pending = []
            slots = [k for k, s in enumerate(spec.template) if s is _TENSOR]
            for j, pos in enumerate(slots):
                t = tensors[j]
                if not ctx.needs_input_grad[j + 2] or not t.dtype.is_floating_point:
                    continue
                prog = _grad_prog2(spec.code, ctx.code2, pos, n_args)
                if prog is None:
                    grads[j + 2] = torch.zeros_like(t)
                elif t.dim() > 0:
                    r = ops.run(prog, full)
                    grads[j + 2] = r._t.to(t.dtype) if r._t.shape == t.shape else _sum_to(r, t.shape).to(t.dtype)
                else:
                    pending.append((j + 2, prog, full, t))
            _fold_scalars(pending, grads)
        return tuple(grads)


class _Launch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec, *tensors):
        with torch.no_grad():
            out = ops.run(spec.code, spec.rebuild(tensors), mode=spec.mode, red=spec.red, out_shape=spec.out_shape,
                          device=spec.device)
        ctx.spec = spec
        ctx.save_for_backward(*tensors, out._t)
        if not out._t.dtype.is_floating_point:
            ctx.mark_non_differentiable(out._t)
        return out._t

    @staticmethod
    def backward(ctx, g):
        spec = ctx.spec
        *tensors, out = ctx.saved_tensors
        with torch.no_grad():
            args = spec.rebuild(tensors)
            n_args = len(args)
            G = DeviceArray(g.contiguous())
            extra = [G]
            gate = None  # tree multiplied into the tangent (max/min folds route the gradient to the arg-extremes)
            if spec.mode == _lib.MODE_SCAN:
                # out_j = sum_{k<=j} f_k  =>  dL/df_k = sum_{j>=k} g_j : a reversed running sum
                extra = [ops.reverse(ops.scan("add", ops.reverse(G)))]
            elif spec.mode == _lib.MODE_REDUCE and spec.red in (_lib.RED_MAX, _lib.RED_MIN):
                f = autodiff.parse(spec.code)
                hit = autodiff.op("eq_b", f, ("arg", n_args + 1))
                cnt = ops.run(autodiff.emit(("op", "to_f64", hit)), args + [G, DeviceArray(out)], mode=_lib.MODE_REDUCE,
                              red=_lib.RED_ADD)
                extra = [G, DeviceArray(out), cnt]  # ties share the gradient evenly, like torch.amax
                gate = True
            grads = [None]
            pending = []
            full = args + extra
            slots = [k for k, s in enumerate(spec.template) if s is _TENSOR]
            for j, pos in enumerate(slots):
                t = tensors[j]
                if not ctx.needs_input_grad[j + 1] or not t.dtype.is_floating_point:
                    grads.append(None)
                    continue
                prog = _grad_prog(spec.code, pos, n_args, gate is not None)
                if prog is None:
                    grads.append(torch.zeros_like(t))
                    continue
                elementwise = t.dim() > 0
                if elementwise:
                    r = ops.run(prog, full, out_shape=tuple(t.shape) if t.numel() == 0 else None)
                    grads.append(r._t.to(t.dtype) if r._t.shape == t.shape else _sum_to(r, t.shape).to(t.dtype))
                else:
                    # a 0-d operand broadcast against arrays: its gradient is the FOLD of the elementwise gradient —
                    # one fused reduce kernel, nothing materialised (two such operands share one pass)
                    if any(isinstance(a, DeviceArray) and a.ndim > 0 for a in full):
                        grads.append(None)
                        pending.append((len(grads) - 1, prog, full, t))
                    else:
                        r = ops.run(prog, full, out_shape=())
                        grads.append(r._t.reshape(t.shape).to(t.dtype))
            _fold_scalars(pending, grads)
        return tuple(grads)


# ------------------------------------------------------------------------------------------- provider hooks
def create_grad_tensor(backend, x):
    """A float64 leaf that tracks gradients (the reference casts to float32, torch_backend.py:692-699)."""
    if isinstance(x, DeviceArray):
        # a detached alias, not a copy: device arrays are never modified in place (amend builds a new array), and a
        # clone of a 1e7-element parameter vector costs as much as the fused forward kernel
        t = x._t.detach()
        t = t.to(torch.float64) if t.dtype != torch.float64 else t
    elif isinstance(x, torch.Tensor):
        t = x.detach().clone().to(device=backend._device, dtype=torch.float64)
    else:
        t = torch.as_tensor(np.asarray(x, dtype=np.float64)).to(backend._device)
    return DeviceArray(t.requires_grad_(True))


def _check_loss(y, what):
    if not isinstance(y, DeviceArray):
        kind = "numpy.ndarray" if isinstance(y, np.ndarray) else type(y).__name__
        raise AutogradChainBrokenError(what, "device array", kind, "Keep the computation on b200 device arrays.")
    if y.size != 1:
        raise NonScalarLossError(tuple(y.shape))
    if not y._t.requires_grad:
        raise AutogradChainBrokenError("gradient computation", "requires_grad=True", "requires_grad=False",
                                       "Output lost gradient tracking (integer result, .item(), or a host round trip).")


def compute_autograd(backend, func, x):
    xt = create_grad_tensor(backend, x)
    y = func(xt)
    _check_loss(y, "function output")
    y._t.backward()
    return DeviceArray(xt._t.grad)


def compute_multi_autograd(backend, func, params):
    leaves = [create_grad_tensor(backend, p) for p in params]
    y = func(leaves)
    _check_loss(y, "loss computation")
    grads = torch.autograd.grad(y._t, [p._t for p in leaves], create_graph=False)
    return [DeviceArray(g) for g in grads]


def compute_jacobian(backend, func, x):
    xt = create_grad_tensor(backend, x)
    y = func(xt)
    if not isinstance(y, DeviceArray):
        raise AutogradChainBrokenError("function output", "device array", type(y).__name__)
    flat = y._t.reshape(-1)
    rows = []
    for i in range(flat.numel()):
        (gi,) = torch.autograd.grad(flat[i], xt._t, retain_graph=True, allow_unused=True)
        rows.append(torch.zeros_like(xt._t) if gi is None else gi)
    jac = torch.stack(rows).reshape(tuple(y.shape) + tuple(xt.shape)) if rows else torch.zeros(0, device=xt._t.device)
    return DeviceArray(jac)
