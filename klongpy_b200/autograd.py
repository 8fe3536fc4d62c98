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


import threading  # noqa: E402

_tls = threading.local()
_parsed_ok = set()


def _parse_checked(code):
    """autodiff.parse as a validity check, once per program"""
    key = tuple(code)
    if key not in _parsed_ok:
        autodiff.parse(code)
        _parsed_ok.add(key)


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
    _parse_checked(code)  # unknown operators fail before anything runs
    tensors = [a._t for a in args if isinstance(a, DeviceArray)]
    template = [_TENSOR if isinstance(a, DeviceArray) else a for a in args]
    spec = _Spec(code, mode, red, template, out_shape, device)
    tape = getattr(_tls, "tape", None)
    if tape is not None:
        # inside compute_autograd / compute_multi_autograd: our own tape — the launch itself plus a list append, instead of a
        # torch.autograd.Function node (~40 us of engine bookkeeping per launch; a fused forward kernel at 1e7 elements
        # takes 25 us).  The result is a fresh leaf that requires grad, so later launches see it as differentiable.
        prev = torch.is_grad_enabled()
        torch._C._set_grad_enabled(False)
        try:
            out = ops.run(spec.code, args, mode=mode, red=red, out_shape=out_shape, device=device, _no_defer=True)
        finally:
            torch._C._set_grad_enabled(prev)
        t = out._t
        if t.dtype.is_floating_point:
            t.requires_grad_(True)
            tape.append((1, spec, tensors, (t,), None))
        return out
    out = _Launch.apply(spec, *tensors)
    return DeviceArray(out)


def run_differentiable2(code, red, code2, red2, args, device):
    """ops.reduce2 (two folds in one pass) for operands that require grad: ONE autograd node with two outputs; its
    backward is one fused kernel per differentiated operand for BOTH incoming gradients."""
    if red != _lib.RED_ADD or red2 != _lib.RED_ADD:
        raise NotDifferentiable("paired reductions are differentiated for +/ only")
    _parse_checked(code)
    _parse_checked(code2)
    tensors = [a._t for a in args if isinstance(a, DeviceArray)]
    template = [_TENSOR if isinstance(a, DeviceArray) else a for a in args]
    spec = _Spec(code, _lib.MODE_REDUCE, red, template, None, device)
    tape = getattr(_tls, "tape", None)
    if tape is not None:
        prev = torch.is_grad_enabled()
        torch._C._set_grad_enabled(False)
        try:
            o1, o2 = ops.reduce2(spec.code, spec.red, code2, _lib.RED_ADD, args)
        finally:
            torch._C._set_grad_enabled(prev)
        outs = []
        for o in (o1, o2):
            if o._t.dtype.is_floating_point:
                o._t.requires_grad_(True)
            outs.append(o._t)
        tape.append((2, spec, tensors, tuple(outs), tuple(code2)))
        return o1, o2
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
        return (None, None) + tuple(_backward2(ctx.spec, ctx.code2, ctx.saved_tensors, g1, g2, ctx.needs_input_grad[2:]))


def _backward2(spec, code2, tensors, g1, g2, needs):
    """Gradients of a two-output fold node w.r.t. its tensor operands: one fused kernel per differentiated operand for
    BOTH incoming gradients (shared by the torch.autograd.Function node and the backend's own tape)."""
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
        grads = [None] * len(tensors)
        pending = []
        slots = [k for k, s in enumerate(spec.template) if s is _TENSOR]
        for j, pos in enumerate(slots):
            t = tensors[j]
            if not needs[j] or not t.dtype.is_floating_point:
                continue
            prog = _grad_prog2(spec.code, code2, pos, n_args)
            if prog is None:
                grads[j] = torch.zeros_like(t)
            elif t.dim() > 0:
                r = ops.run(prog, full)
                grads[j] = r._t.to(t.dtype) if r._t.shape == t.shape else _sum_to(r, t.shape).to(t.dtype)
            else:
                pending.append((j, prog, full, t))
        _fold_scalars(pending, grads)
    return grads


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
        *tensors, out = ctx.saved_tensors
        return (None,) + tuple(_backward1(ctx.spec, tensors, out, g, ctx.needs_input_grad[1:]))


def _backward1(spec, tensors, out, g, needs):
    """Gradients of one launch w.r.t. its tensor operands (shared by the torch.autograd.Function node and the backend's
    own tape): one fused kernel per differentiated array operand, one fused fold per (pair of) 0-d operand(s)."""
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
        grads = []
        pending = []
        full = args + extra
        slots = [k for k, s in enumerate(spec.template) if s is _TENSOR]
        for j, pos in enumerate(slots):
            t = tensors[j]
            if not needs[j] or not t.dtype.is_floating_point:
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
    return grads


def _tape_gradients(tape, y, leaves):
    """Reverse walk of the backend's own tape: d y / d leaf for every leaf tensor, or None where y does not depend on it.

    Tensors that torch arithmetic produced between two launches (views such as flatten / reshape / indexing, the tail of
    dist.sharded_loss_sum) carry a torch grad_fn: their gradient is pushed through torch's graph down to the tape
    results and leaves recorded before them, and the tape takes over again from there."""
    grads = {}

    def add(t, g):
        have = grads.get(id(t))
        grads[id(t)] = g if have is None else have + g

    def through_torch(t, g, upto):
        cand = [o for node in tape[:upto] for o in node[3] if o.requires_grad] + list(leaves)
        for c, gc in zip(cand, torch.autograd.grad(t, cand, grad_outputs=g, allow_unused=True, retain_graph=True)):
            if gc is not None:
                add(c, gc)

    yt = y._t
    if yt.grad_fn is not None:
        through_torch(yt, torch.ones_like(yt), len(tape))
    else:
        grads[id(yt)] = torch.ones_like(yt)
    for k in range(len(tape) - 1, -1, -1):
        kind, spec, tensors, outs, code2 = tape[k]
        gs_out = [grads.pop(id(o), None) for o in outs]
        if all(g is None for g in gs_out):
            continue
        needs = [t.requires_grad for t in tensors]
        if kind == 1:
            gin = _backward1(spec, tensors, outs[0], gs_out[0], needs)
        else:
            gin = _backward2(spec, code2, tensors, gs_out[0], gs_out[1], needs)
        for t, g in zip(tensors, gin):
            if g is None:
                continue
            if t.grad_fn is not None:
                through_torch(t, g, k)
            else:
                add(t, g)
    return [grads.get(id(t)) for t in leaves]


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
    elif isinstance(x, (int, float, np.integer, np.floating)) or (isinstance(x, np.ndarray) and x.ndim == 0):
        # a fill kernel, not a host->device copy: a pageable copy blocks the host until everything queued before it has
        # run, which serialises the CPU with the GPU once per parameter and step
        t = torch.full((), float(x), dtype=torch.float64, device=backend._device)
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


def _with_tape(func, arg):
    """Run the loss function with the backend's own tape recording every differentiable launch."""
    outer = getattr(_tls, "tape", None)
    _tls.tape = tape = []
    try:
        y = func(arg)
        if isinstance(y, DeviceArray):
            y._t  # noqa: B018 - a deferred result is launched HERE, while the tape is still recording
        return y, tape
    finally:
        _tls.tape = outer


def compute_autograd(backend, func, x):
    xt = create_grad_tensor(backend, x)
    y, tape = _with_tape(func, xt)
    _check_loss(y, "function output")
    (g,) = _tape_gradients(tape, y, [xt._t])
    return DeviceArray(g if g is not None else torch.zeros_like(xt._t))


def compute_multi_autograd(backend, func, params):
    leaves = [create_grad_tensor(backend, p) for p in params]
    y, tape = _with_tape(func, leaves)
    _check_loss(y, "loss computation")
    grads = _tape_gradients(tape, y, [p._t for p in leaves])
    if any(g is None for g in grads):
        # torch.autograd.grad's contract (and the reference's behaviour): a parameter the loss does not depend on is an error
        raise RuntimeError("One of the differentiated Tensors appears to not have been used in the graph.")
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
