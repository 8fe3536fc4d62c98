"""Autograd interop: `:>` through the REAL interpreter on the CPU test double, the symbolic
differentiator on its own, and — on the GPU — the fused forward/backward kernels against torch float64 autograd of
the same expression (the reference's torch path is float32, torch_backend.py:692-699; closed forms are the
reference's own test style, tests/test_autograd_parametrized.py)."""
import sys

import numpy as np
import pytest

from conftest import REFERENCE, have_reference


# ------------------------------------------------------------------------------------------- symbolic layer (CPU)
def _eval_tree(t, env):
    import math
    if t[0] == "arg":
        return env[t[1]]
    if t[0] == "lit":
        return float(t[1])
    k = [_eval_tree(c, env) for c in t[2:]]
    n = t[1]
    f = {"add": lambda a, b: a + b, "sub": lambda a, b: a - b, "mul": lambda a, b: a * b, "div": lambda a, b: a / b,
         "pow": lambda a, b: a ** b, "neg": lambda a: -a, "exp": math.exp, "log": math.log, "sin": math.sin, "cos": math.cos,
         "tanh": math.tanh, "sqrt": math.sqrt, "recip": lambda a: 1 / a, "square": lambda a: a * a, "to_f64": float,
         "tan": math.tan, "abs": abs, "sign": lambda a: (a > 0) - (a < 0), "gt_b": lambda a, b: a > b, "lt_b": lambda a, b: a < b,
         "eq_b": lambda a, b: a == b, "where": lambda c, a, b: a if c else b, "max": max, "min": min}[n]
    return f(*k)


@pytest.mark.parametrize("name,build,x", [
    ("poly", lambda K, L: [K.OP_ARG, 0] + L(2) + [K.OPC["mul"]] + L(1) + [K.OPC["add"]] + L(2) + [K.OPC["pow"]], 0.7),
    ("sigmoid", lambda K, L: L(1) + L(1) + L(0) + [K.OP_ARG, 0, K.OPC["sub"], K.OPC["exp"], K.OPC["add"], K.OPC["div"]], 0.3),
    ("quot", lambda K, L: [K.OP_ARG, 0, K.OPC["sin"], K.OP_ARG, 0, K.OPC["square"]] + L(1) + [K.OPC["add"], K.OPC["div"]], 1.1),
    ("sqrt-tanh", lambda K, L: [K.OP_ARG, 0, K.OPC["tanh"], K.OPC["abs"], K.OPC["sqrt"]], 0.9),
    ("pow-half", lambda K, L: [K.OP_ARG, 0] + L(0.5) + [K.OPC["pow"]], 2.3),
    ("max", lambda K, L: [K.OP_ARG, 0] + L(0.5) + [K.OPC["max"], K.OP_ARG, 0, K.OPC["mul"]], 0.8),
])
def test_tangent_matches_finite_difference(name, build, x):
    from klongpy_b200 import autodiff, ops
    code = build(ops, ops.lit)
    tree = autodiff.parse(code)
    assert autodiff.emit(tree) == list(code)  # parse / emit round trip
    tan = autodiff.tangent(tree, 0)
    h = 1e-6
    fd = (_eval_tree(tree, {0: x + h}) - _eval_tree(tree, {0: x - h})) / (2 * h)
    assert abs(_eval_tree(tan, {0: x}) - fd) < 1e-6 * max(1.0, abs(fd)), name


def test_tangent_constant_folds():
    from klongpy_b200 import autodiff, ops
    code = [ops.OP_ARG, 0] + ops.lit(3) + [ops.OPC["add"], ops.OP_ARG, 1, ops.OPC["mul"]]  # (a+3)*b
    tree = autodiff.parse(code)
    assert autodiff.tangent(tree, 0) == ("arg", 1)
    assert autodiff.tangent(tree, 1) == ("op", "add", ("arg", 0), ("lit", 3))
    assert autodiff.grad_program([ops.OP_ARG, 0, ops.OP_ARG, 0, ops.OPC["gt"]], 0, 1) is None  # piecewise constant


# ------------------------------------------------------------------------------------------- through the interpreter (CPU)
needs_ref = pytest.mark.skipif(not have_reference(), reason="reference klongpy not available on this box")


@pytest.fixture(scope="module")
def interp():
    sys.path.insert(0, REFERENCE)
    sys.dont_write_bytecode = True
    import fake_lib
    dev = fake_lib.install()
    import klongpy_b200
    return klongpy_b200.interpreter(device=dev), klongpy_b200


def _np(x, mod):
    return x.to_numpy() if isinstance(x, mod.DeviceArray) else np.asarray(x)


@needs_ref
@pytest.mark.parametrize("src,expect", [
    ("{x^2}:>3.0", 6.0),
    ("{x^3}:>2.0", 12.0),
    ("{+/x^2}:>[1.0 2.0 3.0]", [2.0, 4.0, 6.0]),
    ("{+/x*x}:>[1.0 2.0 3.0]", [2.0, 4.0, 6.0]),
    ("{(+/x*x)%3}:>[1.0 2.0 3.0]", [2 / 3, 4 / 3, 2.0]),
    ("{+/(x*2)+1}:>[1.0 2.0]", [2.0, 2.0]),
    ("{|/x}:>[1.0 5.0 2.0]", [0.0, 1.0, 0.0]),
    ("{+/+\\x}:>[1.0 2.0 3.0]", [3.0, 2.0, 1.0]),
    ("{(+/x)^0.5}:>[1.0 3.0]", [0.25, 0.25]),
])
def test_grad_through_interpreter(interp, src, expect):
    klong, mod = interp
    got = _np(klong(src), mod)
    assert np.allclose(got, np.asarray(expect), rtol=1e-12, atol=1e-12), (src, got)
    assert got.dtype == np.float64  # float64 end to end (the reference's torch path demotes to float32)


@needs_ref
def test_multi_param_sigmoid_loss(interp):
    """BASELINE config 5(a): loss::{+/(sigmoid((w1*X)+b1)-Y)^2}; loss:>[w1 b1] — vs the closed form."""
    klong, mod = interp
    rng = np.random.default_rng(3)
    X, Y = rng.random(1000) * 2, rng.random(1000)
    be = klong._backend
    klong["X"], klong["Y"] = be.kg_asarray(X), be.kg_asarray(Y)
    klong(".bkf(\"exp\")")
    klong("w1::0.1;b1::0.1")
    klong("sigmoid::{1%(1+exp(0-x))}")
    klong("loss::{+/(sigmoid((w1*X)+b1)-Y)^2}")
    g = klong("loss:>[w1 b1]")
    s = 1 / (1 + np.exp(-(0.1 * X + 0.1)))
    d = 2 * (s - Y) * s * (1 - s)
    ref = np.array([np.sum(d * X), np.sum(d)])
    got = np.array([float(_np(x, mod)) for x in g])
    assert np.allclose(got, ref, rtol=1e-10), (got, ref)


@needs_ref
def test_sharpe_gradient(interp):
    """BASELINE config 5(b): sharpe::{(+/x*returns)%((+/((x^2)*(vols^2)))^0.5)}; sharpe:>w."""
    klong, mod = interp
    rng = np.random.default_rng(4)
    n = 500
    r, v = rng.uniform(0.01, 0.15, n), rng.uniform(0.05, 0.4, n)
    w = np.full(n, 1.0 / n)
    be = klong._backend
    klong["returns"], klong["vols"], klong["w"] = be.kg_asarray(r), be.kg_asarray(v), be.kg_asarray(w)
    klong("sharpe::{(+/x*returns)%((+/((x^2)*(vols^2)))^0.5)}")
    got = _np(klong("sharpe:>w"), mod)
    A, B = np.sum(w * r), np.sum(w * w * v * v)
    ref = r / np.sqrt(B) - A * w * v * v / B ** 1.5
    assert np.allclose(got, ref, rtol=1e-10), np.max(np.abs(got - ref))


@needs_ref
def test_non_scalar_loss_raises(interp):
    klong, mod = interp
    from klongpy.autograd import NonScalarLossError
    with pytest.raises(NonScalarLossError):
        klong("{x*2}:>[1.0 2.0]")


def test_paired_reductions_stay_paired_under_autograd(monkeypatch):
    """Compiled trees with two +/ over the same operands (Sharpe, VWAP) run as ONE two-output reduce forward and ONE
    fused gradient kernel backward; two scalar parameters of one node share one two-output fold.  Checked against
    closed forms, with the launches counted."""
    import fake_lib
    dev = fake_lib.install()
    import klongpy_b200 as kb
    from klongpy_b200 import ops
    be = kb.B200BackendProvider(device=dev)
    rng = np.random.default_rng(8)
    n = 4099
    r, v, w = rng.uniform(0.01, 0.15, n), rng.uniform(0.05, 0.4, n), np.full(n, 1.0 / n)
    sharpe_ir = ("binop", "%", ("reduce", "+", ("binop", "*", ("var", "_v0"), ("var", "_v1"))),
                 ("binop", "^", ("reduce", "+", ("binop", "*", ("binop", "^", ("var", "_v0"), ("literal", 2)),
                                                 ("binop", "^", ("var", "_v2"), ("literal", 2)))), ("literal", 0.5)))
    sharpe, _ = be.compile_expr_ir(sharpe_ir, ["x", "returns", "vols"])
    R, V, W = be.kg_asarray(r), be.kg_asarray(v), be.kg_asarray(w)
    calls = {"run": 0, "reduce2": 0}
    real_run, real_r2 = ops.run, ops.reduce2

    def run(*a, **k):
        calls["run"] += 1
        return real_run(*a, **k)

    def reduce2(*a, **k):
        calls["reduce2"] += 1
        return real_r2(*a, **k)

    monkeypatch.setattr(ops, "run", run)
    monkeypatch.setattr(ops, "reduce2", reduce2)
    got = be.compute_autograd(lambda x: sharpe(x, R, V), W).to_numpy()
    A, B = np.sum(w * r), np.sum(w * w * v * v)
    assert np.allclose(got, r / np.sqrt(B) - A * w * v * v / B ** 1.5, rtol=1e-10)
    assert calls["reduce2"] >= 1                       # forward: one two-output kernel, not two reductions
    big = calls["run"]
    # dense layer: two scalar parameters inside ONE compiled tree -> their folds share a pass
    X, Y = rng.random(n), rng.random(n)
    ir = ("reduce", "+", ("binop", "^", ("binop", "-", ("binop", "+", ("binop", "*", ("var", "_v0"), ("var", "_v1")), ("var", "_v2")),
                                         ("var", "_v3")), ("literal", 2)))   # variables are numbered in first-reference order
    fn, _ = be.compile_expr_ir(ir, ["w", "X", "b", "Y"])
    DX, DY = be.kg_asarray(X), be.kg_asarray(Y)
    calls["reduce2"] = 0
    gw, gb = be.compute_multi_autograd(lambda p: fn(p[0], DX, p[1], DY), [0.3, -0.2])
    d = 2 * (0.3 * X - 0.2 - Y)
    assert np.allclose([float(gw.to_numpy()), float(gb.to_numpy())], [np.sum(d * X), np.sum(d)], rtol=1e-10)
    assert calls["reduce2"] == 1, calls                # dL/dw and dL/db: one pass over X, Y
    assert big <= 8, calls


# ------------------------------------------------------------------------------------------- GPU: fused kernels vs torch f64
@pytest.mark.gpu
def test_gpu_sigmoid_and_sharpe_vs_torch(cuda_backend):
    import torch
    from klongpy_b200 import DeviceArray, ops
    be = cuda_backend
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(11)
    n = 1_000_003
    X = torch.rand(n, dtype=torch.float64, device=dev, generator=g) * 2
    Y = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    # (a) loss = +/(sigmoid(w*X+b)-Y)^2 as the interpreter would evaluate it, operator by operator
    def loss(params):
        w, b = params
        z = (w * DeviceArray(X)) + b
        s = 1 / (1 + ops.unary("exp", 0 - z))
        return ops.reduce("add", (s - DeviceArray(Y)) ** 2)
    gw, gb = be.compute_multi_autograd(loss, [0.1, 0.1])
    wt = torch.tensor(0.1, dtype=torch.float64, device=dev, requires_grad=True)
    bt = torch.tensor(0.1, dtype=torch.float64, device=dev, requires_grad=True)
    ref = torch.autograd.grad(((torch.sigmoid(wt * X + bt) - Y) ** 2).sum(), [wt, bt])
    assert abs(gw.item() - ref[0].item()) <= 1e-10 * abs(ref[0].item())
    assert abs(gb.item() - ref[1].item()) <= 1e-10 * abs(ref[1].item())
    # the same loss as ONE compiled IR tree: fused forward, fused backward
    ir = ("reduce", "+", ("binop", "^", ("binop", "-", ("binop", "*", ("var", "_v0"), ("var", "_v1")), ("var", "_v2")),
                          ("literal", 2)))
    fn, _ = be.compile_expr_ir(ir, ["w", "X", "Y"])
    gw2, = be.compute_multi_autograd(lambda p: fn(p[0], DeviceArray(X), DeviceArray(Y)), [0.3])
    wt2 = torch.tensor(0.3, dtype=torch.float64, device=dev, requires_grad=True)
    ref2, = torch.autograd.grad(((wt2 * X - Y) ** 2).sum(), [wt2])
    assert abs(gw2.item() - ref2.item()) <= 1e-10 * abs(ref2.item())
    # (b) Sharpe gradient w.r.t. a vector
    r = torch.rand(n, dtype=torch.float64, device=dev, generator=g) * 0.14 + 0.01
    v = torch.rand(n, dtype=torch.float64, device=dev, generator=g) * 0.35 + 0.05
    w0 = torch.full((n,), 1.0 / n, dtype=torch.float64, device=dev)
    def sharpe(x):
        return ops.reduce("add", x * DeviceArray(r)) / (ops.reduce("add", (x ** 2) * (DeviceArray(v) ** 2)) ** 0.5)
    got = be.compute_autograd(sharpe, DeviceArray(w0))
    wt3 = w0.clone().requires_grad_(True)
    ref3, = torch.autograd.grad((wt3 * r).sum() / ((wt3 ** 2) * (v ** 2)).sum() ** 0.5, [wt3])
    err = ((got.tensor - ref3).abs().max() / ref3.abs().max()).item()  # norm-wise: single elements pass through zero
    assert err <= 1e-10, err
    # running-sum backward = reversed running sum
    got = be.compute_autograd(lambda x: ops.reduce("add", ops.scan("add", x) * DeviceArray(r)), DeviceArray(w0))
    wt4 = w0.clone().requires_grad_(True)
    ref4, = torch.autograd.grad((torch.cumsum(wt4, 0) * r).sum(), [wt4])
    assert ((got.tensor - ref4).abs().max() / ref4.abs().max()).item() <= 1e-10
