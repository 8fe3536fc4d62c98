"""Deferred elementwise expressions (klongpy_b200/lazy.py, on by default): chains of eager operators — including the `.bkf`
unary math and `& |` the reference's compiler IR cannot express — fuse into ONE launch, reductions take a deferred
operand as their prologue, a second consumer materialises, and autograd still matches closed forms.  CPU test double."""
import numpy as np
import pytest


@pytest.fixture()
def lazy_env(fake_backend):
    import fake_lib
    from klongpy_b200 import lazy, ops
    lib = fake_lib.install() and __import__("klongpy_b200")._lib._host_double
    launches = {"n": 0}
    orig = lib.kgb_expr_launch

    def counting(*a, **k):
        launches["n"] += 1
        return orig(*a, **k)

    lib.kgb_expr_launch = counting
    was, was_min = lazy.ENABLED, lazy.MIN_SIZE
    lazy.enable(True)
    lazy.MIN_SIZE = 0   # (the size threshold below which operators launch at once is a tuning knob, not semantics)
    try:
        yield fake_backend, ops, launches
    finally:
        lazy.enable(was)
        lazy.MIN_SIZE = was_min
        lib.kgb_expr_launch = orig


def test_chain_fuses_into_one_launch(lazy_env):
    be, ops, launches = lazy_env
    x = np.linspace(-2, 2, 1001)
    X = be.kg_asarray(x)
    launches["n"] = 0
    s = 1 / (1 + ops.unary("exp", 0 - ((X * 0.5) + 0.25)))  # sigmoid((w*X)+b): 5 operators
    assert s._tensor is None and launches["n"] == 0           # nothing has run yet
    assert s.shape == (1001,) and s.dtype == np.float64      # metadata without a launch (kgb_expr_infer)
    got = s.to_numpy()
    assert launches["n"] == 1                                 # ONE fused kernel
    assert np.allclose(got, 1 / (1 + np.exp(-(x * 0.5 + 0.25))), rtol=1e-15)
    assert s._tensor is not None and s.to_numpy() is not None and launches["n"] == 1  # materialised once


def test_reduce_takes_a_deferred_prologue(lazy_env):
    be, ops, launches = lazy_env
    x, y = np.arange(1.0, 101.0), np.arange(100.0, 0.0, -1.0)
    X, Y = be.kg_asarray(x), be.kg_asarray(y)
    launches["n"] = 0
    r = ops.reduce("add", ops.binary("max", X, Y) * 2 - 1)   # `|` is outside the reference's compiler IR
    assert launches["n"] == 1
    assert r.item() == np.add.reduce(np.maximum(x, y) * 2 - 1)
    launches["n"] = 0
    sc = ops.scan("add", (X * Y) + 1)
    assert launches["n"] == 1 and np.array_equal(sc.to_numpy(), np.cumsum(x * y + 1))


def test_second_consumer_materialises(lazy_env):
    be, ops, launches = lazy_env
    X = be.kg_asarray(np.arange(10.0))
    launches["n"] = 0
    t = ops.unary("exp", X * 0.1)       # deferred
    a = t + 1                           # first consumer: inlines t
    b = t * 2                           # second consumer: t is materialised first, then b defers over the leaf
    assert t._tensor is not None
    ra, rb = a.to_numpy(), b.to_numpy()
    assert launches["n"] == 3           # t, a (recomputing t inline), b
    assert np.allclose(ra, np.exp(np.arange(10.0) * 0.1) + 1) and np.allclose(rb, np.exp(np.arange(10.0) * 0.1) * 2)


def test_dtype_inference_and_shape_errors(lazy_env):
    be, ops, _ = lazy_env
    I, F = be.kg_asarray(np.arange(5)), be.kg_asarray(np.arange(5.0))
    assert (I * 3 - 1).dtype == np.int64 and (I / 2).dtype == np.float64 and (I < 3).dtype == np.bool_
    assert ((I * 2) + F).dtype == np.float64
    with pytest.raises(ValueError):
        (I * 2) + be.kg_asarray(np.arange(6))   # raised at the operator, not at materialisation


def test_autograd_through_a_deferred_chain(lazy_env):
    be, ops, launches = lazy_env
    rng = np.random.default_rng(5)
    X, Y = rng.random(2000) * 2, rng.random(2000)
    dX, dY = be.kg_asarray(X), be.kg_asarray(Y)

    def loss(params):
        w, b = params
        s = 1 / (1 + ops.unary("exp", 0 - ((w * dX) + b)))
        return ops.reduce("add", (s - dY) ** 2)

    launches["n"] = 0
    gw, gb = be.compute_multi_autograd(loss, [0.1, 0.1])
    s = 1 / (1 + np.exp(-(0.1 * X + 0.1)))
    d = 2 * (s - Y) * s * (1 - s)
    assert np.allclose([gw.item(), gb.item()], [np.sum(d * X), np.sum(d)], rtol=1e-10)
    assert launches["n"] <= 4   # fused forward reduce + one fused backward reduce per parameter (was ~30 launches)


# ------------------------------------------------------------------------------------------ on the GPU
@pytest.mark.gpu
def test_lazy_is_the_default_and_matches_eager_bit_for_bit(cuda_backend):
    """Deferred evaluation is the default path on the device: chains fuse into one kernel, and every result equals the
    operator-by-operator (eager) evaluation bit for bit (-fmad=false: fusing never changes a rounding) and the oracle."""
    from klongpy_b200 import lazy, ops
    from oracle import kg_oracle as O
    assert lazy.ENABLED
    be = cuda_backend
    rng = np.random.default_rng(21)
    n = 3_000_003   # above lazy.MIN_SIZE: elementwise results are deferred
    x, y = rng.standard_normal(n), rng.random(n) + 0.5
    i = rng.integers(-50, 50, n)
    X, Y, I = be.kg_asarray(x), be.kg_asarray(y), be.kg_asarray(i)

    def programs(ops_, X, Y, I):
        return {
            "sigmoid": 1 / (1 + ops_.unary("exp", 0 - ((X * 0.5) + 0.25))),
            "maxmin": ops_.binary("min", ops_.binary("max", X, Y) * 2 - 1, Y + X),
            "ints": ((I * 3) - 7) // 4 % 5,
            "mixed": ((I * 2) + X) / Y,
            "cmp": ((X < Y) * 1) + ((I == 3) * 1),
            "reduce": ops_.reduce("add", ops_.unary("abs", X) * Y + 1),
            "reduce_max": ops_.reduce("max", (X * X) - Y),
            "scan": ops_.scan("add", ops_.unary("floor", Y * 256) / 256),
            "scanmax": ops_.scan("max", X - Y),
        }

    lazy_out = programs(ops, X, Y, I)
    deferred = [k for k, v in lazy_out.items() if v._tensor is None]
    assert set(deferred) >= {"sigmoid", "maxmin", "ints", "mixed", "cmp"}      # elementwise results are still descriptions
    lazy_np = {k: v.to_numpy() for k, v in lazy_out.items()}
    lazy.enable(False)
    try:
        eager_np = {k: v.to_numpy() for k, v in programs(ops, X, Y, I).items()}
    finally:
        lazy.enable(True)
    for k in lazy_np:
        assert lazy_np[k].dtype == eager_np[k].dtype and np.array_equal(lazy_np[k], eager_np[k], equal_nan=True), k
    # and against NumPy (the reference's arithmetic)
    assert np.array_equal(lazy_np["maxmin"], np.minimum(np.maximum(x, y) * 2 - 1, y + x))
    assert np.array_equal(lazy_np["ints"], ((i * 3) - 7) // 4 % 5)
    assert np.array_equal(lazy_np["mixed"], ((i * 2) + x) / y)
    assert np.array_equal(lazy_np["cmp"], ((x < y) * 1) + ((i == 3) * 1))
    assert np.allclose(lazy_np["sigmoid"], 1 / (1 + np.exp(-(x * 0.5 + 0.25))), rtol=1e-12)
    assert abs(lazy_np["reduce"] - np.add.reduce(np.abs(x) * y + 1)) <= 1e-12 * np.add.reduce(np.abs(x) * y + 1)
    assert lazy_np["reduce_max"] == np.max(x * x - y)
    assert np.array_equal(lazy_np["scan"], np.cumsum(np.floor(y * 256) / 256))   # dyadic: exact in any association
    assert np.array_equal(lazy_np["scanmax"], np.maximum.accumulate(x - y))
    ir = ("reduce", "+", ("binop", "^", ("binop", "+", ("binop", "*", ("var", "_v0"), ("literal", 2)), ("literal", 1)), ("literal", 2)))
    fn, _ = be.compile_expr_ir(ir, ["x"])
    d = rng.integers(0, 256, n) / 256.0
    assert fn(be.kg_asarray(d)).item() == O.eval_ir(ir, {"_v0": d})


@pytest.mark.gpu
def test_lazy_autograd_on_the_device_matches_torch_float64(cuda_backend):
    """config 5 (sigmoid NN) with deferred operators: 4 launches, gradients equal torch float64 autograd of the same
    expression to 1e-10; the same with lazy switched off."""
    import torch
    from klongpy_b200 import lazy, ops
    be = cuda_backend
    n = 3_000_003
    g = torch.Generator(device="cuda").manual_seed(3)
    Xt = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2
    Yt = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    from klongpy_b200 import DeviceArray
    X, Y = DeviceArray(Xt), DeviceArray(Yt)

    def loss(params):
        w, b = params
        s = 1 / (1 + ops.unary("exp", 0 - ((w * X) + b)))
        return ops.reduce("add", (s - Y) ** 2)

    w = torch.tensor(0.1, dtype=torch.float64, device="cuda", requires_grad=True)
    b = torch.tensor(0.1, dtype=torch.float64, device="cuda", requires_grad=True)
    ref = torch.autograd.grad(((torch.sigmoid(w * Xt + b) - Yt) ** 2).sum(), [w, b])
    for flag in (True, False):
        lazy.enable(flag)
        try:
            gw, gb = be.compute_multi_autograd(loss, [0.1, 0.1])
        finally:
            lazy.enable(True)
        assert abs(gw.item() - ref[0].item()) <= 1e-10 * abs(ref[0].item()), flag
        assert abs(gb.item() - ref[1].item()) <= 1e-10 * abs(ref[1].item()), flag
    gx = be.compute_autograd(lambda v: ops.reduce("add", ops.unary("tanh", v * Y) * v), X)
    xr = Xt.clone().requires_grad_(True)
    (rx,) = torch.autograd.grad((torch.tanh(xr * Yt) * xr).sum(), [xr])
    assert torch.allclose(gx.tensor, rx, rtol=1e-10, atol=1e-12)
