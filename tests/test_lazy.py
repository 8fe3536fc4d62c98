"""Deferred elementwise expressions (klongpy_b200/lazy.py, opt-in): chains of eager operators — including the `.bkf`
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
    lazy.enable(True)
    try:
        yield fake_backend, ops, launches
    finally:
        lazy.enable(False)
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
