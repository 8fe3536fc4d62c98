"""GPU parity: the CUDA path (through the C ABI) vs the oracle and vs the reference's recorded outputs.

Bars (BASELINE.json north_star): bit-exact for integer, index, grade and group results and for IEEE-exact
elementwise float ops (FMA contraction is off); <= 1e-12 relative for float64 reductions, scans and libm-class
functions (pow/exp/...).  Run with `pytest -m gpu` on a B200.
"""
import numpy as np
import pytest

from conftest import to_ir
from oracle import kg_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-12  # float64 reductions / scans / transcendental functions (north_star tolerance)


@pytest.fixture(scope="module")
def be(cuda_backend):
    return cuda_backend


@pytest.fixture(scope="module")
def K(be):
    from klongpy_b200 import ops
    return ops


def dev(be, a):
    return be.kg_asarray(np.asarray(a))


def host(x):
    from klongpy_b200 import DeviceArray
    return x.to_numpy() if isinstance(x, DeviceArray) else np.asarray(x)


def test_native_library_is_loaded(be):
    """The product path is the CUDA library: it is loaded, bound to an sm_100 device, and JIT kernels run."""
    from klongpy_b200 import _lib
    assert _lib._lib is not None and _lib._host_double is None
    with open("/proc/self/maps") as f:
        assert "libkgb200.so" in f.read()


# ------------------------------------------------------------------------------------------ golden: compiled IR
def test_golden_compiled_ir(be, golden):
    arrays = golden["compiled_arrays"]
    scal = {"k": 3, "f": 2.5}
    exact, tol = 0, 0
    for case in golden["compiled"]:
        ir = to_ir(case["ir"])
        res = be.compile_expr_ir(ir, case["params"])
        assert res is not None, case["src"]
        fn, _ = res
        args = [scal[p] if p in scal else dev(be, arrays["in_" + p]) for p in case["params"]]
        ref = arrays[case["out"]]
        if case["src"] in ("+/e", "*/e", "+\\e"):
            got = host(fn(*args))
            assert got.shape == ref.shape and got.dtype == ref.dtype and np.array_equal(got, ref), case["src"]
            continue
        got = host(fn(*args))
        assert got.shape == ref.shape, case["src"]
        assert got.dtype == ref.dtype, (case["src"], got.dtype, ref.dtype)
        if ref.dtype.kind == "i":
            assert np.array_equal(got, ref), case["src"]
            exact += 1
        elif np.array_equal(got, ref):
            exact += 1
        else:
            # only folds over non-dyadic floats and pow() may differ, and then by <= 1e-12 relative
            assert any(t in case["src"] for t in ("/", "\\", "^")), f"elementwise float result differs: {case['src']}"
            err = np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300))
            assert err <= RTOL, (case["src"], err)
            tol += 1
    assert exact >= 35, (exact, tol)


def test_golden_dyadic_folds_are_bit_exact(be, golden):
    """x = k/256: every partial sum is exactly representable, so any association must match NumPy bit for bit."""
    arrays = golden["compiled_arrays"]
    for case in golden["compiled"]:
        if case["params"] != ["x"]:
            continue
        fn, _ = be.compile_expr_ir(to_ir(case["ir"]), case["params"])
        got = host(fn(dev(be, arrays["in_x"])))
        assert np.array_equal(got, arrays[case["out"]]), case["src"]


# ------------------------------------------------------------------------------------------ golden: verbs
def test_golden_verbs(be, K, golden):
    from klongpy_b200 import verbs
    arrays = golden["verb_arrays"]
    d = {k[3:]: v for k, v in arrays.items() if k.startswith("in_")}
    D = {k: dev(be, v) for k, v in d.items()}
    shim = be.np
    done = 0
    for case in golden["verbs"]:
        src, kind, key = case["src"], case["kind"], case["out"]
        if kind == "grade":
            got = be.argsort(D[src[1:]], descending=(src[0] == ">"))
        elif kind == "group":
            g = verbs.group_indices(D[src[1:]])
            if case["ragged"]:
                flat = np.concatenate([host(x) for x in g])
                starts = np.concatenate([[0], np.cumsum([len(x) for x in g])])
                assert np.array_equal(flat, arrays[key + "_flat"]) and np.array_equal(starts, arrays[key + "_starts"]), src
                done += 1
                continue
            got = g
        elif kind == "range":
            got = K.unique_first(D[src[1:]])
        elif kind == "find":
            name, needle = src.split("?")
            got = K.find_equal(D[name], float(needle) if "." in needle else int(needle))
        elif kind == "enumerate":
            got = K.iota(int(src[1:]))
        elif kind == "expand" and src[1:] in D:
            got = K.expand(D[src[1:]])
        elif kind == "reverse":
            got = K.reverse(D[src[1:]])
        elif kind == "at" and src in ("g1@ix", "f1@ix"):
            got = K.gather(D[src[:2]], D["ix"])
        elif kind == "at" and src == "f1@<f1":
            got = K.gather(D["f1"], be.argsort(D["f1"]))
        elif kind == "over":
            uf = {"+": shim.add, "-": shim.subtract, "*": shim.multiply, "%": shim.divide}.get(src[0])
            got = uf.reduce(D[src[2:]]) if uf else (shim.max if src[0] == "|" else shim.min)(D[src[2:]])
        elif kind == "scan":
            uf = {"+": shim.add, "-": shim.subtract, "*": shim.multiply, "%": shim.divide}.get(src[0])
            got = uf.accumulate(D[src[2:]]) if uf else K.scan("max" if src[0] == "|" else "min", D[src[2:]])
        elif kind == "monad" and src[0] in "-_%":
            a = D[src[1:]]
            got = {"-": lambda: shim.negative(a), "_": lambda: be.floor_to_int(a), "%": lambda: shim.reciprocal(a)}[src[0]]()
        elif kind == "dyad":
            table = {"g2+g1": lambda: shim.add(D["g2"], D["g1"]), "g2*2.5": lambda: shim.multiply(D["g2"], 2.5),
                     "g2|10": lambda: shim.maximum(D["g2"], 10), "g2&10": lambda: shim.minimum(D["g2"], 10),
                     "g1!7": lambda: shim.fmod(D["g1"], 7), "g2<10": lambda: shim.less(D["g2"], 10) * 1,
                     "g2>10": lambda: shim.greater(D["g2"], 10) * 1, "g2=10": lambda: be.safe_equal(D["g2"], 10) * 1,
                     "f1^2": lambda: be.power(D["f1"], 2), "f1^0.5": lambda: be.power(D["f1"], 0.5)}
            if src not in table:
                continue
            got = table[src]()
        elif kind == "take" and src == "5#g1":
            got = D["g1"][:5]
        elif kind == "drop" and src == "5_g1":
            got = D["g1"][5:]
        elif kind == "drop" and src == "(-5)_g1":
            got = D["g1"][:-5]
        else:
            continue
        ref = arrays[key]
        got = host(got)
        assert got.shape == ref.shape, (src, got.shape, ref.shape)
        assert got.dtype == ref.dtype, (src, got.dtype, ref.dtype)
        if ref.dtype.kind == "f":
            assert np.allclose(got, ref, rtol=RTOL, atol=0, equal_nan=True), src
        else:
            assert np.array_equal(got, ref), src
        done += 1
    assert done >= 45, done


def test_reference_suite_vectors_on_gpu(be, K, golden):
    from klongpy_b200 import verbs
    for v in golden["suite_vectors"]:
        src, exp = v["src"], v["expect"]
        if src[0] in "<>=?" and src[1] == "[":
            a = dev(be, np.array([int(t) for t in src[2:-1].split()]))
            if src[0] in "<>":
                got = host(be.argsort(a, descending=src[0] == ">")).tolist()
            elif src[0] == "=":
                g = verbs.group_indices(a)
                got = [host(x).tolist() for x in g] if not hasattr(g, "to_numpy") else host(g).tolist()
            else:
                got = host(K.unique_first(a)).tolist()
        else:
            lst, needle = src.split("]?")
            got = host(K.find_equal(dev(be, np.array([int(t) for t in lst[1:].split()])), int(needle))).tolist()
        assert got == exp, v


# ------------------------------------------------------------------------------------------ elementwise
@pytest.mark.parametrize("n", [1, 2, 3, 31, 1000, 4097, 1_000_003])
def test_elementwise_bit_exact(be, K, n):
    rng = np.random.default_rng(n)
    a, b, c = rng.random(n), rng.random(n) - 0.5, rng.random(n) + 1
    i, j = rng.integers(-1000, 1000, n), rng.integers(1, 50, n)
    A, B, C, I, J = (dev(be, x) for x in (a, b, c, i, j))
    with np.errstate(all="ignore"):
        checks = [
            (A + B, a + b), (A - B, a - b), (A * B, a * b), (A / C, a / c), (A * 2 + B, a * 2 + b),
            ((A * 2 + B * 3) / (A + 1) - (B * C), (a * 2 + b * 3) / (a + 1) - (b * c)),
            (I + J, i + j), (I * J, i * j), (I - 7, i - 7), (I / J, i / j), (I * 2.5, i * 2.5), (A + I, a + i),
            (A ** 2, a ** 2), (I ** 2, i ** 2), (I ** 3, i ** 3), (-A, -a), (-I, -i), (abs(B), abs(b)), (abs(I), abs(i)),
            (be.np.maximum(A, B), np.maximum(a, b)), (be.np.minimum(I, 5), np.minimum(i, 5)),
            (be.np.fmod(I, J), np.fmod(i, j)), (be.np.fmod(I, -7), np.fmod(i, -7)), (be.np.fmod(B, 0.3), np.fmod(b, 0.3)),
            (A < B, a < b), (I > J, i > j), (I == J, i == j), ((A < B) * 1, (a < b) * 1),
            (be.floor_to_int(B * 10), np.floor(b * 10).astype(int)), (be.np.reciprocal(C), np.reciprocal(c)),
            (be.np.sqrt(A), np.sqrt(a)), (be.np.trunc(B * 10), np.trunc(b * 10)), (be.to_int_array(B * 10), (b * 10).astype(int)),
            (A / 0, a / 0), (I / 0, i / 0), (be.np.abs(I), np.abs(i)),
        ]
    for k, (got, ref) in enumerate(checks):
        g = host(got)
        assert g.dtype == ref.dtype, (k, g.dtype, ref.dtype)
        assert np.array_equal(g, ref, equal_nan=True), (k, n)
    for f in ("exp", "log", "sin", "cos", "tanh"):
        g, r = host(getattr(be.np, f)(C)), getattr(np, f)(c)
        assert np.allclose(g, r, rtol=RTOL, atol=1e-300), f
    g, r = host(A ** B), a ** b
    assert np.allclose(g, r, rtol=RTOL, atol=0)


def test_elementwise_views_scalars_and_special_values(be, K):
    a = np.array([0.0, -0.0, 1.5, -2.5, np.inf, -np.inf, np.nan, 1e308, 5e-324])
    A = dev(be, a)
    with np.errstate(all="ignore"):
        for got, ref in [(A + 1, a + 1), (A * 2, a * 2), (A / A, a / a), (be.np.maximum(A, 0.0), np.maximum(a, 0.0)),
                         (be.np.minimum(A, 0.0), np.minimum(a, 0.0)), (A * A, a * a), (A - A, a - a)]:
            assert np.array_equal(host(got), ref, equal_nan=True)
    x = np.arange(10, dtype=np.float64)
    X = dev(be, x)
    assert np.array_equal(host(X[1:] + X[:-1]), x[1:] + x[:-1])          # 8-byte-aligned-only views
    assert np.array_equal(host(X[3:] * 2), x[3:] * 2)
    i = np.array([2**62, -2**62, 2**63 - 1], dtype=np.int64)
    with np.errstate(over="ignore"):
        assert np.array_equal(host(dev(be, i) * 2), i * 2)                  # int64 wraps like numpy
        assert np.array_equal(host(dev(be, i) + dev(be, i)), i + i)
    # 0-d device scalars as operands (results of reductions)
    s = K.reduce("add", X)
    assert np.array_equal(host(X - s / 10), x - x.sum() / 10)
    # empty and length-1
    E = dev(be, np.zeros(0))
    assert host(E + 1).shape == (0,) and host(E * E).shape == (0,)
    assert np.array_equal(host(dev(be, [5.0]) * 3), [15.0])
    # 2-d same-shape elementwise
    m = np.arange(12.0).reshape(3, 4)
    assert np.array_equal(host(dev(be, m) * 2 + 1), m * 2 + 1)
    # whole-result power -> int64 rule is host logic in dyads.py; backend.power keeps float for float input
    assert host(be.power(dev(be, [1.0, 2.0]), 2)).dtype == np.float64


# ------------------------------------------------------------------------------------------ reduce / scan
@pytest.mark.parametrize("n", [1, 2, 255, 256, 4096, 4097, 100_003, 10_000_019])
def test_reduce_parity(be, K, n):
    rng = np.random.default_rng(n + 1)
    x = rng.integers(0, 256, n) / 256.0
    u = rng.random(n) * 100 - 50
    i = rng.integers(-10**6, 10**6, n)
    X, U, I = dev(be, x), dev(be, u), dev(be, i)
    assert K.reduce("add", X).item() == np.add.reduce(x)                       # dyadic: exact
    ir = ("reduce", "+", ("binop", "^", ("binop", "+", ("binop", "*", ("var", "_v0"), ("literal", 2)), ("literal", 1)), ("literal", 2)))
    fn, _ = be.compile_expr_ir(ir, ["x"])
    assert fn(X).item() == O.eval_ir(ir, {"_v0": x})                           # headline expression, exact on dyadic data
    ref = np.add.reduce(u)
    exact = float(np.sum(u.astype(np.longdouble)))
    got = K.reduce("add", U).item()
    assert abs(got - exact) <= RTOL * max(abs(exact), np.sum(np.abs(u)) * 1e-4), (got, ref, exact)
    assert K.reduce("add", I).item() == int(i.sum())
    assert K.reduce("max", U).item() == u.max() and K.reduce("min", U).item() == u.min()
    assert K.reduce("max", I).item() == i.max() and K.reduce("min", I).item() == i.min()
    assert K.reduce("add", X).item() == K.reduce("add", X).item()              # run-to-run deterministic
    small = rng.integers(1, 3, min(n, 40))
    assert K.reduce("mul", dev(be, small)).item() == int(np.multiply.reduce(small))
    cmp_ir = ("reduce", "+", ("cmp", ">", ("var", "_v0"), ("literal", 0.0)))
    fn, _ = be.compile_expr_ir(cmp_ir, ["u"])
    r = fn(U)
    assert host(r).dtype == np.int64 and r.item() == int((u > 0).sum())


def test_reduce_nan_and_overflow(be, K):
    assert np.isnan(K.reduce("max", dev(be, [1.0, np.nan, 3.0])).item())      # |/[1 nan 3] -> nan
    assert np.isnan(K.reduce("min", dev(be, [1.0, np.nan, 3.0])).item())
    big = np.array([2**62, 2**62, 2**62], dtype=np.int64)
    with np.errstate(over="ignore"):
        assert K.reduce("add", dev(be, big)).item() == int(np.add.reduce(big))  # wraps like numpy
    with pytest.raises(ValueError):
        K.reduce("max", dev(be, np.zeros(0)))
    assert K.reduce("add", dev(be, np.zeros(0))).item() == 0.0


@pytest.mark.parametrize("n", [257, 70_001, 6_000_011])
def test_max_min_with_nan_inf_and_signed_zero(be, K, n):
    """np.maximum / np.minimum semantics (NaN from either side propagates and sticks; +-inf; +-0) through every kernel
    family that folds with max/min: elementwise, reduce, scan (generic and pipelined) and scan+reduce."""
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    y = rng.standard_normal(n)
    sp = rng.choice(n, size=min(64, n // 4), replace=False)
    x[sp[0::4]] = np.inf
    x[sp[1::4]] = -np.inf
    x[sp[2::4]] = 0.0
    y[sp[2::4]] = -0.0
    y[sp[3::4]] = np.nan
    X, Y = dev(be, x), dev(be, y)
    for name, f in (("max", np.maximum), ("min", np.minimum)):
        assert np.array_equal(host(K.binary(name, X, Y)), f(x, y), equal_nan=True), name
        assert np.array_equal(host(K.binary(name, Y, X)), f(y, x), equal_nan=True), name
        assert np.array_equal(host(K.scan(name, X)), f.accumulate(x), equal_nan=True), name     # no NaN: +-inf stick
        assert np.array_equal(host(K.scan(name, Y)), f.accumulate(y), equal_nan=True), name     # NaN from its first index on
        assert np.array_equal(host(K.reduce(name, X)), f.reduce(x), equal_nan=True), name
        assert np.isnan(K.reduce(name, Y).item()), name
    # drawdown |/(|\x)-x before the first NaN-free prefix ends, and with a clean vector
    ir = ("reduce", "|", ("binop", "-", ("scan", "|", ("var", "_v0")), ("var", "_v0")))
    fn, _ = be.compile_expr_ir(ir, ["x"])
    z = rng.standard_normal(n)
    assert fn(dev(be, z)).item() == float(np.max(np.maximum.accumulate(z) - z))
    assert np.isnan(fn(Y).item())


@pytest.mark.parametrize("n", [1, 2, 4095, 4096, 4097, 8193, 100_003, 10_000_019])
def test_scan_parity(be, K, n):
    rng = np.random.default_rng(n + 2)
    x = rng.integers(0, 256, n) / 256.0
    u = rng.random(n)
    i = rng.integers(-10**6, 10**6, n)
    X, U, I = dev(be, x), dev(be, u), dev(be, i)
    assert np.array_equal(host(K.scan("add", X)), np.cumsum(x))                # dyadic: bit-exact vs np.cumsum
    assert np.array_equal(host(K.scan("add", I)), np.cumsum(i))
    assert np.array_equal(host(K.scan("max", U)), np.maximum.accumulate(u))
    assert np.array_equal(host(K.scan("min", U)), np.minimum.accumulate(u))
    assert np.array_equal(host(K.scan("max", I)), np.maximum.accumulate(i))
    got = host(K.scan("add", U))
    exact = O.scan_exact(u)
    assert np.max(np.abs(got - exact) / np.abs(exact)) <= RTOL                 # vs long-double oracle
    j = rng.integers(1, 3, n)
    with np.errstate(over="ignore"):
        assert np.array_equal(host(K.scan("mul", dev(be, j))), np.cumprod(j))  # wraps like numpy
    assert np.array_equal(host(K.scan("add", I[1:])), np.cumsum(i[1:]))        # unaligned input view
    ir = ("scan", "+", ("binop", "+", ("binop", "*", ("var", "_v0"), ("literal", 2)), ("literal", 1)))
    fn, _ = be.compile_expr_ir(ir, ["x"])
    assert np.array_equal(host(fn(X)), np.cumsum(x * 2 + 1))                   # fused prologue


def test_scan_reduce_fusion_max_drawdown(be, K):
    for n in (5, 4097, 1_000_003):
        ts = 100 + np.random.default_rng(n).standard_normal(n) * 0.5
        ir = ("reduce", "|", ("binop", "-", ("scan", "|", ("var", "_v0")), ("var", "_v0")))
        fn, _ = be.compile_expr_ir(ir, ["ts"])
        from klongpy_b200 import _lib
        assert [s.mode for s in fn.plan.stages] == [_lib.MODE_SCANRED]
        assert fn(dev(be, ts)).item() == np.max(np.maximum.accumulate(ts) - ts)


def test_nested_stage_dag(be, K):
    """VWAP: two reductions under a scalar divide; results stay 0-d device arrays."""
    rng = np.random.default_rng(9)
    p = 50 + rng.standard_normal(100_000) * 10
    s = (np.floor(rng.random(100_000) * 9900) + 100).astype(np.int64)
    ir = ("binop", "%", ("reduce", "+", ("binop", "*", ("var", "_v0"), ("var", "_v1"))), ("reduce", "+", ("var", "_v1")))
    fn, _ = be.compile_expr_ir(ir, ["p", "s"])
    r = fn(dev(be, p), dev(be, s))
    assert r.ndim == 0
    ref = np.add.reduce(p * s) / np.add.reduce(s)
    assert abs(r.item() - ref) <= RTOL * abs(ref)


# ------------------------------------------------------------------------------------------ grade / group / range / find
def _key_sets(rng, n):
    return {
        "perm": rng.permutation(n),
        "ties": rng.integers(0, max(2, n // 100 + 2), n),
        "full": rng.integers(-2**63, 2**63 - 1, n),
        "const": np.full(n, 7, dtype=np.int64),
        "neg": rng.integers(-1000, 1000, n),
        "f64": rng.standard_normal(n),
        "f64ties": np.round(rng.standard_normal(n), 1),
    }


@pytest.mark.parametrize("n", [1, 2, 33, 4096, 4097, 12289, 1_000_003])
def test_grade_bit_exact_vs_stable_oracle(be, K, n):
    rng = np.random.default_rng(n + 3)
    for name, keys in _key_sets(rng, n).items():
        D = dev(be, keys)
        up = host(be.argsort(D))
        assert up.dtype == np.int64
        assert np.array_equal(up, O.grade_up(keys)), (name, n)
        down = host(be.argsort(D, descending=True))
        assert np.array_equal(down, O.grade_down(keys)), (name, n)


def test_grade_special_floats_and_dtypes(be, K):
    k = np.array([3.0, np.nan, -0.0, 0.0, -np.inf, np.inf, 1.0, np.nan, -1e-300, 1e-300, 0.0, -0.0])
    assert np.array_equal(host(be.argsort(dev(be, k))), np.argsort(k, kind="stable"))
    rng = np.random.default_rng(4)
    k32 = rng.integers(-100, 100, 10_001).astype(np.int32)
    assert np.array_equal(host(K.argsort(K.asarray(k32))), np.argsort(k32, kind="stable"))
    f32 = rng.standard_normal(10_001).astype(np.float32)
    assert np.array_equal(host(K.argsort(K.asarray(f32))), np.argsort(f32, kind="stable"))
    assert host(be.argsort(dev(be, np.zeros(0)))).shape == (0,)


@pytest.mark.parametrize("n,G", [(1, 1), (10, 3), (4097, 100), (100_003, 10_000), (1_000_003, 1000)])
def test_group_bit_exact(be, K, n, G):
    rng = np.random.default_rng(n + G)
    k = rng.integers(0, G, n)
    order, starts, g = K.group(dev(be, k))
    ro, rs = O.group_csr(k)
    assert g == len(rs) - 1
    assert np.array_equal(host(order), ro) and np.array_equal(host(starts), rs)
    f = np.round(rng.standard_normal(n), 1)
    order, starts, g = K.group(dev(be, f))
    ro, rs = O.group_csr(f)
    assert np.array_equal(host(order), ro) and np.array_equal(host(starts), rs)


@pytest.mark.parametrize("n,G", [(1, 1), (10, 3), (4097, 100), (1_000_003, 5000)])
def test_range_and_find_bit_exact(be, K, n, G):
    rng = np.random.default_rng(n * 7 + G)
    a = rng.integers(-G, G, n)
    assert np.array_equal(host(K.unique_first(dev(be, a))), O.range_unique(a))
    f = np.round(rng.standard_normal(n), 1)   # contains -0.0 and 0.0: distinct elements for `?`
    assert np.array_equal(host(K.unique_first(dev(be, f))), O.range_unique(f))
    needle = int(a[n // 2])
    assert np.array_equal(host(K.find_equal(dev(be, a), needle)), O.find(a, needle))
    assert host(K.find_equal(dev(be, a), 10**9)).shape == (0,)
    assert np.array_equal(host(K.find_equal(dev(be, f), float(f[0]))), O.find(f, f[0]))
    flags = rng.integers(0, 3, n)
    assert np.array_equal(host(K.nonzero(dev(be, flags))), np.flatnonzero(flags))
    assert np.array_equal(host(K.nonzero(dev(be, flags) > 1)), np.flatnonzero(flags > 1))


def _first_appearance(a):
    """np.unique on the raw bits + first index: the oracle's `?a` without its Python loop (large inputs)."""
    bits = a.view(np.int64) if a.dtype == np.float64 else a
    _, first = np.unique(bits, return_index=True)
    return a[np.sort(first)]


@pytest.mark.parametrize("case", ["10k", "minus_one", "float_special", "near_limit", "over_limit", "all_distinct", "int32"])
def test_range_hash_path_and_fallback(be, K, case):
    """`?a` at sizes where the L2-resident hash table is tried (n >= 2^21): low cardinality stays on it, the key whose
    bits are all ones (-1) takes its side slot, NaN / -0.0 / +0.0 are distinct raw-bit keys, and more than 2^19 distinct
    keys overflow into the sort-based path — every variant bit-exact against first-appearance order."""
    rng = np.random.default_rng(len(case))
    n = 3_000_001
    if case == "10k":
        a = rng.integers(0, 10_000, n)
    elif case == "minus_one":
        a = rng.integers(-3, 40, n)
    elif case == "float_special":
        a = rng.integers(0, 500, n) / 8.0
        a[::1001] = np.nan
        a[5::997] = -0.0
        a[7::991] = 0.0
    elif case == "near_limit":
        a = rng.integers(0, 250_000, n)
    elif case == "over_limit":
        a = rng.integers(0, 900_000, n)
    elif case == "all_distinct":
        a = rng.permutation(n)
    else:
        a = rng.integers(-5, 3000, n).astype(np.int32)
    got = host(K.unique_first(dev(be, a)))
    want = _first_appearance(a)
    assert got.dtype == want.dtype and got.shape == want.shape
    gb = got.view(np.int64) if got.dtype == np.float64 else got
    wb = want.view(np.int64) if want.dtype == np.float64 else want
    assert np.array_equal(gb, wb), case


@pytest.mark.parametrize("c", [1, 2, 255, 256, 257, 4097, 16_383, 16_384, 16_385, 40_000])
@pytest.mark.parametrize("n", [50_000, 2_500_003])
def test_range_distinct_count_boundaries(be, K, c, n):
    """The number of distinct keys around the edges of the first-position ranking kernel (tiles of 256, at most 16 384
    keys; above that the radix sort orders the first positions), on the sort path (small n) and the hash path."""
    if c > n:
        pytest.skip("more distinct keys than rows")
    rng = np.random.default_rng(c + n)
    vals = rng.permutation(10 * c)[:c] - 3 * c               # distinct, scattered, negative and positive
    a = vals[rng.integers(0, c, n)]
    a[rng.permutation(n)[:c]] = vals                         # every key occurs
    got = host(K.unique_first(dev(be, a)))
    assert np.array_equal(got, _first_appearance(a))
    f = a / 4.0                                              # the same as float64 keys (information in the high word)
    assert np.array_equal(host(K.unique_first(dev(be, f))), _first_appearance(f))


def test_range_cache_adversaries(be, K):
    """Keys built to fight the per-CTA key cache: all of them folding onto ONE cache entry (multiples of 2^28 xor-fold
    to the same index), small keys next to the patterns 'empty' entries are initialised with (c ^ 1), the all-ones key
    as the majority key, and a key that first appears in the very last row."""
    rng = np.random.default_rng(77)
    n = 2_600_000
    same_entry = (rng.integers(0, 300, n) << 28) | 5
    assert np.array_equal(host(K.unique_first(dev(be, same_entry))), _first_appearance(same_entry))
    small = rng.integers(0, 16_384, n) ^ 1
    assert np.array_equal(host(K.unique_first(dev(be, small))), _first_appearance(small))
    ones = np.where(rng.random(n) < 0.9, -1, rng.integers(0, 50, n))
    ones[-1] = 123_456_789
    assert np.array_equal(host(K.unique_first(dev(be, ones))), _first_appearance(ones))
    late = np.zeros(n, dtype=np.int64)
    late[-1] = -1
    assert np.array_equal(host(K.unique_first(dev(be, late))), np.array([0, -1]))


def test_support_verbs(be, K):
    rng = np.random.default_rng(5)
    assert np.array_equal(host(K.iota(0)), O.enumerate_(0)) and np.array_equal(host(K.iota(100_001)), O.enumerate_(100_001))
    a = rng.random(100_003)
    idx = rng.integers(-100_003, 100_003, 50_001)
    assert np.array_equal(host(K.gather(dev(be, a), dev(be, idx))), a[idx])
    with pytest.raises(IndexError):
        K.gather(dev(be, a), dev(be, [0, 100_003]))
    assert np.array_equal(host(K.reverse(dev(be, a))), O.reverse(a))
    assert np.array_equal(host(dev(be, a)[::-1]), a[::-1]) and np.array_equal(host(dev(be, a)[10:2:-3]), a[10:2:-3])
    c = rng.integers(0, 4, 10_001)
    assert np.array_equal(host(K.expand(dev(be, c))), O.expand_where(c))
    assert be.array_equal(dev(be, a), dev(be, a.copy()))
    b = a.copy()
    b[77_777] += 1
    assert not be.array_equal(dev(be, a), dev(be, b))
    assert not be.array_equal(dev(be, [np.nan]), dev(be, [np.nan]))            # np.array_equal: NaN != NaN
    m = rng.random((7, 5))
    assert np.array_equal(host(K.gather(dev(be, m), dev(be, [3, 0, 6, 3]))), m[[3, 0, 6, 3]])
    assert np.array_equal(host(be.np.concatenate((dev(be, [1, 2]), dev(be, [3.5])))), np.array([1, 2, 3.5]))
    assert np.array_equal(host(be.np.tile(dev(be, [1, 2, 3]), 3)), np.tile([1, 2, 3], 3))


# ------------------------------------------------------------------------------------------ grouped aggregation
@pytest.mark.parametrize("n,G", [(10, 3), (100_003, 100), (3_000_001, 10_000)])
def test_grouped_min_mean_max(be, K, n, G):
    rng = np.random.default_rng(n + G + 6)
    k = rng.integers(0, G, n)
    v = np.clip(np.round(rng.normal(10, 15, n), 1), -99.9, 99.9)
    mn, mx, sm, ct = (host(t) for t in K.groupagg(dev(be, k), dev(be, v), G))
    rmn, rmx, rsm, rct = O.grouped_stats(k, v, G)
    assert np.array_equal(mn, rmn) and np.array_equal(mx, rmx) and np.array_equal(ct, rct)   # bit-exact / exact
    assert np.allclose(sm, rsm, rtol=RTOL, atol=1e-9)
    # the 1brc report lines (par.kg:4-6).  The mean is sum%count printed with one decimal; a sum that differs in its
    # last bits (association order) may flip the print only when the mean sits on an x.x5 rounding boundary.
    rep, ref = O.report_1brc(mn, mx, sm, ct), O.report_1brc(rmn, rmx, rsm, rct)
    assert rep.keys() == ref.keys()
    for s in rep:
        if rep[s] != ref[s]:
            m10 = rsm[s] / rct[s] * 10
            assert abs(abs(m10 - np.floor(m10)) - 0.5) < 1e-9, (s, rep[s], ref[s])
            assert rep[s].split("/")[0::2] == ref[s].split("/")[0::2]
    with pytest.raises(Exception):
        K.groupagg(dev(be, [0, G]), dev(be, [1.0, 2.0]), G)


@pytest.mark.parametrize("case", ["narrow", "wide", "rising", "falling", "G12800", "G12801", "G40000"])
def test_grouped_extremes_filter_stress(be, K, case):
    """The per-CTA filters hold only the top 16 bits of each running extreme: values crowded into one filter bucket,
    exponents from denormal to 1e300 with infinities and signed zeros, strictly rising / falling streams (every row is
    a new extreme) and station counts at and beyond what fits the shared-memory tables — min/max bit-exact, counts
    exact, sums in tolerance."""
    rng = np.random.default_rng(len(case))
    n, G = 2_000_003, 1000
    if case == "narrow":
        v = 64.0 + rng.integers(0, 1 << 20, n) * 2.0 ** -40
    elif case == "wide":
        v = rng.standard_normal(n) * 10.0 ** rng.integers(-300, 300, n)
        v[::1013] = 0.0
        v[1::1013] = -0.0
        v[2::5003] = np.inf
        v[3::5003] = -np.inf
        v[4::7001] = 5e-324
    elif case == "rising":
        v = np.arange(n, dtype=np.float64) * 0.25 - 1000.0
    elif case == "falling":
        v = -np.arange(n, dtype=np.float64) * 0.25 + 1000.0
    else:
        G = int(case[1:])
        v = np.round(rng.normal(10, 15, n), 1)
    k = rng.integers(0, G, n)
    mn, mx, sm, ct = (host(t) for t in K.groupagg(dev(be, k), dev(be, v), G))
    rmn, rmx, rsm, rct = O.grouped_stats_fast(k, v, G)
    assert np.array_equal(ct, rct)
    assert np.array_equal(mn.view(np.int64), rmn.view(np.int64)) or np.array_equal(mn, rmn)
    assert np.array_equal(mx.view(np.int64), rmx.view(np.int64)) or np.array_equal(mx, rmx)
    if case != "wide":
        assert np.allclose(sm, rsm, rtol=RTOL, atol=1e-9)
    else:  # +inf and -inf in one station sum to NaN in both; compare where finite
        fin = np.isfinite(rsm)
        assert np.array_equal(np.isnan(sm), np.isnan(rsm)) and np.allclose(sm[fin], rsm[fin], rtol=1e-9, atol=0)


def test_grouped_agg_merge_and_nan(be, K):
    k = np.array([0, 1, 0, 2, 1])
    v = np.array([1.0, np.nan, -3.0, 4.0, 2.0])
    mn, mx, sm, ct = (host(t) for t in K.groupagg(dev(be, k), dev(be, v), 4))
    assert mn[0] == -3.0 and mx[0] == 1.0 and np.isnan(mn[1]) and np.isnan(mx[1]) and ct.tolist() == [2, 2, 1, 0]
    assert mn[3] == np.inf and mx[3] == -np.inf and sm[3] == 0.0
    # merge step (worker.kg `merge`): accumulate a second chunk into the same tables
    t = K.groupagg(dev(be, k[:3]), dev(be, np.array([1.0, 5.0, -3.0])), 4)
    t = K.groupagg(dev(be, k[3:]), dev(be, np.array([4.0, 2.0])), 4, tables=t)
    mn, mx, sm, ct = (host(x) for x in t)
    assert mn.tolist()[:3] == [-3.0, 2.0, 4.0] and mx.tolist()[:3] == [1.0, 5.0, 4.0] and ct.tolist() == [2, 2, 1, 0]


# ------------------------------------------------------------------------------------------ full-size properties
def test_full_size_properties_1e9(be, K):
    """BASELINE.json config 2 size (1e9 float64): size-independent properties instead of a CPU oracle."""
    import torch
    n = 1_000_000_000
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs 40 GB of free HBM")
    from klongpy_b200 import DeviceArray
    g = torch.Generator(device="cuda").manual_seed(7)
    ki = torch.randint(0, 256, (n,), dtype=torch.int64, device="cuda", generator=g)
    KI = DeviceArray(ki)
    X = KI / 256                                            # dyadic float64 data
    isum = K.reduce("add", KI).item()                       # exact integer sum
    assert K.reduce("add", X).item() == isum / 256.0        # float fold is exact on dyadic data
    ir = ("reduce", "+", ("binop", "^", ("binop", "+", ("binop", "*", ("var", "_v0"), ("literal", 2)), ("literal", 1)), ("literal", 2)))
    fn, _ = be.compile_expr_ir(ir, ["x"])
    sq = K.reduce("add", (KI * 2 + 256) * (KI * 2 + 256)).item()   # ((x*2)+1)^2 * 65536, exact in int64
    assert fn(X).item() == sq / 65536.0
    del ki, KI
    S = K.scan("add", X)
    assert DeviceArray(S.tensor[-1]).item() == isum / 256.0                         # last == total
    # adjacent differences reproduce the input exactly (dyadic): max |(S[i]-S[i-1]) - x[i]| == 0
    d = DeviceArray(S.tensor[1:]) - DeviceArray(S.tensor[:-1]) - DeviceArray(X.tensor[1:])
    assert K.reduce("max", abs(d)).item() == 0.0
    assert DeviceArray(S.tensor[0]).item() == DeviceArray(X.tensor[0]).item()


def test_full_size_grade_group_1e8(be, K):
    """config 3 size (1e8 int64 keys): permutation validity, sortedness, stability, group boundaries."""
    import torch
    from klongpy_b200 import DeviceArray
    n = 100_000_000
    g = torch.Generator(device="cuda").manual_seed(8)
    keys = DeviceArray(torch.randint(0, 10_000, (n,), dtype=torch.int64, device="cuda", generator=g))
    perm = be.argsort(keys)
    s = K.gather(keys, perm)
    assert K.reduce("min", DeviceArray(s.tensor[1:]) - DeviceArray(s.tensor[:-1])).item() >= 0      # sorted
    # stable: inside runs of equal keys positions ascend  <=>  (key, pos) pairs strictly increase
    dk = DeviceArray(s.tensor[1:]) - DeviceArray(s.tensor[:-1])
    dp = DeviceArray(perm.tensor[1:]) - DeviceArray(perm.tensor[:-1])
    bad = K.reduce("add", ((dk == 0) * 1) * ((dp <= 0) * 1)).item()
    assert bad == 0
    assert K.reduce("add", perm).item() == n * (n - 1) // 2                                          # a permutation
    assert K.array_equal(K.gather(perm, K.argsort(perm)), K.iota(n))
    down = be.argsort(keys, descending=True)
    assert K.array_equal(down, K.reverse(perm))                                                      # grade-down contract
    order, starts, G = K.group(keys)
    assert G == 10_000 and K.array_equal(order, perm)
    heads = K.gather(s, DeviceArray(starts.tensor[:-1]))
    assert K.array_equal(heads, K.iota(10_000))                                                      # ascending key order
    sizes = DeviceArray(starts.tensor[1:]) - DeviceArray(starts.tensor[:-1])
    assert K.reduce("add", sizes).item() == n and K.reduce("min", sizes).item() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 1000, 100_003, 3_000_001])
def test_two_output_reduce_and_vwap(be, K, n):
    """`(+/p*s)%+/s` (examples/bench_compiler.kg VWAP): both sums from ONE pass over p and s, through the compiled IR
    and through ops.reduce2 directly; exact on dyadic prices and integer sizes."""
    rng = np.random.default_rng(n + 77)
    p = rng.integers(0, 4096, n) / 64.0
    s = rng.integers(100, 10_000, n)
    P, S = dev(be, p), dev(be, s)
    a, b = K.reduce2([K.OP_ARG, 0, K.OP_ARG, 1, K.OPC["mul"]], "+", [K.OP_ARG, 1], "+", [P, S])
    assert a.item() == np.add.reduce(p * s) and b.item() == np.add.reduce(s)
    assert host(a).dtype == np.float64 and host(b).dtype == np.int64
    ir = ("binop", "%", ("reduce", "+", ("binop", "*", ("var", "_v0"), ("var", "_v1"))), ("reduce", "+", ("var", "_v1")))
    fn, _ = be.compile_expr_ir(ir, ["p", "s"])
    assert any(st.pair is not None for st in fn.plan.stages)  # the two reductions were paired
    assert fn(P, S).item() == O.eval_ir(ir, {"_v0": p, "_v1": s})
    mx, mn = K.reduce2([K.OP_ARG, 0], "|", [K.OP_ARG, 0], "&", [P])
    assert mx.item() == p.max() and mn.item() == p.min()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2_500_000, 20_000_001])
def test_pipelined_scan_with_carried_seed(be, K, n):
    """The persistent pipelined scan (>= 4 tiles per SM) with the exclusive offset a shard / chunk carries
    (dist.sharded_scan, stream.HostStream): bit-exact on dyadic data, repeated to give a race a chance to show."""
    import torch
    from klongpy_b200 import _lib
    rng = np.random.default_rng(n)
    k = rng.integers(0, 256, n)
    x = dev(be, k / 256.0)
    from klongpy_b200 import DeviceArray
    seed = DeviceArray(torch.tensor(12345.5, dtype=torch.float64, device="cuda:0"))  # 0-d device scalar
    ref = np.cumsum(k) / 256.0 + 12345.5
    for _ in range(5):
        s = K.run([K.OP_ARG, 0], [x, seed], mode=_lib.MODE_SCAN, red=_lib.RED_ADD, code2=[K.OP_ARG, 1])
        assert np.array_equal(host(s), ref)
    i = dev(be, k)
    for _ in range(3):
        assert np.array_equal(host(K.scan("add", i)), np.cumsum(k))
        assert np.array_equal(host(K.scan("max", x)), np.maximum.accumulate(k / 256.0))


@pytest.mark.gpu
def test_grade_large_float_keys_with_nan_inf_and_signed_zero(be, K):
    """The persistent pipelined sort pass on float64 keys with every special value mixed in: NaN sorts last, -0.0 and
    +0.0 tie (stable by position), +-inf at the ends — np.argsort(kind='stable') is the contract (writer.py:97-108)."""
    rng = np.random.default_rng(99)
    n = 3_000_017
    k = np.round(rng.standard_normal(n), 2)
    special = rng.integers(0, n, 60_000)
    k[special[:10_000]] = np.nan
    k[special[10_000:20_000]] = np.inf
    k[special[20_000:30_000]] = -np.inf
    k[special[30_000:45_000]] = -0.0
    k[special[45_000:]] = 0.0
    d = dev(be, k)
    ref = np.argsort(k, kind="stable")
    assert np.array_equal(host(K.argsort(d)), ref)
    assert np.array_equal(host(K.argsort(d, descending=True)), ref[::-1])
    order, starts, g = K.group(d)
    assert np.array_equal(host(order), ref)
    u = host(K.unique_first(dev(be, k[:200_000])))
    kk = k[:200_000]
    # first-appearance order; NaNs collapse to one entry, -0.0 and 0.0 stay distinct (the reference keys on str(x))
    seen, want = set(), []
    for x in kk:
        key = "nan" if x != x else repr(float(x))
        if key not in seen:
            seen.add(key)
            want.append(x)
    want = np.array(want)
    assert len(u) == len(want) and np.array_equal(np.isnan(u), np.isnan(want))
    assert np.array_equal(u[~np.isnan(u)], want[~np.isnan(want)]) and np.array_equal(np.signbit(u), np.signbit(want))


# ------------------------------------------------------------------------------------------ packed radix pass
def _packed_key_sets(rng, n):
    """Key sets aimed at the packed sort's planner (csrc/kgb_sort_packed.cuh): key ranges that cross zero, trailing
    constant bits, keys too wide for the packed word (prefix mode) with short and with long prefix ties."""
    sets = dict(_key_sets(rng, n))
    sets["cross0"] = rng.integers(-5, 6, n)
    sets["stride1024"] = rng.integers(0, max(2, n), n) * 1024 + 7
    sets["two"] = rng.integers(0, 2, n) * (2**62) - 2**61
    hi = rng.integers(0, 3, n).astype(np.int64) << 45
    sets["prefix_long_ties"] = hi | rng.integers(0, 2**20, n)          # few prefixes, many low parts: fix-up overflows
    sets["prefix_short_ties"] = (rng.integers(-2**62, 2**62, n) >> 9) << 9 | rng.integers(0, 4, n)
    f = rng.standard_normal(n)
    f[rng.integers(0, n, max(1, n // 50))] = np.nan
    f[rng.integers(0, n, max(1, n // 50))] = -0.0
    f[rng.integers(0, n, max(1, n // 50))] = 0.0
    f[rng.integers(0, n, max(1, n // 100))] = np.inf
    sets["f64special"] = f
    sets["f64temps"] = np.round(rng.normal(10, 15, n), 1)               # wide live range, few distinct values
    sets["i32"] = rng.integers(-1000, 1000, n).astype(np.int32)
    sets["f32"] = rng.standard_normal(n).astype(np.float32)
    return sets


@pytest.mark.parametrize("n", [2, 100, 8191, 8192, 8193, 24577, 300_001, 1_000_003])
def test_grade_packed_path_bit_exact(be, K, n, monkeypatch):
    """Every size through the packed pass (threshold forced down), every planner branch, both directions, sorted keys."""
    monkeypatch.setenv("KGB200_SORT_PACKED_MIN", "2")
    rng = np.random.default_rng(n + 11)
    for name, keys in _packed_key_sets(rng, n).items():
        D = K.asarray(keys)
        ref = np.argsort(keys, kind="stable")
        up, ks = K.argsort(D, return_sorted=True)
        assert np.array_equal(host(up), ref), (name, n)
        assert np.array_equal(host(ks), keys[ref], equal_nan=True), (name, n)
        assert np.array_equal(host(K.argsort(D, descending=True)), ref[::-1]), (name, n)
        if keys.dtype == np.int64 or name in ("f64ties", "f64temps"):
            order, starts, g = K.group(D)
            ro, rs = O.group_csr(keys)
            assert np.array_equal(host(order), ro) and np.array_equal(host(starts), rs), (name, n)


@pytest.mark.parametrize("n", [300_001, 1_000_003])
def test_grade_general_pass_still_bit_exact(be, K, n, monkeypatch):
    """The 12-byte general pass is the fallback of the packed one: keep it covered at sizes the packed path now takes."""
    monkeypatch.setenv("KGB200_SORT_PACKED", "0")
    rng = np.random.default_rng(n + 12)
    for name, keys in _key_sets(rng, n).items():
        D = dev(be, keys)
        assert np.array_equal(host(be.argsort(D)), O.grade_up(keys)), (name, n)
        assert np.array_equal(host(be.argsort(D, descending=True)), O.grade_down(keys)), (name, n)


# ------------------------------------------------------------------------------------------ integer //, %, float32 folds
def test_floor_divide_and_remainder_match_numpy(be, K):
    """`//` and `%` of the array type: exact integer kernels (also beyond 2^53, x//0 == x%0 == 0 like NumPy's loops) and
    NumPy's divmod for floats, bit for bit (signed zeros, infinities and NaN included)."""
    rng = np.random.default_rng(5)
    a = np.concatenate([rng.integers(-2**62, 2**62, 50_000), [2**62 + 1, -2**62 - 1, -2**63, 2**63 - 1, 0, -7, 7]])
    b = np.concatenate([rng.integers(-1000, 1000, 50_000), [1, 1, -1, -1, 5, 3, -3]])
    A, B = dev(be, a), dev(be, b)
    with np.errstate(all="ignore"):
        fd, md = np.floor_divide(a, b), np.remainder(a, b)
    assert np.array_equal(host(A // B), fd) and host(A // B).dtype == np.int64
    assert np.array_equal(host(A % B), md) and host(A % B).dtype == np.int64
    assert np.array_equal(host(A // 7), a // 7) and np.array_equal(host(A % -7), a % -7)
    with np.errstate(all="ignore"):
        assert np.array_equal(host(10**18 // B), np.floor_divide(10**18, b))
    x = np.concatenate([rng.standard_normal(50_000) * 1e3, [0.0, -0.0, 7.5, -7.5, np.inf, -np.inf, np.nan, 1e308, 5.0]])
    y = np.concatenate([rng.standard_normal(50_000), [3.0, 3.0, 2.5, 2.5, 2.0, 2.0, 2.0, 1e-308, 0.0]])
    X, Y = dev(be, x), dev(be, y)
    with np.errstate(all="ignore"):
        fd, md = np.floor_divide(x, y), np.remainder(x, y)
    assert np.array_equal(host(X // Y), fd, equal_nan=True) and np.array_equal(np.signbit(host(X // Y)), np.signbit(fd))
    assert np.array_equal(host(X % Y), md, equal_nan=True) and np.array_equal(np.signbit(host(X % Y)), np.signbit(md))
    with pytest.raises(TypeError):
        np.hypot(X, Y)  # no kernel: raises, never a silent host copy


@pytest.mark.parametrize("n", [1000, 1_000_003, 20_000_001])
def test_float32_reduce_and_scan_within_1e6(be, K, n):
    """north_star tolerance for float32 folds: 1e-6 relative against a float64 evaluation of the same float32 data."""
    rng = np.random.default_rng(n + 32)
    x = rng.random(n, dtype=np.float32)
    X = K.asarray(x)
    assert X.dtype == np.float32
    ref = np.add.reduce(x.astype(np.float64))
    s = K.reduce("add", X)
    assert host(s).dtype == np.float32
    assert abs(float(s.item()) - ref) <= 1e-6 * abs(ref)
    assert float(K.reduce("max", X).item()) == float(x.max()) and float(K.reduce("min", X).item()) == float(x.min())
    c = host(K.scan("add", X))
    assert c.dtype == np.float32
    cref = np.cumsum(x.astype(np.float64))
    assert np.max(np.abs(c - cref) / cref) <= 1e-6
    assert np.array_equal(host(K.scan("max", X)), np.maximum.accumulate(x))
    # a fused float32 tree: +/((x*2)+1)^2 keeps float32 with a python-scalar (weak) literal, like NumPy 2
    ir = ("reduce", "+", ("binop", "^", ("binop", "+", ("binop", "*", ("var", "_v0"), ("literal", 2)), ("literal", 1)), ("literal", 2)))
    fn, _ = be.compile_expr_ir(ir, ["x"])
    got = fn(X)
    want = np.add.reduce(((x.astype(np.float64) * 2) + 1) ** 2)
    assert host(got).dtype == np.float32 and abs(float(got.item()) - want) <= 1e-6 * want
    # (a running product rounds once per element: 200 factors keep the float32 chain itself inside 1e-6)
    p = host(K.scan("mul", K.asarray((1 + x[:200] * 1e-3).astype(np.float32))))
    pref = np.cumprod((1 + x[:200] * 1e-3).astype(np.float32).astype(np.float64))
    assert np.max(np.abs(p - pref) / pref) <= 1e-6


def test_range_hash_path_collapses_nans_like_the_sort_path(be, K):
    """`?a` must not depend on which path ran: NaNs of either sign / any payload are ONE value on the hash path (n >= 2M,
    low cardinality) as on the sort path and in the reference (a set of str(x)); -0.0 and +0.0 stay distinct."""
    rng = np.random.default_rng(17)
    n = 3_000_000
    x = np.round(rng.standard_normal(n), 1)
    nan_neg = np.frombuffer(np.uint64(0xfff8000000000000).tobytes(), dtype=np.float64)[0]
    nan_pay = np.frombuffer(np.uint64(0x7ff8000000000123).tobytes(), dtype=np.float64)[0]
    x[rng.integers(0, n, 5000)] = np.nan
    x[rng.integers(0, n, 5000)] = nan_neg
    x[rng.integers(0, n, 5000)] = nan_pay
    x[rng.integers(0, n, 5000)] = -0.0
    u = host(K.unique_first(dev(be, x)))
    small = host(K.unique_first(dev(be, x[:100_000])))  # sort path
    assert np.isnan(u).sum() == 1 and np.isnan(small).sum() == 1
    want = O.range_unique(x)
    assert len(u) == len(want)
    assert np.array_equal(u[~np.isnan(u)], want[~np.isnan(want)]) and np.array_equal(np.signbit(u[~np.isnan(u)]), np.signbit(want[~np.isnan(want)]))


@pytest.mark.parametrize("case", ["dense", "negative", "stride", "int32", "outlier", "edge_2p15", "too_wide"])
def test_range_direct_address_path(be, K, case):
    """`?a` on integer keys spanning <= 2^15 values: first positions from a direct-address table in shared memory, the key
    range guessed from a sample and verified — an outlier the sample missed must fall back, never mis-answer."""
    rng = np.random.default_rng(hash(case) % 1000)
    n = 1_500_007
    if case == "dense":
        k = rng.integers(0, 10_000, n)
    elif case == "negative":
        k = rng.integers(-3000, 3000, n)
    elif case == "stride":
        k = rng.integers(0, 500, n) * 4096 + 17
    elif case == "int32":
        k = rng.integers(-100, 30_000, n).astype(np.int32)
    elif case == "outlier":
        k = rng.integers(0, 1000, n)
        k[n // 3] = 10**15          # one key far outside what a 64 K sample sees
        k[n // 2] = -5
    elif case == "edge_2p15":
        k = rng.integers(0, 2**15, n)
    else:
        k = rng.integers(0, 2**20, n)
    got = host(K.unique_first(K.asarray(k)))
    want = O.range_unique(k)
    assert got.dtype == k.dtype and np.array_equal(got, want), case


@pytest.mark.gpu
def test_two_folds_with_epilogue_single_kernel(cuda_backend):
    """`(+/p*s)%+/s` and relatives through kgb_expr_compile_epi (both folds + the epilogue in ONE kernel) against NumPy:
    float64 to 1e-12, integer folds exact, min / max folds bit-exact, at a size with many CTAs and at a tiny one."""
    be = cuda_backend
    rng = np.random.default_rng(31)
    vwap = ("binop", "%", ("reduce", "+", ("binop", "*", ("var", "_v0"), ("var", "_v1"))), ("reduce", "+", ("var", "_v1")))
    spread = ("binop", "-", ("reduce", "|", ("var", "_v0")), ("reduce", "&", ("var", "_v1")))
    scaled = ("binop", "+", ("binop", "*", ("reduce", "+", ("var", "_v0")), ("literal", 3)), ("reduce", "+", ("var", "_v1")))
    f_vwap, _ = be.compile_expr_ir(vwap, ["p", "s"])
    f_spread, _ = be.compile_expr_ir(spread, ["a", "b"])
    f_scaled, _ = be.compile_expr_ir(scaled, ["a", "b"])
    for n in (7, 100_000, 1_500_001):
        p, s = rng.random(n) * 100, rng.random(n) + 0.5
        P, S = be.kg_asarray(p), be.kg_asarray(s)
        ref = np.add.reduce(p * s) / np.add.reduce(s)
        got = f_vwap(P, S)
        assert got.dtype == np.float64 and abs(got.item() - ref) <= 1e-12 * abs(ref), n
        assert f_spread(P, S).item() == np.max(p) - np.min(s), n
        pi, si = rng.integers(-1000, 1000, n), rng.integers(1, 9, n)
        PI, SI = be.kg_asarray(pi), be.kg_asarray(si)
        assert f_scaled(PI, SI).item() == int(np.add.reduce(pi)) * 3 + int(np.add.reduce(si)), n
        gi = f_vwap(PI, SI)
        assert gi.dtype == np.float64 and gi.item() == np.add.reduce(pi * si) / np.add.reduce(si), n
