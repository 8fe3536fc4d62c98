"""Host-side logic through the REAL KlongPy interpreter (CPU container only: needs /root/reference).

The provider, np-shim, IR lowering and verb overrides run unchanged; only the device library is replaced by
the NumPy test double (tests/fake_lib.py).  Results must equal the reference NumPy backend's.  The GPU box has
no klongpy, so these tests skip there; the kernels themselves are covered by tests/test_gpu_parity.py.
"""
import sys

import numpy as np
import pytest

from conftest import REFERENCE, have_reference

pytestmark = pytest.mark.skipif(not have_reference(), reason="reference klongpy not available on this box")


@pytest.fixture(scope="module")
def interp():
    sys.path.insert(0, REFERENCE)
    sys.dont_write_bytecode = True
    import fake_lib
    dev = fake_lib.install()
    import klongpy_b200
    from klongpy import KlongInterpreter
    return klongpy_b200.interpreter(device=dev), KlongInterpreter(backend="numpy"), klongpy_b200


def norm(x, mod):
    if isinstance(x, mod.DeviceArray):
        return x.to_numpy()
    if isinstance(x, np.ndarray) and x.dtype == object:
        out = np.empty(len(x), dtype=object)
        for i, y in enumerate(x):
            out[i] = norm(y, mod)
        return out
    return x


PROGRAMS = [
    "1+2", "[1 2 3]+1", "a::[1.0 2.0 3.0];{+/((x*2)+1)^2}(a)", "+\\[1 2 3 4]", "<[3 1 2 1]", ">[3 1 2 1]",
    "=[1 2 1 3 2]", "?[1 1 2 3 3]", "[1 2 3 1]?1", "!5", "&[0 1 0 1]", "[1 2 3]@[0 2]", "|[1 2 3]",
    "[1 2 3],[4 5]", "[1 2 3],4", "2^[1 2 3]", "[1 2 3]^2", "[1.5 2.5]^2", "-[1 2 3]", "_[1.5 2.5]", "%[1 2 4]",
    "#[1 2 3]", "[1 2 3]%2", "[1 2 3]:%2", "[5 6 7]!3", "|/[1 5 2]", "&/[4 2 8]", "|\\[1 5 2 7]",
    "a::!10;+/a*a", "3#[1 2 3 4 5]", "7#[1 2 3]", "2_[1 2 3 4]", "1:+[1 2 3 4]", "[1 2 3]=[1 5 3]",
    "[1 2 3]<[2 2 2]", "[1 2 3]~[1 2 3]", "*[4 5 6]", "^[1 2 3]", "[[1 2] [3 4]]+1", "+[[1 2] [3 4]]",
    "+/[[1 2] [3 4]]", "a::[1 2 3];b::[4 5 6];(a*2)+b", "ts::[3.0 1.0 4.0 1.0 5.0];|/(|\\ts)-ts",
    "{x*2}'[1 2 3]", "10{x+1}:*0", "[1 2 3]|2", "[1 2 3]&2", "+/[]", "#[]", '"abc","def"', "=[1 1 2 2]",
    "[1 [2 3] 4]+1", "(+/[1.0 2.0 3.0])%#[1 2 3]", "p::[1.0 2.0 3.0];s::[4 5 6];(+/p*s)%+/s",
    "a::[1.0 2.0 3.0];a-(+/a)%#a", "-\\[10 1 2 3]", "-/[10 1 2 3]", "*/[1 2 3 4]", "%/[8.0 2.0 2.0]",
    "a::[3 1 2];a@<a", "a::[3 1 2];a@>a", "&5", "&0", "[1 2 3]?4", "(-2)#[1 2 3 4]", "(-1)_[1 2 3 4]",
    "[1.0 2.0]=[1.0 3.0]", "~[1 0 2]", "a::[1 2 3];+/a>1", "x::[0.5 0.25 0.125];+\\(x*2)+1",
]


@pytest.mark.parametrize("src", PROGRAMS)
def test_program_matches_numpy_backend(interp, src):
    kb, kn, mod = interp
    got = norm(kb(src), mod)
    ref = kn(src)
    assert kn._backend.kg_equal(got, ref), (src, got, ref)
    if isinstance(ref, np.ndarray) and ref.dtype != object and isinstance(got, np.ndarray):
        assert got.dtype == ref.dtype, (src, got.dtype, ref.dtype)


def test_arrays_stay_on_the_backend(interp):
    """tests/test_backend_abstraction.py:12-31 pattern: literals and `+` results are backend arrays."""
    kb, _, mod = interp
    assert isinstance(kb("[1 2 3]"), mod.DeviceArray)
    assert isinstance(kb("[1 2 3]+[4 5 6]"), mod.DeviceArray)
    assert isinstance(kb("!4"), mod.DeviceArray)
    assert isinstance(kb("<[3 1 2]"), mod.DeviceArray)


def test_compiled_path_is_taken_and_fused(interp):
    """`compile_expr` hands our provider the IR; the headline tree lowers to ONE reduce stage and
    max-drawdown to ONE scan+reduce stage."""
    kb, _, mod = interp
    from klongpy.compiler import compile_expr
    from klongpy_b200 import _lib
    kb("a::[1.0 2.0 3.0 4.0]")
    _, ast = kb.prog("+/((a*2)+1)^2")
    fn, syms = compile_expr(ast[0], kb)
    assert [s.mode for s in fn.plan.stages] == [_lib.MODE_REDUCE]
    r = fn(kb["a"])
    assert isinstance(r, mod.DeviceArray) and r.ndim == 0 and r.item() == 164.0
    _, ast = kb.prog("|/(|\\a)-a")
    fn, _ = compile_expr(ast[0], kb)
    assert [s.mode for s in fn.plan.stages] == [_lib.MODE_SCANRED]
    kb("p::[1.0 2.0];s::[3 4]")
    _, ast = kb.prog("(+/p*s)%+/s")
    fn, _ = compile_expr(ast[0], kb)
    assert [s.mode for s in fn.plan.stages] == [_lib.MODE_REDUCE, _lib.MODE_REDUCE, _lib.MODE_MAP]


def test_compiled_callable_raises_on_unsupported(interp):
    """interpreter.py:697-701 falls back silently on any exception: host arrays / scalars-only must RAISE."""
    kb, _, mod = interp
    from klongpy_b200 import lowering
    fn = lowering.make_callable(("binop", "+", ("var", "_v0"), ("literal", 1)))
    with pytest.raises(Exception):
        fn(np.arange(3))
    with pytest.raises(Exception):
        fn(3)
    # (N-d operands are compiled since round 2: test_compiled_expressions_take_nd_operands)
    two = lowering.make_callable(("binop", "+", ("var", "_v0"), ("var", "_v1")))
    with pytest.raises(Exception):
        two(kb("[[1 2] [3 4]]"), kb("[[1 2 3] [4 5 6]]"))   # N-d operands of different shapes


def test_host_numpy_symbols_are_not_compiled(interp):
    """tests/test_compiler.py:42-52: a host ndarray symbol makes compile_expr return None on a device backend."""
    kb, _, _ = interp
    from klongpy.compiler import compile_expr
    kb["h"] = np.arange(4.0)
    _, ast = kb.prog("h+1")
    assert compile_expr(ast[0], kb) is None


def test_golden_programs_through_interpreter(interp, golden):
    kb, kn, mod = interp
    arrays = golden["compiled_arrays"]
    for name, v in arrays.items():
        if name.startswith("in_"):
            kb[name[3:]] = kb._backend.kg_asarray(v)
    kb["k"] = 3
    kb["f"] = 2.5
    for case in golden["compiled"]:
        if case["src"] in ("+/e", "*/e"):
            continue  # empty folds: the eager interpreter returns [] (adverbs.py:225), compiled returns 0.0/1.0
        got = norm(kb(case["src"]), mod)
        ref = arrays[case["out"]]
        got = np.asarray(got)
        assert got.shape == ref.shape, case["src"]
        assert got.dtype == ref.dtype, (case["src"], got.dtype, ref.dtype)
        assert np.allclose(got, ref, rtol=1e-12, atol=0, equal_nan=True), case["src"]


def test_golden_verbs_through_interpreter(interp, golden):
    kb, kn, mod = interp
    arrays = golden["verb_arrays"]
    for name, v in arrays.items():
        if name.startswith("in_"):
            kb[name[3:]] = kb._backend.kg_asarray(v)
    for case in golden["verbs"]:
        got = norm(kb(case["src"]), mod)
        key = case["out"]
        if case["ragged"]:
            flat = np.concatenate([np.asarray(x) for x in got])
            starts = np.concatenate([[0], np.cumsum([len(x) for x in got])])
            assert np.array_equal(flat, arrays[key + "_flat"]), case["src"]
            assert np.array_equal(starts, arrays[key + "_starts"]), case["src"]
            continue
        ref = arrays[key]
        got = np.asarray(got)
        assert got.shape == ref.shape, (case["src"], got.shape, ref.shape)
        if not case.get("object_scalars"):
            assert got.dtype == ref.dtype, (case["src"], got.dtype, ref.dtype)
        assert np.allclose(got, ref, rtol=1e-12, atol=0, equal_nan=True), case["src"]


def test_bench_compiler_example_runs(interp):
    """examples/bench_compiler.kg (config 1) expressions on device-resident inputs, vs the NumPy backend."""
    kb, kn, mod = interp
    rng = np.random.default_rng(0)
    n = 1000
    data = {"a": rng.random(n), "b": rng.random(n), "c": rng.random(n), "ts": 100 + rng.standard_normal(n) * 0.5,
            "p": 50 + rng.standard_normal(n) * 10, "s": (np.floor(rng.random(n) * 9900) + 100).astype(np.int64)}
    for k, v in data.items():
        kb[k] = kb._backend.kg_asarray(v)
        kn[k] = v
    exprs = ["a+b", "a*2+b", "(a+b)*(a-b)", "(a*2+b)%(a+1)", "(a*2+b*3)%(a+1)-(b*c)", "{x+y}(a;b)",
             "{(x+y)*(x-y)}(a;b)", "{+/x*y}(a;b)", "+/a", "+/a*b", "|/a-b", "+/a*b+c", "+\\ts", "|\\ts",
             "(|\\ts)-ts", "|/(|\\ts)-ts", "(+/p*s)%+/s", "a-(+/a)%#a", "3{x;a+b;x+1}:*0"]
    for e in exprs:
        got, ref = np.asarray(norm(kb(e), mod)), np.asarray(kn(e))
        assert got.shape == ref.shape and np.allclose(got, ref, rtol=1e-12, atol=0), e


TABLE_PROGRAM = r'''
.py("klongpy_b200.table")
e::[]
e::e,,"a",,[1 2 3]
e::e,,"b",,[2 3 4]
e::e,,"c",,[3 4 5]
T::.table(e)
'''


def table_program_checks(klong, mod):
    """The reference's tests/kgtests/db scenarios (test_create_multi_col_table, test_multi_col_insert_with_single_index,
    test_rindex, test_table_via_dict_behavior) with `.py("klongpy_b200.table")` in place of `.py("klongpy.db")`; the rows
    the reference reads back through `db("select * from T")` are read with `T.rows()` here (no SQL engine)."""
    klong(TABLE_PROGRAM)
    T = klong["T"]
    assert list(klong(".schema(T)")) == ["a", "b", "c"]
    assert T.rows().tolist() == [[1, 2, 3], [2, 3, 4], [3, 4, 5]]
    klong(".insert(T;[4 5 6])")
    klong(".insert(T;[4 5 6])")
    assert T.rows().tolist() == [[1, 2, 3], [2, 3, 4], [3, 4, 5], [4, 5, 6], [4, 5, 6]]
    assert list(klong('.index(T;["a"])')) == ["a"]
    klong(".insert(T;[4 5 6])")
    assert T.rows().tolist() == [[1, 2, 3], [2, 3, 4], [3, 4, 5], [4, 5, 6]]
    klong(".insert(T;[4 6 7])")
    assert T.rows().tolist() == [[1, 2, 3], [2, 3, 4], [3, 4, 5], [4, 6, 7]]
    assert klong(".rindex(T)") == 1 and klong(".rindex(T)") == 0
    klong(".insert(T;[[4 6 7] [9 9 9]])")
    assert T.rows().tolist() == [[1, 2, 3], [2, 3, 4], [3, 4, 5], [4, 6, 7], [4, 6, 7], [9, 9, 9]]
    klong('T,"d",,[1.5 2.5 3.5 4.5 5.5 6.5]')
    assert list(klong(".schema(T)")) == ["a", "b", "c", "d"]
    col = klong('T?"a"')                                    # a column is a device array: the verbs take it as is
    assert isinstance(col, mod.DeviceArray) and col.to_numpy().tolist() == [1, 2, 3, 4, 4, 9]
    assert klong('+/T?"d"') == 24.0
    g = klong('=T?"a"')
    assert [np.asarray(x.to_numpy() if isinstance(x, mod.DeviceArray) else x).tolist() for x in g] == [[0], [1], [2], [3, 4], [5]]
    keys, mn, mean, mx, ct = klong('.gstats(T;"a";"d")')
    assert keys.to_numpy().tolist() == [1, 2, 3, 4, 9] and ct.to_numpy().tolist() == [1, 1, 1, 2, 1]
    assert mn.to_numpy().tolist() == [1.5, 2.5, 3.5, 4.5, 6.5] and mx.to_numpy().tolist() == [1.5, 2.5, 3.5, 5.5, 6.5]
    assert mean.to_numpy().tolist() == [1.5, 2.5, 3.5, 5.0, 6.5]


def test_table_program_through_interpreter(interp):
    klong, _, mod = interp
    table_program_checks(klong, mod)


def test_array_type_has_no_silent_host_fallback(fake_backend):
    """NumPy ufuncs without a kernel RAISE on device arrays instead of copying them to the host (north_star: no CPU
    fallback); `//` and `%` follow NumPy's floored semantics exactly, also beyond 2^53; sum/max/min honour `axis`."""
    be = fake_backend
    a = be.kg_asarray(np.array([-7, 7, 2**62 + 1], dtype=np.int64))
    with pytest.raises(TypeError):
        np.hypot(a, a)
    with pytest.raises(TypeError):
        np.add(a, a, dtype=np.float32)
    h = a.to_numpy()
    assert np.array_equal((a // 1).to_numpy(), h // 1) and (a // 1).to_numpy()[2] == 2**62 + 1
    assert np.array_equal((a // 3).to_numpy(), h // 3) and np.array_equal((a % 3).to_numpy(), h % 3)
    assert np.array_equal((7 % be.kg_asarray(np.array([3, -3, 5]))).to_numpy(), 7 % np.array([3, -3, 5]))
    assert np.array_equal((a % -3).to_numpy(), h % -3)
    f = be.kg_asarray(np.array([-7.5, 7.5, 0.0, -0.0]))
    assert np.array_equal((f % 2.0).to_numpy(), f.to_numpy() % 2.0) and np.array_equal((f // 2.0).to_numpy(), f.to_numpy() // 2.0)
    m = be.kg_asarray(np.arange(6.0).reshape(2, 3))
    assert m.sum().item() == 15.0 and np.array_equal(m.sum(axis=0).to_numpy(), [3.0, 5.0, 7.0])
    assert be.np.sum(m).item() == 15.0 and be.np.max(m).item() == 5.0 and be.np.min(m, axis=0).to_numpy().tolist() == [0.0, 1.0, 2.0]
    with pytest.raises(NotImplementedError):
        m.sum(axis=1)
    from klongpy_b200 import ops
    ints = be.kg_asarray(np.array([1, 2, 3], dtype=np.int64))
    for needle in (float("inf"), float("nan"), 2.5, 2**70):
        assert ops.find_equal(ints, needle).to_numpy().tolist() == []
    assert ops.find_equal(ints, 2.0).to_numpy().tolist() == [1]


def test_sort_by_grade_takes_sorted_keys_from_the_sort(interp, monkeypatch):
    """`a@<a` / `a@>a` (monads.py: "To sort a list a, use a@<a"): the grade is deferred and the `@` on the same array
    asks the sort for its sorted keys — no permutation is built and no gather runs; every other use of `<a`
    materialises the permutation."""
    b200, ref, mod = interp
    from klongpy_b200 import ops
    calls = {"gather": 0}
    real_gather = ops.gather
    monkeypatch.setattr(ops, "gather", lambda *a, **k: (calls.__setitem__("gather", calls["gather"] + 1), real_gather(*a, **k))[1])
    rng = np.random.default_rng(3)
    keys = rng.integers(-50, 50, 2000)
    for kl in (b200, ref):
        kl["a"] = kl._backend.kg_asarray(keys) if kl is b200 else keys
    for src in ("a@<a", "a@>a"):
        got = b200(src)
        assert isinstance(got, mod.DeviceArray)
        assert np.array_equal(got.to_numpy(), np.asarray(ref(src))), src
    assert calls["gather"] == 0
    p = b200("<a")
    assert ops.grade_source(p) is not None            # still deferred
    assert np.array_equal(p.to_numpy(), np.argsort(keys, kind="stable")) and ops.grade_source(p) is None
    assert np.array_equal(b200("(<a)+1").to_numpy(), np.argsort(keys, kind="stable") + 1)
    assert np.array_equal(b200("a@<a+0").to_numpy(), np.sort(keys, kind="stable")) and calls["gather"] == 1  # another array: gather


def test_compiled_expressions_take_nd_operands(interp):
    """N-d operands through the COMPILED path with the reference's own semantics (numpy_backend.py:163-173): elementwise
    trees keep the shape, `+/ */ |/ &/` fold axis 0 (np.<ufunc>.reduce), `+\\` / `*\\` scan the flattened array (np.cumsum)."""
    b200, ref, mod = interp
    rng = np.random.default_rng(12)
    for shape in ((2, 3), (30, 7), (5, 4, 3)):
        m = rng.integers(1, 9, shape).astype(np.float64)
        w = rng.integers(1, 5, shape)
        b200["m"], b200["w"] = b200._backend.kg_asarray(m), b200._backend.kg_asarray(w)
        ref["m"], ref["w"] = m, w
        for src in ("+/m", "*/w", "|/m", "&/m", "+/(m*2)+w", "+\\m", "*\\w", "+\\m*w", "(m*2)+w", "+/m=w", "(+/m)%+/w"):
            got, want = b200(src), np.asarray(ref(src))
            assert isinstance(got, mod.DeviceArray), src
            g = got.to_numpy()
            assert g.shape == want.shape and g.dtype == want.dtype and np.array_equal(g, want), (src, shape)
        # the compiled callable itself (not the interpreter's eager fallback) takes them
        fn, _ = b200._backend.compile_expr_ir(("reduce", "+", ("binop", "*", ("var", "_v0"), ("var", "_v1"))), ["m", "w"])
        assert np.array_equal(fn(b200["m"], b200["w"]).to_numpy(), np.add.reduce(m * w))


def test_find_with_a_list_needle(interp):
    """`a?b` with a list b (dyads.py:340-341): positions of the elements of `a` that equal the whole list."""
    b200, ref, mod = interp
    for src in ("[[1 2] [3 4] [1 2] [5 6]]?[1 2]", "[[1 2] [3 4]]?[9 9]", "[1 2 3 1]?[1 2]", "[[1.5 2.0] [1.5 2.0] [0.0 1.0]]?[1.5 2.0]",
                "[[1 2 3] [4 5 6]]?[4 5 6]", "[[[1 2] [3 4]] [[5 6] [7 8]] [[1 2] [3 4]]]?[[1 2] [3 4]]"):
        got, want = b200(src), np.asarray(ref(src))
        g = got.to_numpy() if isinstance(got, mod.DeviceArray) else np.asarray(got)
        assert list(g) == list(want), src
    rng = np.random.default_rng(2)
    m = rng.integers(0, 3, (5000, 4))
    b200["m"], ref["m"] = b200._backend.kg_asarray(m), m
    b200["b"], ref["b"] = b200._backend.kg_asarray(m[17]), m[17]
    got = b200("m?b")
    assert isinstance(got, mod.DeviceArray)
    assert np.array_equal(got.to_numpy(), np.flatnonzero((m == m[17]).all(axis=1)))


def test_two_folds_and_their_epilogue_are_one_launch(fake_backend):
    """`(+/p*s)%+/s` (compiler.py's VWAP example): both sums share one pass AND the divide runs in the fold kernel's last
    CTA (kgb_expr_compile_epi) — one launch for small vectors; operands the pre-resolved launcher does not take (a float32
    vector: narrow accumulator) still give the right answer through the general stages."""
    import klongpy_b200
    be = fake_backend
    lib = klongpy_b200._lib._host_double
    ir = ("binop", "%", ("reduce", "+", ("binop", "*", ("var", "_v0"), ("var", "_v1"))), ("reduce", "+", ("var", "_v1")))
    fn, _ = be.compile_expr_ir(ir, ["p", "s"])
    rng = np.random.default_rng(2)
    p, s = rng.random(1000), rng.random(1000) + 0.5
    P, S = be.kg_asarray(p), be.kg_asarray(s)
    calls = {"n": 0}
    orig = lib.kgb_expr_launch

    def counting(*a, **k):
        calls["n"] += 1
        return orig(*a, **k)

    lib.kgb_expr_launch = counting
    try:
        r = fn(P, S)
    finally:
        lib.kgb_expr_launch = orig
    assert calls["n"] == 1
    assert abs(r.item() - np.add.reduce(p * s) / np.add.reduce(s)) <= 1e-12 * abs(r.item())
    # integer operands: the epilogue divides in float64 like NumPy
    pi, si = rng.integers(1, 50, 777), rng.integers(1, 9, 777)
    ri = fn(be.kg_asarray(pi), be.kg_asarray(si))
    assert ri.dtype == np.float64 and ri.item() == np.add.reduce(pi * si) / np.add.reduce(si)
    # (+/a) - (+/a)*2 style trees that use a fold twice, and literals in the epilogue
    ir2 = ("binop", "-", ("binop", "*", ("reduce", "+", ("var", "_v0")), ("literal", 2)), ("reduce", "|", ("var", "_v1")))
    fn2, _ = be.compile_expr_ir(ir2, ["a", "b"])
    assert abs(fn2(P, S).item() - (np.add.reduce(p) * 2 - np.max(s))) <= 1e-12 * np.add.reduce(p)
