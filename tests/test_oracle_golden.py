"""The oracle (oracle/kg_oracle.py) is pinned here against (1) outputs of the REAL reference recorded by
tests/golden/make_golden.py and (2) the reference's own golden vectors lifted from its test-suite.  CPU only."""
import numpy as np
import pytest

from conftest import to_ir
from oracle import kg_oracle as O


SCALARS = {"k": 3, "f": 2.5}


def make_env(case, arrays):
    env = {}
    for i, p in enumerate(case["params"]):
        env[f"_v{i}"] = SCALARS[p] if p in SCALARS else arrays["in_" + p]
    return env


def test_compiled_ir_matches_reference(golden):
    arrays = golden["compiled_arrays"]
    n_exact = 0
    for case in golden["compiled"]:
        ir = to_ir(case["ir"])
        ref = arrays[case["out"]]
        if case["src"] in ("*/e",) or case["how"].startswith("interpreter (") :
            continue
        got = np.asarray(O.eval_ir(ir, make_env(case, arrays)))
        assert got.dtype == ref.dtype, case["src"]
        assert got.shape == ref.shape, case["src"]
        assert np.array_equal(got, ref, equal_nan=True), case["src"]
        n_exact += 1
    assert n_exact >= 50


def _verb_inputs(golden):
    return {k[3:]: v for k, v in golden["verb_arrays"].items() if k.startswith("in_")}


def test_verbs_match_reference(golden):
    arrays = golden["verb_arrays"]
    d = _verb_inputs(golden)
    checked = 0
    for case in golden["verbs"]:
        src, kind, key = case["src"], case["kind"], case["out"]
        if kind == "grade":
            a = d[src[1:]]
            got = O.grade_up(a) if src[0] == "<" else O.grade_down(a)
        elif kind == "group":
            order, starts = O.group_csr(d[src[1:]])
            if case["ragged"]:
                assert np.array_equal(order, arrays[key + "_flat"]), src
                assert np.array_equal(starts, arrays[key + "_starts"]), src
            else:  # equal-sized groups: the reference returns a 2-d int64 array
                ref = arrays[key]
                assert np.array_equal(order.reshape(ref.shape), ref), src
            checked += 1
            continue
        elif kind == "range":
            got = O.range_unique(d[src[1:]])
        elif kind == "find":
            name, needle = src.split("?")
            got = O.find(d[name], float(needle) if "." in needle else int(needle))
        elif kind == "enumerate":
            got = O.enumerate_(int(src[1:]))
        elif kind == "expand":
            got = O.expand_where(d[src[1:]] if src[1:] in d else int(src[1:]))
        elif kind == "reverse":
            got = O.reverse(d[src[1:]])
        elif kind == "at" and src in ("g1@ix", "f1@ix"):
            got = O.at_index(d[src[:2]], d["ix"])
        elif kind == "at" and src == "f1@<f1":
            got = O.at_index(d["f1"], O.grade_up(d["f1"]))
        elif kind == "over":
            got = O.over(src[0], d[src[2:]])
        elif kind == "scan":
            got = O.scan(src[0], d[src[2:]])
        elif kind == "monad" and src[0] in "-_%~":
            got = O.monad(src[0], d[src[1:]])
        elif kind == "dyad":
            table = {"g2+g1": ("+", "g2", "g1"), "g2*2.5": ("*", "g2", 2.5), "g2|10": ("|", "g2", 10),
                     "g2&10": ("&", "g2", 10), "g1!7": ("!", "g1", 7), "g2<10": ("<", "g2", 10),
                     "g2>10": (">", "g2", 10), "g2=10": ("=", "g2", 10), "g2^2": ("^", "g2", 2),
                     "f1^2": ("^", "f1", 2), "f1^0.5": ("^", "f1", 0.5)}
            if src not in table:
                continue
            op, a, b = table[src]
            got = O.dyad(op, d[a], d[b] if isinstance(b, str) else b)
        else:
            continue
        ref = arrays[key]
        got = np.asarray(got)
        assert got.shape == ref.shape, src
        if not case.get("object_scalars"):
            assert got.dtype == ref.dtype, (src, got.dtype, ref.dtype)
        assert np.array_equal(got, ref, equal_nan=True), src
        checked += 1
    assert checked >= 50


def test_reference_suite_vectors(golden):
    """Golden vectors of the reference's own suite (tests/test_suite.py:71-114, 184-189, 451-459)."""
    for v in golden["suite_vectors"]:
        assert v["reference_agrees"], v
        src, exp = v["src"], v["expect"]
        if src[0] in "<>" and src[1] == "[":
            a = np.array([int(t) for t in src[2:-1].split()])
            got = (O.grade_up(a) if src[0] == "<" else O.grade_down(a)).tolist()
        elif src[0] == "=":
            a = np.array([int(t) for t in src[2:-1].split()])
            got = [g.tolist() for g in O.group(a)]
        elif src[0] == "?":
            a = np.array([int(t) for t in src[2:-1].split()])
            got = O.range_unique(a).tolist()
        else:
            lst, needle = src.split("]?")
            a = np.array([int(t) for t in lst[1:].split()])
            got = O.find(a, int(needle)).tolist()
        assert got == exp, v


def test_group_restatement_equals_reference_algorithm():
    """The O(N log N) restatement == the reference's literal algorithm (monads.py:202-203) on random inputs."""
    rng = np.random.default_rng(1)
    for _ in range(100):
        n = int(rng.integers(1, 60))
        a = rng.integers(0, int(rng.integers(1, 12)), n)
        vals, inv = np.unique(a, return_inverse=True)
        ref = [np.where(inv == i)[0] for i in range(len(vals))]
        got = O.group(a)
        assert len(ref) == len(got)
        for x, y in zip(ref, got):
            assert np.array_equal(x, y)


def test_range_restatement_equals_reference_algorithm():
    """First-appearance unique == the reference's set walk (monads.py:284-294)."""
    rng = np.random.default_rng(2)
    for _ in range(100):
        a = rng.integers(-5, 5, int(rng.integers(1, 50)))
        seen, ref = set(), []
        for x in a:
            if str(x) not in seen:
                seen.add(str(x))
                ref.append(x)
        assert np.array_equal(O.range_unique(a), np.asarray(ref))


def test_grouped_stats_variants_agree():
    rng = np.random.default_rng(3)
    k = rng.integers(0, 50, 5000)
    v = np.round(rng.normal(10, 15, 5000), 1)
    a = O.grouped_stats(k, v, 60)
    b = O.grouped_stats_fast(k, v, 60)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3])
    assert np.allclose(a[2], b[2], rtol=1e-12, atol=1e-9)
    # worker.kg semantics on the doc example: {'a': [-7. 6. 1. 3.]} -> '-7.0/0.3/6.0' 
    mn, mx, sm, ct = O.grouped_stats(np.array([0, 0, 0]), np.array([-7.0, 6.0, 2.0]), 1)
    assert O.report_1brc(mn, mx, sm, ct) == {0: "-7.0/0.3/6.0"}
