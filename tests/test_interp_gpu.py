"""Whole stack on a real B200: the REAL KlongPy interpreter (backend='b200', device cuda:0) against the same
interpreter on its NumPy backend, program by program.  Needs klongpy importable on the GPU box — e.g. the optional
install `pip install --no-deps --target baseline/_ref /root/reference` (git-ignored, travels with gpurun); skipped
otherwise.  The CPU twin of this test (tests/test_host_logic.py) runs the same programs on the test double."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def _have_klongpy():
    if os.path.isdir(os.path.join(REF, "klongpy")) and REF not in sys.path:
        sys.path.insert(0, REF)
    try:
        import klongpy  # noqa: F401
        return True
    except Exception:
        return False


pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not _have_klongpy(), reason="klongpy is not importable on this box")]


@pytest.fixture(scope="module")
def interps():
    import klongpy_b200
    from klongpy import KlongInterpreter
    return klongpy_b200.interpreter(device="cuda:0"), KlongInterpreter(backend="numpy"), klongpy_b200


def norm(x, mod):
    if isinstance(x, mod.DeviceArray):
        return x.to_numpy()
    if isinstance(x, np.ndarray) and x.dtype == object:
        out = np.empty(len(x), dtype=object)
        for i, y in enumerate(x):
            out[i] = norm(y, mod)
        return out
    return x


def test_programs_match_numpy_backend(interps):
    from test_host_logic import PROGRAMS
    gpu, ref, mod = interps
    bad = []
    for src in PROGRAMS:
        g = norm(gpu(src), mod)
        r = ref(src)
        ok = ref._backend.kg_equal(g, r)  # the reference's own equality (tests/utils.py:134)
        if ok and isinstance(r, np.ndarray) and r.dtype != object and isinstance(g, np.ndarray):
            ok = g.dtype == r.dtype
        if not ok:
            bad.append((src, g, r))
    assert not bad, bad[:5]


def test_large_vectors_through_the_interpreter(interps):
    gpu, ref, mod = interps
    rng = np.random.default_rng(1)
    n = 3_000_001
    x = rng.integers(0, 256, n) / 256.0
    k = rng.integers(0, 1000, n)
    for klong in (gpu, ref):
        be = klong._backend
        klong["x"], klong["k"] = be.kg_asarray(x), be.kg_asarray(k)
    # grade on tied keys: the reference's np.argsort is unstable (SURVEY.md §0 trap 1) - the contract is the stable one
    stable = np.argsort(k, kind="stable")
    assert np.array_equal(norm(gpu("<k"), mod), stable) and np.array_equal(norm(gpu(">k"), mod), stable[::-1])
    for src in ("+/((x*2)+1)^2", "+\\x", "|/(|\\x)-x", "(+/x*k)%+/k", "k?77", "#?k", "x@<x", "+/x>0.5", "&/x", "|\\x"):
        g, r = norm(gpu(src), mod), ref(src)
        assert np.array_equal(g, r) or (np.asarray(g).dtype.kind == "f" and np.allclose(g, r, rtol=1e-12, atol=0)), src
    # group: ascending key order, ascending indices inside a group
    gg, rg = norm(gpu("=k"), mod), ref("=k")
    assert len(gg) == len(rg) and all(np.array_equal(a, b) for a, b in zip(gg, rg))


def test_autograd_through_the_interpreter(interps):
    gpu, _, mod = interps
    rng = np.random.default_rng(2)
    X, Y = rng.random(100_000) * 2, rng.random(100_000)
    be = gpu._backend
    gpu["X"], gpu["Y"] = be.kg_asarray(X), be.kg_asarray(Y)
    gpu('.bkf("exp")')
    gpu("w1::0.1;b1::0.1")
    gpu("sigmoid::{1%(1+exp(0-x))}")
    gpu("loss::{+/(sigmoid((w1*X)+b1)-Y)^2}")
    g = gpu("loss:>[w1 b1]")
    s = 1 / (1 + np.exp(-(0.1 * X + 0.1)))
    d = 2 * (s - Y) * s * (1 - s)
    ref = np.array([np.sum(d * X), np.sum(d)])
    got = np.array([float(norm(v, mod)) for v in g])
    assert np.allclose(got, ref, rtol=1e-10), (got, ref)
    assert np.allclose(norm(gpu("{+/x^2}:>[1.0 2.0 3.0]"), mod), [2.0, 4.0, 6.0])


def test_table_program_through_the_interpreter(interps):
    from test_host_logic import table_program_checks
    gpu, _, mod = interps
    table_program_checks(gpu, mod)
