"""One interpreter, several devices (klongpy_b200/sharded.py): long vectors are block-sharded behind the backend seam and
Klong programs — compiled expressions, eager verbs, adverbs — run unchanged.  CPU suite: two / three logical devices on
the NumPy test double (all blocks live on the double's device; the partition, the exchanges and the seeded scans are
the real code).  GPU suite: the same programs on two real B200s (`gpurun --gpus 2`)."""
import sys
import warnings

import numpy as np
import pytest

from conftest import REFERENCE, have_reference


def _programs(rng, n):
    x = rng.integers(0, 256, n) / 256.0
    y = rng.integers(1, 64, n) / 8.0
    k = rng.integers(0, 37, n)
    progs = [
        ("+/x", lambda: np.add.reduce(x)),
        ("{+/((x*2)+1)^2}(x)", lambda: np.add.reduce(((x * 2) + 1) ** 2)),
        ("+\\x", lambda: np.cumsum(x)),
        ("|/x", lambda: x.max()),
        ("&/y", lambda: y.min()),
        ("|\\x", lambda: np.maximum.accumulate(x)),
        ("(x*2)+y", lambda: (x * 2) + y),
        ("(+/x*y)%+/y", lambda: np.add.reduce(x * y) / np.add.reduce(y)),
        ("x-(+/x)%#x", lambda: x - np.add.reduce(x) / len(x)),
        ("+/x>y", lambda: np.add.reduce((x > y) * 1)),
        ("+\\(x*y)+1", lambda: np.cumsum(x * y + 1)),
        ("#x", lambda: len(x)),
    ]
    return x, y, k, progs


def _check(klong, mod, progs, sharded_names):
    from klongpy_b200 import sharded
    for src, want in progs:
        got = klong(src)
        w = want()
        if isinstance(got, mod.DeviceArray):
            if src in sharded_names:
                assert sharded.is_sharded(got), src      # elementwise / scan results stay sharded
            got = got.to_numpy()
        if np.ndim(w) == 0:
            assert float(got) == float(w), src            # dyadic data: exact in any association
        else:
            assert np.array_equal(np.asarray(got), w), src


@pytest.mark.skipif(not have_reference(), reason="reference klongpy not available on this box")
@pytest.mark.parametrize("k", [2, 3])
def test_klong_programs_on_sharded_vectors_cpu_double(k):
    sys.path.insert(0, REFERENCE)
    import fake_lib
    dev = fake_lib.install()
    import klongpy_b200
    from klongpy_b200 import lazy, ops, sharded
    klong = klongpy_b200.interpreter(device=dev)
    be = klong._backend
    be._devices = [dev] * k
    old_min, sharded.SHARD_MIN = sharded.SHARD_MIN, 1000
    try:
        rng = np.random.default_rng(5 + k)
        n = 10_007
        x, y, keys, progs = _programs(rng, n)
        klong["x"], klong["y"] = be.kg_asarray(x), be.kg_asarray(y)
        assert sharded.is_sharded(klong["x"]) and len(sharded.shards_of(klong["x"])) == k
        assert [s.size for s in sharded.shards_of(klong["x"])] == [hi - lo for lo, hi in sharded.bounds(n, k)]
        with warnings.catch_warnings():
            warnings.simplefilter("error", sharded.ShardGatherWarning)   # none of these may gather onto one device
            _check(klong, klongpy_b200, progs, {"+\\x", "|\\x", "(x*2)+y", "x-(+/x)%#x", "+\\(x*y)+1"})
            mn, mx, sm, ct = (t.to_numpy() for t in ops.groupagg(be.kg_asarray(keys), klong["y"], 37))
        from oracle import kg_oracle as O
        rmn, rmx, rsm, rct = O.grouped_stats(keys, y, 37)
        assert np.array_equal(mn, rmn) and np.array_equal(mx, rmx) and np.array_equal(ct, rct) and np.array_equal(sm, rsm)
        # no sharded kernel (grade): the vector is gathered, announced, and the answer is still right
        with pytest.warns(sharded.ShardGatherWarning):
            assert np.array_equal(klong("<x").to_numpy(), np.argsort(x, kind="stable"))
        # a one-device vector of the same length combines with a sharded one
        z = ops.asarray(np.arange(n, dtype=np.float64), device=dev)
        assert np.array_equal((klong["x"] + z).to_numpy(), x + np.arange(n))
    finally:
        sharded.SHARD_MIN = old_min
        lazy.enable(True)


@pytest.mark.gpu
def test_klong_programs_on_two_gpus():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    try:
        import klongpy  # noqa: F401
    except ImportError:
        import klongpy_b200  # noqa: F401 - puts baseline/_ref on the path
    import klongpy_b200
    from klongpy import KlongInterpreter
    from klongpy_b200 import ops, sharded, verbs
    klongpy_b200.register()
    klong = verbs.install(KlongInterpreter(backend="b200", device="cuda:*"))
    be = klong._backend
    assert be._devices and len(be._devices) >= 2
    rng = np.random.default_rng(9)
    n = 20_000_003
    x, y, keys, progs = _programs(rng, n)
    klong["x"], klong["y"] = be.kg_asarray(x), be.kg_asarray(y)
    sh = sharded.shards_of(klong["x"])
    assert len({s.tensor.device for s in sh}) == len(be._devices)
    with warnings.catch_warnings():
        warnings.simplefilter("error", sharded.ShardGatherWarning)
        _check(klong, klongpy_b200, progs, {"+\\x", "|\\x", "(x*2)+y", "x-(+/x)%#x", "+\\(x*y)+1"})
        mn, mx, sm, ct = (t.to_numpy() for t in ops.groupagg(be.kg_asarray(keys), klong["y"], 37))
    from oracle import kg_oracle as O
    rmn, rmx, rsm, rct = O.grouped_stats(keys, y, 37)
    assert np.array_equal(mn, rmn) and np.array_equal(mx, rmx) and np.array_equal(ct, rct)
    assert np.allclose(sm, rsm, rtol=1e-12)
    with pytest.warns(sharded.ShardGatherWarning):
        assert np.array_equal(klong("<x").to_numpy(), np.argsort(x, kind="stable"))
