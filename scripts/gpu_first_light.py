"""First-light check of every libkgb200 kernel family on a real B200 (numpy used only as the checker)."""
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, ".")
from klongpy_b200 import _lib, ops  # noqa: E402
from klongpy_b200.ops import OPC, OP_ARG, lit  # noqa: E402

dev = torch.device("cuda:0")
ops.set_default_device(dev)
rng = np.random.default_rng(0)
fails = 0


def check(name, fn):
    global fails
    t = time.time()
    try:
        fn()
        torch.cuda.synchronize()
        print(f"ok   {name}  ({time.time() - t:.2f}s)", flush=True)
    except Exception:
        fails += 1
        print(f"FAIL {name}", flush=True)
        traceback.print_exc()


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


def t_map():
    for n in (1, 3, 1000, 4097, 1_000_003):
        a = rng.random(n)
        b = rng.random(n)
        da, db = ops.asarray(a), ops.asarray(b)
        r = ops.binary("add", da, db).to_numpy()
        assert np.array_equal(r, a + b)
        r = (da * 2 + 1).to_numpy()
        assert np.array_equal(r, a * 2 + 1), "fma contraction?"
        r = ((da * 2 + db * 3) / (da + 1) - (db * 0.5)).to_numpy()
        assert np.array_equal(r, (a * 2 + b * 3) / (a + 1) - (b * 0.5))
        # unaligned view
        r = (da[1:] + db[1:]).to_numpy()
        assert np.array_equal(r, a[1:] + b[1:])
        i = rng.integers(-50, 50, n)
        di = ops.asarray(i)
        assert np.array_equal((di * 3 - 1).to_numpy(), i * 3 - 1)
        assert (di / 2).dtype == np.float64 and np.array_equal((di / 2).to_numpy(), i / 2)
        assert np.array_equal((di < 3).to_numpy(), i < 3)
        assert np.array_equal(ops.binary("max", da, 0.5).to_numpy(), np.maximum(a, 0.5))
        assert np.array_equal((da ** 2).to_numpy(), a ** 2)
        assert np.allclose(ops.unary("exp", da).to_numpy(), np.exp(a), rtol=1e-14)


def t_reduce():
    for n in (1, 2, 1000, 4097, 1_000_003, 20_000_001):
        a = rng.integers(0, 256, n) / 256.0
        da = ops.asarray(a)
        code = [OP_ARG, 0] + lit(2) + [OPC["mul"]] + lit(1) + [OPC["add"]] + lit(2) + [OPC["pow"]]
        r = ops.run(code, [da], mode=_lib.MODE_REDUCE, red=_lib.RED_ADD).item()
        ref = np.add.reduce(((a * 2) + 1) ** 2)
        assert r == ref, (n, r, ref)
        u = rng.random(n)
        du = ops.asarray(u)
        r = ops.reduce("add", du).item()
        assert abs(r - np.add.reduce(u)) <= 1e-12 * abs(np.add.reduce(u))
        assert ops.reduce("max", du).item() == u.max() and ops.reduce("min", du).item() == u.min()
        i = rng.integers(-1000, 1000, n)
        assert ops.reduce("add", ops.asarray(i)).item() == i.sum()
    x = np.array([1.0, np.nan, 3.0])
    assert np.isnan(ops.reduce("max", ops.asarray(x)).item())


def t_scan():
    for n in (1, 2, 1000, 4096, 4097, 1_000_003, 20_000_001):
        a = rng.integers(0, 256, n) / 256.0
        r = ops.scan("add", ops.asarray(a)).to_numpy()
        assert np.array_equal(r, np.cumsum(a)), n
        i = rng.integers(-1000, 1000, n)
        assert np.array_equal(ops.scan("add", ops.asarray(i)).to_numpy(), np.cumsum(i)), n
        u = rng.random(n)
        assert np.array_equal(ops.scan("max", ops.asarray(u)).to_numpy(), np.maximum.accumulate(u)), n
        assert np.array_equal(ops.scan("min", ops.asarray(u)).to_numpy(), np.minimum.accumulate(u)), n
        r = ops.scan("add", ops.asarray(u)).to_numpy()
        ref = np.cumsum(u.astype(np.longdouble)).astype(np.float64)
        assert np.max(np.abs(r - ref) / np.abs(ref)) < 1e-12
        # unaligned
        assert np.array_equal(ops.scan("add", ops.asarray(i)[1:]).to_numpy(), np.cumsum(i[1:])), n


def t_scanred():
    for n in (5, 4097, 1_000_003):
        ts = 100 + rng.standard_normal(n) * 0.5
        d = ops.asarray(ts)
        r = ops.run([OP_ARG, 0], [d], mode=_lib.MODE_SCANRED, red=_lib.RED_MAX,
                    code2=[4, OP_ARG, 0, OPC["sub"]], red2=_lib.RED_MAX).item()
        ref = np.max(np.maximum.accumulate(ts) - ts)
        assert r == ref, (n, r, ref)


def t_sort():
    for n in (1, 2, 33, 4096, 4097, 100_003, 3_000_001):
        for keys in (rng.permutation(n), rng.integers(0, 10_000, n), rng.integers(-2**63, 2**63 - 1, n),
                     np.zeros(n, dtype=np.int64), rng.standard_normal(n), rng.integers(0, 7, n).astype(np.float64) - 3):
            d = ops.asarray(keys)
            r = ops.argsort(d).to_numpy()
            ref = np.argsort(keys, kind="stable")
            assert np.array_equal(r, ref), (n, keys.dtype)
            r = ops.argsort(d, descending=True).to_numpy()
            assert np.array_equal(r, ref[::-1]), (n, keys.dtype, "desc")
    k = np.array([3.0, np.nan, -0.0, 0.0, -np.inf, np.inf, 1.0, np.nan])
    assert np.array_equal(ops.argsort(ops.asarray(k)).to_numpy(), np.argsort(k, kind="stable"))
    k32 = rng.integers(-100, 100, 10_001).astype(np.int32)
    assert np.array_equal(ops.argsort(ops.asarray(k32)).to_numpy(), np.argsort(k32, kind="stable"))


def t_group():
    for n, G in ((1, 1), (10, 3), (4097, 100), (1_000_003, 10_000)):
        k = rng.integers(0, G, n)
        order, starts, g = ops.group(ops.asarray(k))
        o = np.argsort(k, kind="stable")
        b = np.flatnonzero(k[o][1:] != k[o][:-1]) + 1
        assert np.array_equal(order.to_numpy(), o)
        assert np.array_equal(starts.to_numpy(), np.concatenate([[0], b, [n]])), (n, G)
        assert g == len(b) + 1


def t_select():
    for n in (1, 31, 4096, 4097, 1_000_003):
        a = rng.integers(0, 5, n)
        d = ops.asarray(a)
        assert np.array_equal(ops.find_equal(d, 3).to_numpy(), np.flatnonzero(a == 3))
        assert np.array_equal(ops.nonzero(d).to_numpy(), np.flatnonzero(a))
        f = a.astype(np.float64) / 2
        assert np.array_equal(ops.find_equal(ops.asarray(f), 1.5).to_numpy(), np.flatnonzero(f == 1.5))
        assert np.array_equal(ops.nonzero(ops.asarray(a > 2)).to_numpy(), np.flatnonzero(a > 2))


def t_unique():
    for n, G in ((1, 1), (10, 3), (4097, 100), (1_000_003, 5000)):
        a = rng.integers(-G, G, n)
        r = ops.unique_first(ops.asarray(a)).to_numpy()
        ref = a[np.sort(np.unique(a, return_index=True)[1])]
        assert np.array_equal(r, ref), (n, G)


def t_misc():
    assert np.array_equal(ops.iota(10).to_numpy(), np.arange(10))
    assert np.array_equal(ops.iota(100_001).to_numpy(), np.arange(100_001))
    a = rng.random(100_003)
    idx = rng.integers(-100_003, 100_003, 50_001)
    assert np.array_equal(ops.gather(ops.asarray(a), ops.asarray(idx)).to_numpy(), a[idx])
    assert np.array_equal(ops.reverse(ops.asarray(a)).to_numpy(), a[::-1])
    assert np.array_equal(ops.asarray(a)[::-1].to_numpy(), a[::-1])
    assert np.array_equal(ops.asarray(a)[10:5:-2].to_numpy(), a[10:5:-2])
    c = rng.integers(0, 4, 10_001)
    assert np.array_equal(ops.expand(ops.asarray(c)).to_numpy(), np.repeat(np.arange(len(c)), c))
    da = ops.asarray(a)
    assert ops.array_equal(da, ops.asarray(a.copy()))
    b = a.copy()
    b[77_777] += 1
    assert not ops.array_equal(da, ops.asarray(b))
    m = rng.random((7, 5))
    rows = np.array([3, 0, 6, 3])
    assert np.array_equal(ops.gather(ops.asarray(m), ops.asarray(rows)).to_numpy(), m[rows])


def t_groupagg():
    for n, G in ((10, 3), (100_003, 100), (3_000_001, 10_000)):
        k = rng.integers(0, G, n)
        v = np.round(rng.normal(10, 15, n), 1)
        mn, mx, sm, ct = ops.groupagg(ops.asarray(k), ops.asarray(v), G)
        rmin = np.full(G, np.inf)
        rmax = np.full(G, -np.inf)
        rsum = np.zeros(G)
        np.minimum.at(rmin, k, v)
        np.maximum.at(rmax, k, v)
        np.add.at(rsum, k, v)
        rcnt = np.bincount(k, minlength=G)
        assert np.array_equal(mn.to_numpy(), rmin) and np.array_equal(mx.to_numpy(), rmax)
        assert np.array_equal(ct.to_numpy(), rcnt)
        assert np.allclose(sm.to_numpy(), rsum, rtol=1e-12, atol=1e-9)


def perf():
    peak = 6546.0
    n = 1 << 28
    x = ops.asarray(torch.rand(n, dtype=torch.float64, device=dev))
    code = [OP_ARG, 0] + lit(2) + [OPC["mul"]] + lit(1) + [OPC["add"]] + lit(2) + [OPC["pow"]]
    ms = timeit(lambda: ops.run(code, [x], mode=_lib.MODE_REDUCE, red=_lib.RED_ADD))
    print(f"perf fused reduce  n=2^28: {ms:.3f} ms  {8 * n / ms / 1e6:.0f} GB/s  frac {8 * n / ms / 1e6 / peak:.2f}")
    ms = timeit(lambda: ops.scan("add", x))
    print(f"perf scan +\\x      n=2^28: {ms:.3f} ms  {16 * n / ms / 1e6:.0f} GB/s  frac {16 * n / ms / 1e6 / peak:.2f}")
    y = ops.asarray(torch.rand(n, dtype=torch.float64, device=dev))
    ms = timeit(lambda: ops.binary("add", x, y))
    print(f"perf a+b           n=2^28: {ms:.3f} ms  {24 * n / ms / 1e6:.0f} GB/s  frac {24 * n / ms / 1e6 / peak:.2f}")
    del y
    ns = 100_000_000
    for name, keys in (("perm", torch.randperm(ns, device=dev)), ("10k", torch.randint(0, 10_000, (ns,), device=dev)),
                       ("full", torch.randint(-2**62, 2**62, (ns,), device=dev))):
        k = ops.asarray(keys)
        ms = timeit(lambda: ops.argsort(k), reps=3)
        print(f"perf argsort {name:5s} n=1e8: {ms:.3f} ms  {ns / ms / 1e3:.0f} Mkeys/s  (16B/elem: {16 * ns / ms / 1e6:.0f} GB/s)")
    k = ops.asarray(torch.randint(0, 10_000, (ns,), device=dev))
    ms = timeit(lambda: ops.group(k), reps=3)
    print(f"perf group 10k     n=1e8: {ms:.3f} ms  {ns / ms / 1e3:.0f} Mkeys/s")
    v = ops.asarray(torch.randn(ns, dtype=torch.float64, device=dev))
    ms = timeit(lambda: ops.groupagg(k, v, 10_000), reps=3)
    print(f"perf groupagg 10k  n=1e8: {ms:.3f} ms  {ns / ms / 1e3:.0f} Mrows/s  {16 * ns / ms / 1e6:.0f} GB/s frac {16 * ns / ms / 1e6 / peak:.2f}")
    ms = timeit(lambda: ops.find_equal(k, 77), reps=3)
    print(f"perf find          n=1e8: {ms:.3f} ms  {8 * ns / ms / 1e6:.0f} GB/s")


for nm, f in (("map", t_map), ("reduce", t_reduce), ("scan", t_scan), ("scanred", t_scanred), ("misc", t_misc),
              ("select", t_select), ("sort", t_sort), ("group", t_group), ("unique", t_unique), ("groupagg", t_groupagg)):
    check(nm, f)
if "--perf" in sys.argv:
    check("perf", perf)
print("FAILS", fails)
sys.exit(1 if fails else 0)
