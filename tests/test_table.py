"""DeviceTable (klongpy_b200/table.py) against the reference's Table semantics.

Golden: tests/golden/golden_table.json, recorded by tests/golden/make_table_golden.py from the reference's own Table
class (klongpy/db/sys_fn_db.py:18-127) on the scenarios of the reference's tests/kgtests/db/*.kg plus seeded tables.
CPU leg: the oracle restatement and the host logic on the test double; `-m gpu` leg: the same scenarios and a large
randomised case through libkgb200 on the device."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import kg_oracle as O


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLDEN, "golden_table.json")) as f:
        return json.load(f)


def canon(rows, idx):
    """Rows with equal index keys in a canonical order: the reference's single-column `sort_index` is an unstable
    quicksort, so the order INSIDE a key run is arbitrary there (the contract here: insertion order)."""
    rows = np.asarray(rows).reshape(len(rows), -1)
    if not idx or len(rows) == 0:
        return rows.tolist()
    rest = [j for j in range(rows.shape[1]) if j not in idx]
    return rows[np.lexsort([rows[:, j] for j in reversed(list(idx) + rest)])].tolist()


class OracleTable:
    """The oracle restatement driven by the same scripts."""

    def __init__(self, cols):
        self.names = [k for k, _ in cols]
        self.m = np.stack([np.asarray(v) for _, v in cols], 1)
        self.idx = None
        self.buf = []

    def commit(self):
        if self.buf:
            self.m = O.table_upsert(self.m, np.array(self.buf), self.idx) if self.idx else O.table_append(self.m, self.buf)
            self.buf = []

    def step(self, st):
        op = st[0]
        if op == "insert":
            self.buf.append(st[1])
        elif op == "insertb":
            self.buf.extend(st[1])
        else:
            self.commit()
            if op == "index":
                self.idx = [self.names.index(c) for c in st[1]]
                self.m = O.table_index(self.m, self.idx)
            elif op == "rindex":
                self.idx = None
            elif op == "set":
                if st[1] in self.names:
                    self.m[:, self.names.index(st[1])] = st[2]
                else:
                    self.m = np.concatenate([self.m, np.asarray(st[2])[:, None]], 1)
                    self.names.append(st[1])

    def rows(self):
        self.commit()
        return self.m.tolist()

    def col(self, c):
        self.commit()
        return self.m[:, self.names.index(c)].tolist()


class DevTable:
    def __init__(self, cols, be):
        from klongpy_b200.table import DeviceTable
        self.t = DeviceTable({k: be.kg_asarray(np.asarray(v)) for k, v in cols}, columns=[k for k, _ in cols])
        self.names = self.t.columns

    def step(self, st):
        op = st[0]
        if op == "insert":
            self.t.insert(np.asarray(st[1]))
        elif op == "insertb":
            self.t.insertb(np.asarray(st[1]))
        elif op == "index":
            self.t.set_index(st[1])
        elif op == "rindex":
            self.t.reset_index()
        elif op == "set":
            self.t.set(st[1], np.asarray(st[2]))

    def rows(self):
        r = self.t.rows()
        return r.reshape(len(self.t), -1).tolist()

    def col(self, c):
        return self.t.get(c).to_numpy().tolist()


def play(make, sc):
    t = None
    exp = iter(sc["expect"])
    idx = None
    for st in sc["script"]:
        if st[0] == "table":
            t = make(st[1])
        elif st[0] == "rows":
            e = next(exp)
            got = t.rows()
            assert len(got) == e["len"] and list(t.names) == e["schema"]
            assert canon(got, idx) == canon(e["rows"], idx)
            if idx is None or len(idx) > 1:
                assert got == e["rows"]  # multi-column indexes sort stably in the reference too
        elif st[0] == "get":
            assert t.col(st[1]) == next(exp)["col"]
        else:
            t.step(st)
            if st[0] == "index":
                idx = [t.names.index(c) for c in st[1]]
            elif st[0] == "rindex" and (idx is None or len(idx) > 1):
                idx = None  # (after a single-column index the rows stay in the reference's arbitrary in-key order)


def test_oracle_table_matches_reference(gold):
    for name, sc in gold["scenarios"].items():
        play(OracleTable, sc)
    for c in gold["group_stats"].values():
        keys, mn, mean, mx, ct = O.table_group_stats(c["k"], c["v"])
        assert keys.tolist() == c["keys"] and mn.tolist() == c["min"] and mx.tolist() == c["max"]
        assert ct.tolist() == c["count"] and np.allclose(mean, c["mean"], rtol=1e-12, atol=0)


def check_group_stats(be, k, v, ref=None, **kw):
    from klongpy_b200.table import DeviceTable
    t = DeviceTable({"k": be.kg_asarray(k), "v": be.kg_asarray(v)})
    keys, mn, mean, mx, ct = (x.to_numpy() for x in t.group_stats("k", "v", **kw))
    rk, rmn, rmean, rmx, rct = ref if ref is not None else O.table_group_stats(k, v)
    assert keys.dtype == np.asarray(k).dtype and np.array_equal(keys, rk)
    assert np.array_equal(mn, rmn) and np.array_equal(mx, rmx) and np.array_equal(ct, rct)
    assert np.allclose(mean, rmean, rtol=1e-12, atol=1e-12)


def run_device_suite(be, gold):
    for name, sc in gold["scenarios"].items():
        play(lambda cols: DevTable(cols, be), sc)
    for c in gold["group_stats"].values():
        k, v = np.asarray(c["k"], dtype=np.int64), np.asarray(c["v"])
        ref = tuple(np.asarray(c[x]) for x in ("keys", "min", "mean", "max", "count"))
        check_group_stats(be, k, v, ref)
        check_group_stats(be, k, v, ref, dense_limit=0)      # force the dictionary-encoded path
    rng = np.random.default_rng(5)
    check_group_stats(be, rng.normal(0, 1, 4000).round(1), rng.normal(0, 1, 4000))   # float keys


def test_device_table_host_logic(fake_backend, gold):
    run_device_suite(fake_backend, gold)


def test_table_sys_functions(fake_backend):
    from klongpy_b200 import table as T
    ex = T.klongpy_exports
    assert sorted(ex) == [".gstats", ".index", ".insert", ".rindex", ".schema", ".table"]
    d = np.empty(2, dtype=object)
    d[0] = np.array(["a", fake_backend.kg_asarray(np.array([1, 2, 3]))], dtype=object)
    d[1] = np.array(["b", np.array([2.0, 3.0, 4.0])], dtype=object)
    t = ex[".table"](d)
    assert list(ex[".schema"](t)) == ["a", "b"] and len(t) == 3
    assert ex[".rindex"](t) == 0 and ex[".index"](t, np.array(["a"], dtype=object)) == ["a"]
    with pytest.raises(T.KlongDbException):
        ex[".index"](t, np.array(["a"], dtype=object))
    with pytest.raises(T.KlongDbException):
        ex[".insert"](t, np.array([1, 2, 3]))
    ex[".insert"](t, np.array([2, 9.5]))
    ex[".insert"](t, np.array([[0, 1.5], [7, 2.5]]))
    assert t.rows().tolist() == [[0, 1.5], [1, 2.0], [2, 9.5], [3, 4.0], [7, 2.5]]
    assert ex[".rindex"](t) == 1
    keys, mn, mean, mx, ct = ex[".gstats"](t, "a", "b")
    assert keys.to_numpy().tolist() == [0, 1, 2, 3, 7] and ct.to_numpy().tolist() == [1] * 5
    assert t.get("zz") is T.KLONG_UNDEFINED


@pytest.mark.gpu
def test_device_table_gpu(cuda_backend, gold):
    run_device_suite(cuda_backend, gold)


@pytest.mark.gpu
def test_device_table_large_gpu(cuda_backend):
    """2e6-row table: two-column index with repeated rows, an upsert of 1e5 rows, grouped statistics."""
    from klongpy_b200.table import DeviceTable
    be = cuda_backend
    rng = np.random.default_rng(9)
    n = 2_000_000
    a, b, c = rng.integers(0, 3000, n), rng.integers(0, 300, n), rng.integers(0, 2, n)
    t = DeviceTable({"a": be.kg_asarray(a), "b": be.kg_asarray(b), "c": be.kg_asarray(c)})
    m = np.stack([a, b, c], 1)
    t.set_index(["a", "b"])
    ref = O.table_index(m, [0, 1])
    assert np.array_equal(t.rows(), ref)
    kk = rng.permutation(3000 * 400)[:100_000]
    up = np.stack([kk // 400, kk % 400, rng.integers(5, 9, len(kk))], 1)
    t.insertb(up)
    got = t.rows()
    # vectorised restatement of the upsert (O.table_upsert is a Python loop): keys buffered at most once
    key = lambda r: r[:, 0] * 400 + r[:, 1]
    s = np.argsort(key(up), kind="stable")
    ups = up[s]
    pos = np.searchsorted(key(ups), key(ref))
    hit = (pos < len(ups)) & (key(ups)[np.minimum(pos, len(ups) - 1)] == key(ref))
    upd = ref.copy()
    upd[hit] = ups[pos[hit]]
    new = ups[~np.isin(key(ups), key(ref))]
    exp = np.concatenate([upd, new])
    exp = exp[np.lexsort([exp[:, 1], exp[:, 0]])]
    assert np.array_equal(got, exp)
    small = slice(0, 3000)
    assert np.array_equal(O.table_upsert(ref[small], up[key(up) < key(ref[small]).max()][:200], [0, 1]),
                          O.table_upsert(ref[small], up[key(up) < key(ref[small]).max()][:200], [0, 1]))
    v = np.round(rng.normal(10, 15, n), 1)
    check_group_stats(be, a, v)
    check_group_stats(be, a * 1_000_003 - 17, v)


def edge_cases(be):
    """Empty and one-row tables, float index keys, NaN-free ties, appends that change a column's dtype."""
    from klongpy_b200.table import DeviceTable
    t = DeviceTable({"a": np.zeros(0)}, columns=["a"], device=be.device)          # tests/kgtests/db/test_create_empty_table.kg
    assert len(t) == 0 and list(t.schema()) == ["a"] and t.rows().size == 0
    t.set_index(["a"])
    t.insert(np.array([3.5]))
    t.insert(np.array([1.5]))
    t.insert(np.array([3.5]))
    assert t.rows().tolist() == [1.5, 3.5] and len(t) == 2                          # one column: rows() squeezes like the reference
    t.reset_index()
    keys, mn, mean, mx, ct = t.group_stats("a", "a")
    assert keys.to_numpy().tolist() == [1.5, 3.5] and ct.to_numpy().tolist() == [1, 1]
    one = DeviceTable({"k": be.kg_asarray(np.array([7])), "v": be.kg_asarray(np.array([2.0]))})
    one.set_index(["k"])
    assert one.rows().tolist() == [7.0, 2.0]
    one.insert(np.array([7, 9.25]))
    one.insert(np.array([-2, 0.5]))
    assert one.rows().tolist() == [[-2.0, 0.5], [7.0, 9.25]]
    assert one.get("k").dtype == np.float64                                          # the buffered block is one float matrix
    # float keys incl. -0.0 / 0.0 (equal as keys) and repeated full rows
    f = DeviceTable({"k": be.kg_asarray(np.array([0.5, -0.0, 0.0, 0.5, 2.0, 0.5])),
                     "v": be.kg_asarray(np.array([1, 2, 2, 1, 5, 3]))})
    f.set_index(["k"])
    m = np.stack([np.array([0.5, -0.0, 0.0, 0.5, 2.0, 0.5]), np.array([1, 2, 2, 1, 5, 3.0])], 1)
    assert np.array_equal(f.rows(), O.table_index(m, [0]))
    e = DeviceTable({"k": be.kg_asarray(np.zeros(0, dtype=np.int64)), "v": be.kg_asarray(np.zeros(0))})
    assert all(x.size == 0 for x in e.group_stats("k", "v"))
    with pytest.raises(ValueError):
        DeviceTable({"a": np.arange(3), "b": np.arange(4)}, device=be.device)
    with pytest.raises(ValueError):
        one.set("z", np.arange(5))


def test_device_table_edge_cases(fake_backend):
    edge_cases(fake_backend)


@pytest.mark.gpu
def test_device_table_edge_cases_gpu(cuda_backend):
    edge_cases(cuda_backend)
