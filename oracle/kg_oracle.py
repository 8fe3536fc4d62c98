"""CPU oracle for the KlongPy array hot path — TEST INFRASTRUCTURE ONLY.

A NumPy restatement of the reference's algorithm for every primitive on the path.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
module; the product (`klongpy_b200/`) never does and has no CPU fallback.

Pinning: `tests/golden/make_golden.py` runs the REAL reference (`/root/reference`, NumPy backend) on seeded
inputs and commits inputs + outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks every function
here against those vectors and against the reference's own golden vectors lifted from
`tests/test_suite.py` (grade :71-101, group :103-114, range :184-189, find :451-459, over :1176-1196,
scan :1243-1262).  All arithmetic the reference delegates to NumPy (unpinned `numpy>=2.0`,
pyproject.toml:31) is delegated to NumPy here too, so the restatement is exact wherever NumPy is
deterministic; where the reference itself is not (unstable `np.argsort` on ties) the oracle
implements the language contract (stable) and says so.
"""
import itertools

import numpy as np

_BINOPS = {"+": np.add, "-": np.subtract, "*": np.multiply, "%": np.divide, "^": np.power}
_CMPS = {"=": np.equal, ">": np.greater, "<": np.less}
_REDUCE = {"+": np.add, "*": np.multiply, "|": np.maximum, "&": np.minimum}


# --------------------------------------------------------------------------------------------- compiled IR
def eval_ir(ir, env):
    """Evaluate a compiler IR tree (klongpy/compiler.py:37-113) exactly as the NumPy backend's generated
    source does (klongpy/backends/numpy_backend.py:116-178): `%`->`/`, `^`->`**`, cmp->`((l==r)*1)`,
    reduce->`np.<ufunc>.reduce`, scan->`np.cumsum/np.cumprod`.  `|\\` and `&\\` are not compilable on NumPy
    (:173-175); the interpreter then runs `itertools.accumulate` with max/min (adverbs.py:333), which equals
    `np.maximum/minimum.accumulate`."""
    with np.errstate(all="ignore"):
        return _eval(ir, env)


def _eval(ir, env):
    t = ir[0]
    if t == "literal":
        return ir[1]
    if t == "var":
        return env[ir[1]]
    if t == "binop":
        l, r = _eval(ir[2], env), _eval(ir[3], env)
        op = ir[1]
        if op == "+":
            return l + r
        if op == "-":
            return l - r
        if op == "*":
            return l * r
        if op == "%":
            return l / r
        if op == "^":
            return l ** r
        raise ValueError(op)
    if t == "cmp":
        l, r = _eval(ir[2], env), _eval(ir[3], env)
        return _CMPS[ir[1]](l, r) * 1
    if t == "negate":
        return -_eval(ir[1], env)
    if t == "reduce":
        return _REDUCE[ir[1]].reduce(_eval(ir[2], env))
    if t == "scan":
        a = _eval(ir[2], env)
        if ir[1] == "+":
            return np.cumsum(a)
        if ir[1] == "*":
            return np.cumprod(a)
        return _REDUCE[ir[1]].accumulate(a)
    raise ValueError(f"unknown IR node {t}")


# --------------------------------------------------------------------------------------------- eager dyads / monads
def dyad(op, a, b):
    """Eager dyads as the verbs compute them (klongpy/dyads.py): `+ - *` (:6-21, :976-991, :725-739),
    `%` true divide (:214-232), `|`/`&` NaN-propagating max/min (:649-700), `!` np.fmod (:785-803),
    `< > =` -> int64 0/1 via kg_truth (:588-609, :703-722, :264-294; types.py:339-340),
    `^` power with the all-whole -> int64 rule (:742-761)."""
    with np.errstate(all="ignore"):
        if op == "+":
            return np.add(a, b)
        if op == "-":
            return np.subtract(a, b)
        if op == "*":
            return np.multiply(a, b)
        if op == "%":
            return np.divide(a, b)
        if op == "|":
            return np.maximum(a, b)
        if op == "&":
            return np.minimum(a, b)
        if op == "!":
            return np.fmod(a, b)
        if op == "<":
            return np.less(a, b) * 1
        if op == ">":
            return np.greater(a, b) * 1
        if op == "=":
            return np.equal(a, b) * 1
        if op == "^":
            r = np.power(float(a) if isinstance(a, (int, np.integer)) else a, b)
            if np.ndim(r) > 0:
                whole = bool((np.trunc(r) == r).all())
            else:
                whole = bool(np.trunc(r) == r)
            return np.asarray(r, dtype=int) if (whole and np.ndim(r) > 0) else (int(r) if whole else r)
    raise ValueError(op)


def monad(op, a):
    """Eager monads (klongpy/monads.py): `-` (:224-237), `_` floor->int64 (:101-119, base.py:198-203),
    `%` reciprocal (:298-316), `#` abs for numbers (:408-427), `~` not (:240-257)."""
    with np.errstate(all="ignore"):
        if op == "-":
            return np.negative(a)
        if op == "_":
            return np.floor(np.asarray(a, dtype=float)).astype(int)
        if op == "%":
            return np.reciprocal(np.asarray(a, dtype=float))
        if op == "#":
            return np.abs(a)
        if op == "~":
            return np.logical_not(np.asarray(a)) * 1
    raise ValueError(op)


# --------------------------------------------------------------------------------------------- adverbs
def over(op, a):
    """f/a (klongpy/adverbs.py:210-246): `+ - * %` -> ufunc.reduce, `& |` -> np.min / np.max (1-d)."""
    a = np.asarray(a)
    if a.ndim == 0 or a.size == 1:
        return a if a.ndim == 0 else a[0]
    with np.errstate(all="ignore"):
        if op == "+":
            return np.add.reduce(a)
        if op == "-":
            return np.subtract.reduce(a)
        if op == "*":
            return np.multiply.reduce(a)
        if op == "%":
            return np.divide.reduce(a)
        if op == "&":
            return np.min(a)
        if op == "|":
            return np.max(a)
    raise ValueError(op)


def scan(op, a):
    """f\\a (klongpy/adverbs.py:316-334): `+ - * %` -> ufunc.accumulate; others itertools.accumulate."""
    a = np.asarray(a)
    with np.errstate(all="ignore"):
        if op == "+":
            return np.add.accumulate(a)
        if op == "-":
            return np.subtract.accumulate(a)
        if op == "*":
            return np.multiply.accumulate(a)
        if op == "%":
            return np.divide.accumulate(a)
        if op == "|":
            return np.asarray(list(itertools.accumulate(a, np.maximum))) if a.size < 4096 else np.maximum.accumulate(a)
        if op == "&":
            return np.asarray(list(itertools.accumulate(a, np.minimum))) if a.size < 4096 else np.minimum.accumulate(a)
    raise ValueError(op)


def scan_exact(a):
    """High-precision running sum (long double) used as the accuracy reference for `+\\` on float64: NumPy's
    own sequential cumsum is 6e-13 off at 1e8 elements (SURVEY.md §0 trap 3)."""
    return np.cumsum(np.asarray(a, dtype=np.longdouble)).astype(np.float64)


# --------------------------------------------------------------------------------------------- grade / group / find / range
def grade_up(a):
    """<a (klongpy/monads.py:143-168 -> writer.py:97-122 -> numpy_backend.py:75-80).  The reference calls
    np.argsort with the default (unstable) kind; the language contract (writer.py:97-108 docstring,
    tests/test_suite.py:82,98) is ties-by-position, i.e. the STABLE permutation — identical to the reference
    whenever keys are distinct."""
    return np.argsort(np.asarray(a), kind="stable")


def grade_down(a):
    """>a = reverse of the stable ascending grade (numpy_backend.py:78-79: indices[::-1].copy())."""
    return grade_up(a)[::-1].copy()


def group(a):
    """=a (klongpy/monads.py:182-204): np.unique(return_inverse) + one np.where per value, i.e. groups in
    ascending key order with ascending indices inside.  Restated as stable argsort split at key changes
    (O(N log N) instead of O(G*N)); equality with the reference is asserted in tests at small N."""
    a = np.asarray(a)
    if a.size == 0:
        return []
    o = np.argsort(a, kind="stable")
    s = a[o]
    if s.dtype.kind == "f":
        neq = ~((s[1:] == s[:-1]) | (np.isnan(s[1:]) & np.isnan(s[:-1])))
    else:
        neq = s[1:] != s[:-1]
    b = np.flatnonzero(neq) + 1
    return np.split(o, b)


def group_csr(a):
    """group() in CSR form: (order, starts) with starts[G] = n."""
    a = np.asarray(a)
    g = group(a)
    order = np.concatenate(g) if len(g) else np.zeros(0, dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum([len(x) for x in g])]).astype(np.int64)
    return order.astype(np.int64), starts


def find(a, b):
    """a?b for an atom b (klongpy/dyads.py:342): np.where(np.asarray(a) == b)[0]."""
    return np.where(np.asarray(a) == b)[0]


def range_unique(a):
    """?a (klongpy/monads.py:260-295): unique elements in order of first appearance (the reference walks the
    list with a Python set, :284-294)."""
    a = np.asarray(a)
    if a.size == 0:
        return a
    keys = a
    if a.dtype.kind == "f":
        # the reference keys its set on str(x): -0.0 and 0.0 are DIFFERENT elements, all NaNs are one
        b = a.astype(np.float64, copy=True)
        b[np.isnan(b)] = np.nan
        keys = b.view(np.int64)
    _, idx = np.unique(keys, return_index=True)
    return a[np.sort(idx)]


def enumerate_(n):
    """!n (klongpy/monads.py:52): np.arange(int(n))."""
    return np.arange(int(n))


def expand_where(a):
    """&a (klongpy/monads.py:79): np.repeat(np.arange(len(a)), a)."""
    arr = a if np.ndim(a) > 0 else [a]
    return np.repeat(np.arange(len(arr)), arr)


def at_index(a, idx):
    """a@b with an index list (klongpy/dyads.py:176-181): [a[x] for x in b]."""
    return np.asarray(a)[np.asarray(idx)]


def reverse(a):
    """|a (klongpy/monads.py:319-340)."""
    return np.asarray(a)[::-1].copy()


# --------------------------------------------------------------------------------------------- grouped aggregation
def grouped_stats(keys, vals, n_groups):
    """examples/1brc/worker.kg:3-5 — collect = [min max sum count] per station, merge = (&, |, +, +) —
    restated over int64 station ids in [0, n_groups): returns (min, max, sum, count) ordered by id, with
    +inf / -inf / 0 / 0 for absent stations.  sum is accumulated per station in row order (np.add.at)."""
    keys = np.asarray(keys)
    vals = np.asarray(vals, dtype=np.float64)
    mn = np.full(n_groups, np.inf)
    mx = np.full(n_groups, -np.inf)
    sm = np.zeros(n_groups)
    np.minimum.at(mn, keys, vals)
    np.maximum.at(mx, keys, vals)
    np.add.at(sm, keys, vals)
    ct = np.bincount(keys, minlength=n_groups).astype(np.int64)
    return mn, mx, sm, ct


def grouped_stats_fast(keys, vals, n_groups):
    """Same result as grouped_stats for min/max/count (sum up to rounding), via sort + reduceat: the variant
    timed as the CPU baseline at sizes where ufunc.at is too slow."""
    keys = np.asarray(keys)
    vals = np.asarray(vals, dtype=np.float64)
    o = np.argsort(keys, kind="stable")
    ks, vs = keys[o], vals[o]
    heads = np.flatnonzero(np.concatenate([[True], ks[1:] != ks[:-1]]))
    ids = ks[heads]
    mn = np.full(n_groups, np.inf)
    mx = np.full(n_groups, -np.inf)
    sm = np.zeros(n_groups)
    ct = np.zeros(n_groups, dtype=np.int64)
    if len(heads):
        mn[ids] = np.minimum.reduceat(vs, heads)
        mx[ids] = np.maximum.reduceat(vs, heads)
        sm[ids] = np.add.reduceat(vs, heads)
        ct[ids] = np.diff(np.concatenate([heads, [len(ks)]]))
    return mn, mx, sm, ct


def parse_1brc(text):
    """examples/1brc/par.py:57-88: `for line in f: location, measurement = line.split(";")`, stations kept as
    strings, `np.asarray(temps, dtype=float)`; worker.kg:5 then groups by station name (`=s`) and par.kg:4 sorts the
    names.  Restated for a whole buffer: returns (names sorted ascending, station id per row as an index into that
    list, float64 temperature per row).  `float()` is Python's correctly rounded decimal conversion."""
    if isinstance(text, (bytes, bytearray, memoryview)):
        text = bytes(text).decode("utf-8")
    stations, temps = [], []
    for line in text.split("\n"):
        if not line:
            continue
        location, measurement = line.split(";")
        stations.append(location)
        temps.append(float(measurement))
    names = sorted(set(stations), key=lambda s: s.encode("utf-8"))
    index = {s: i for i, s in enumerate(names)}
    ids = np.fromiter((index[s] for s in stations), dtype=np.int64, count=len(stations))
    return names, ids, np.asarray(temps, dtype=float)


def report_1brc(mn, mx, sm, ct):
    """examples/1brc/par.kg:4-6: 'min/mean/max' with one decimal, stations ascending, mean = sum % count."""
    out = {}
    for s in range(len(ct)):
        if ct[s]:
            out[s] = f"{mn[s]:.1f}/{sm[s] / ct[s]:.1f}/{mx[s]:.1f}"
    return out


# --------------------------------------------------------------------------------------------- tables
# Restatement of the reference's columnar Table (klongpy/db/sys_fn_db.py:18-127; pandas underneath) as plain
# NumPy row-matrix operations.  Pinned by tests/golden/golden_table.json, which tests/golden/make_table_golden.py
# records from the reference's own Table class (the reference's tests/kgtests/db/*.kg scenarios + seeded tables).
def table_index(rows, idx):
    """`Table.set_index` (sys_fn_db.py:55-69): sort by the index columns (first is the major key; rows with equal
    keys keep their order) and drop rows that repeat an earlier row in every column."""
    rows = np.asarray(rows)
    if rows.shape[0] == 0:
        return rows
    order = np.lexsort([rows[:, j] for j in reversed(list(idx))])
    rows = rows[order]
    _, first = np.unique(rows, axis=0, return_index=True)
    return rows[np.sort(first)]


def table_append(rows, buf):
    """`Table.commit` without an index (sys_fn_db.py:107-109): buffered rows are appended in order."""
    return np.concatenate([np.asarray(rows)] + [np.asarray(b).reshape(1, -1) for b in buf])


def table_upsert(rows, buf, idx):
    """`Table.commit` with an index (sys_fn_db.py:97-106): every existing row whose key appears in the buffer takes
    the buffered row's values, buffered rows with new keys are appended, then everything is sorted by key again.
    Several different buffered rows with the SAME existing key inside one commit are not pinned by the reference
    (its `.loc` assignment needs unique labels); the restatement lets the last one win."""
    rows = np.asarray(rows)
    idx = list(idx)
    buf = table_index(np.asarray(buf).reshape(-1, rows.shape[1]), idx)
    last = {tuple(r[idx]): r for r in buf}
    have = {tuple(r[idx]) for r in rows}
    upd = np.array([last.get(tuple(r[idx]), r) for r in rows]).reshape(-1, rows.shape[1])
    new = np.array([r for r in buf if tuple(r[idx]) not in have]).reshape(-1, rows.shape[1])
    out = np.concatenate([upd, new.astype(upd.dtype)]) if len(new) else upd
    return out[np.lexsort([out[:, j] for j in reversed(idx)])]


def table_group_stats(k, v):
    """`select k, min(v), avg(v), max(v), count(*) from T group by k order by k` — the grouped statistics of the db /
    1brc examples (examples/1brc/worker.kg:3-5 computes the same per chunk): -> (keys, min, mean, max, count)."""
    k = np.asarray(k)
    v = np.asarray(v, dtype=np.float64)
    keys, inv = np.unique(k, return_inverse=True)
    mn, mx, sm, ct = grouped_stats(inv.astype(np.int64), v, len(keys))
    return keys, mn, sm / ct, mx, ct
