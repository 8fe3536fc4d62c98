"""DeviceTable — the reference's columnar Table with device-resident columns.

The reference keeps a table as a pandas DataFrame (klongpy/db/sys_fn_db.py:18-127) and answers grouped queries by
handing that frame to DuckDB (:129-147), so a column that lives on the GPU makes a host round trip before anything
is grouped or aggregated.  Here the table is a set of equally long 1-d device arrays: `T?"col"` hands the column to
the device verbs as is, an index is a stable multi-column radix sort, an indexed insert is one sort + a handful of
fused elementwise kernels over (existing rows ++ buffered rows), and `group_stats` — the
`select k, min(v), avg(v), max(v), count(*) … group by k order by k` of the db / 1brc examples — runs on the
grouped-aggregation kernel.  There is no SQL engine here: `.db` stays the reference's (DuckDB) business.

Klong programs load it the way they load the reference's module:

    .py("klongpy_b200.table")      :" instead of .py(""klongpy.db"") "
    T::.table(d); .schema(T); .insert(T;[4 5 6]); .index(T;["a"]); .rindex(T); T?"a"; .gstats(T;"k";"v")

Semantics follow the reference and are pinned by tests/golden/golden_table.json (recorded from the reference's own
Table class).  Two places where the reference's result is arbitrary or undefined are fixed by contract: rows with
equal index keys keep their insertion order (pandas' single-column `sort_index` is an unstable quicksort), and when
one commit buffers several different rows for the same existing key the last one wins.
"""
import numpy as np
import torch

from . import ops
from .array import DeviceArray

try:
    from klongpy.core import KLONG_UNDEFINED, KlongException
except Exception:  # pragma: no cover - GPU box without klongpy
    KLONG_UNDEFINED = None

    class KlongException(Exception):
        pass


class KlongDbException(KlongException):
    """Same name and shape as klongpy/db/sys_fn_db.py:12-15."""

    def __init__(self, x, e):
        self.x = x
        super().__init__(e)


def _concat(a, b):
    """a ++ b for 1-d device arrays (two D2D copies), promoting like NumPy."""
    d = np.result_type(a.dtype, b.dtype)
    a, b = (ops.astype(x, d) if x.dtype != d else x for x in (a, b))
    out = torch.empty(a.size + b.size, dtype=a._t.dtype, device=a._t.device)
    out[:a.size].copy_(a._t.reshape(-1))
    out[a.size:].copy_(b._t.reshape(-1))
    return DeviceArray(out)


def _lexsort(cols, order=None):
    """Stable row order by `cols` (first = major key): LSD over columns, one radix grade per column."""
    for c in reversed(cols):
        k = c if order is None else ops.gather(c, order)
        o = ops.argsort(k)
        order = o if order is None else ops.gather(order, o)
    return order


def _changes(cols, order):
    """bool[n-1]: sorted row i+1 differs from sorted row i in any of `cols` (one fused kernel per column)."""
    flag = None
    for c in cols:
        s = ops.gather(c, order)
        f = ops.binary("ne_b", DeviceArray(s._t[1:]), DeviceArray(s._t[:-1]))
        flag = f if flag is None else ops.binary("max", flag, f)
    return flag


class DeviceTable(dict):
    """Drop-in for `Table` (sys_fn_db.py:18-127) over device columns.  `data`: {name: 1-d array}; host arrays are
    uploaded once, device arrays are adopted without a copy (columns are never mutated in place)."""

    def __init__(self, data, columns=None, device=None):
        super().__init__()
        if not isinstance(data, dict):
            raise ValueError("Input must be a dictionary of columns.")
        self.columns = list(columns) if columns is not None else list(data.keys())
        self.device = torch.device(device) if device is not None else None
        self._cols = {}
        for k in self.columns:
            self._cols[k] = self._adopt(data[k])
        self._check_lengths()
        self.idx_cols = None
        self.buffer = []

    # ---------------------------------------------------------------- columns
    def _adopt(self, v):
        if isinstance(v, DeviceArray):
            if self.device is None:
                self.device = v._t.device
            return v if v.ndim == 1 else v.flatten()
        a = np.asarray(v)
        if a.dtype == object or a.dtype.kind in "US":
            raise KlongDbException(v, "device tables hold numeric columns only")
        if a.size == 0 and a.dtype.kind not in "if":
            a = a.astype(np.float64)
        return ops.asarray(np.ascontiguousarray(a.reshape(-1)), device=self.device or ops.default_device())

    def _check_lengths(self):
        n = {c.size for c in self._cols.values()}
        if len(n) > 1:
            raise ValueError("All arrays must be of the same length")  # pandas' message for the same mistake

    def __getitem__(self, x):
        return self.get(x)

    def __setitem__(self, x, y):
        return self.set(x, y)

    def __contains__(self, x):
        raise NotImplementedError()

    def get(self, x):
        self.commit()
        v = self._cols.get(x)
        return KLONG_UNDEFINED if v is None else v

    def set(self, x, y):
        self.commit()
        v = self._adopt(y)
        if self._cols and v.size != len(self):
            raise ValueError(f"Length of values ({v.size}) does not match length of index ({len(self)})")
        self._cols[x] = v
        if x not in self.columns:
            self.columns.append(x)

    def schema(self):
        return np.array(self.columns, dtype=object)

    def __len__(self):
        self.commit()
        return next(iter(self._cols.values())).size if self._cols else 0

    # ---------------------------------------------------------------- index
    def has_index(self):
        return self.idx_cols is not None

    def set_index(self, idx_cols):
        """sys_fn_db.py:55-69: rows sorted by the index columns; rows repeating an earlier row in EVERY column dropped."""
        self.commit()
        idx_cols = list(idx_cols)
        n = len(self)
        if n > 1:
            key = [self._cols[c] for c in idx_cols]
            rest = [self._cols[c] for c in self.columns if c not in idx_cols]
            full = _lexsort(key + rest)
            new = _changes(key + rest, full)                                   # sorted row i+1 differs from row i
            if not bool(ops.all_(new)):
                keep = _concat(ops.full(1, 1, new.dtype, self.device), new)
                pos = ops.gather(full, ops.nonzero(keep))                       # first occurrences, key-major order
                pos = ops.argsort(pos, return_sorted=True)[1]                   # ... back in insertion order
                order = ops.gather(pos, _lexsort([ops.gather(c, pos) for c in key]))
            else:
                order = _lexsort(key) if rest else full
            for c in self.columns:
                self._cols[c] = ops.gather(self._cols[c], order)
        self.idx_cols = idx_cols

    def reset_index(self):
        self.commit()
        self.idx_cols = None

    # ---------------------------------------------------------------- inserts
    def insert(self, y):
        self.buffer.append(np.asarray(y.to_numpy() if isinstance(y, DeviceArray) else y).reshape(1, -1))

    def insertb(self, y):
        y = np.asarray(y.to_numpy() if isinstance(y, DeviceArray) else y)
        self.buffer.append(y.reshape(-1, len(self.columns)))   # one block, not one Python object per row

    def commit(self):
        """sys_fn_db.py:93-110.  Without an index: append.  With one: existing rows whose key is buffered take the
        buffered row's values, rows with new keys are added, key order is restored — one stable sort of
        (existing ++ buffered) keys and elementwise kernels over the sorted order; the rows never leave the device."""
        if not self.buffer:
            return
        buf = np.concatenate(self.buffer)
        self.buffer = []
        n_old = next(iter(self._cols.values())).size
        if self.has_index():
            ki = [self.columns.index(c) for c in self.idx_cols]
            # the reference sorts the buffer by key and drops rows that repeat an earlier row first (:98-99)
            rest = [j for j in range(buf.shape[1]) if j not in ki]
            o = np.lexsort([buf[:, j] for j in reversed(ki + rest)])
            s = buf[o]
            first = np.concatenate(([True], (s[1:] != s[:-1]).any(axis=1)))
            buf = buf[np.sort(o[first])]                                        # first occurrences, insertion order
            buf = buf[np.lexsort([buf[:, j] for j in reversed(ki)])]            # ... stably by key
        both = {c: _concat(self._cols[c], self._adopt(np.ascontiguousarray(buf[:, j])))
                for j, c in enumerate(self.columns)}
        if self.has_index():
            key = [both[c] for c in self.idx_cols]
            n = n_old + buf.shape[0]
            order = _lexsort(key)                                               # old rows precede buffered rows per key
            chg = _changes(key, order)
            head = _concat(ops.full(1, 1, chg.dtype, self.device), chg)
            starts = ops.nonzero(head)                                          # first sorted slot of every key
            gid = ops.binary("sub", ops.scan("add", ops.astype(head, np.int64)), 1)
            ends = _concat(DeviceArray(starts._t[1:]), ops.full(1, n, np.int64, self.device))
            last = ops.gather(order, ops.binary("sub", ops.gather(ends, gid), 1))      # row closing the slot's key run
            frst = ops.gather(order, ops.gather(starts, gid))                          # row opening it
            is_old = ops.binary("lt_b", order, n_old)
            upd = ops.binary("min", is_old, ops.binary("ge_b", last, n_old))           # old row whose key is buffered
            src = ops.where3(upd, last, order)
            keep = ops.binary("max", is_old, ops.binary("ge_b", frst, n_old))          # old rows + rows of new keys
            src = ops.gather(src, ops.nonzero(keep))
            both = {c: ops.gather(v, src) for c, v in both.items()}
        self._cols = both

    # ---------------------------------------------------------------- queries
    def rows(self):
        """Host row matrix — what the reference's `db("select * from T")` returns (`df.values.squeeze()`, :143-144)."""
        self.commit()
        cols = [self._cols[c].to_numpy() for c in self.columns]
        return np.stack(cols, axis=1).squeeze() if cols else np.zeros(0)

    def group_stats(self, key, val, dense_limit=1 << 22):
        """`select key, min(val), avg(val), max(val), count(*) from T group by key order by key` on the device.
        -> (keys, min, mean, max, count) device arrays.  Integer keys in [0, dense_limit) go straight into the
        grouped-aggregation kernel as station ids; anything else is dictionary-encoded by one radix sort first."""
        self.commit()
        k, v = self._cols[key], self._cols[val]
        n = k.size
        if n == 0:
            e = ops.asarray(np.zeros(0), device=self.device)
            return k, e, e, e, ops.asarray(np.zeros(0, dtype=np.int64), device=self.device)
        if k._t.dtype == torch.int64:
            arg = [ops.OP_ARG, 0]                                               # min and max in ONE pass over the keys,
            lo, hi = ops.reduce2(arg, "min", arg, "max", [k])                   # one 16-byte read-back
            lo, hi = (int(x) for x in torch.stack((lo._t, hi._t)).tolist())
        elif k.dtype.kind == "i":
            lo, hi = int(ops.reduce("min", k).item()), int(ops.reduce("max", k).item())
        else:
            lo, hi = -1, 0
        if 0 <= lo and hi < dense_limit:
            mn, mx, sm, ct = ops.groupagg(k, v, hi + 1)
            if bool(ops.all_(ct)):                                              # every id in [0, hi] occurs: nothing to drop
                keys = ops.iota(hi + 1, self.device)
            else:
                keys = ops.nonzero(ct)
                mn, mx, sm, ct = (ops.gather(t, keys) for t in (mn, mx, sm, ct))
            if keys.dtype != k.dtype:
                keys = ops.astype(keys, k.dtype)
        else:
            order, starts, g = ops.group(k)
            sizes = ops.binary("sub", DeviceArray(starts._t[1:]), DeviceArray(starts._t[:-1]))
            ids = ops.expand(sizes)                                             # dense id of every sorted row
            mn, mx, sm, ct = ops.groupagg(ids, ops.gather(v, order), g)
            keys = ops.gather(k, ops.gather(order, DeviceArray(starts._t[:-1])))
        return keys, mn, ops.binary("div", sm, ct), mx, ct

    def __str__(self):
        n = len(self)
        idx = self.idx_cols or []
        header = " ".join(f"{k}{'*' if k in idx else ''}" for k in self.columns)
        if n == 0:
            return header
        head = np.stack([self._cols[c]._t[:10].cpu().numpy() for c in self.columns], axis=1)
        s = header + "\n" + "\n".join(" ".join(str(x) for x in r) for r in head.tolist())
        if n > 10:
            s += f"\n...\nrows={n}\n"
        return s


# ------------------------------------------------------------------------------------- Klong system functions
def _is_list(x):
    return isinstance(x, (np.ndarray, list, DeviceArray))


def eval_sys_fn_create_table(x):
    """

        .table(x)                                         [Table-Create]

        Device counterpart of sys_fn_db.py:153-185: "x" is a list of [name column] pairs.

    """
    if isinstance(x, DeviceArray):
        raise KlongDbException(x, "tables must be created from an array of column data map.")
    if _is_list(x):
        return DeviceTable({k: v for k, v in x}, columns=[k for k, _ in x])
    raise KlongDbException(x, "tables must be created from an array of column data map or a Pandas DataFrame.")


def eval_sys_fn_index(x, y):
    """

        .index(x, y)                                [Create-Table-Index]

        sys_fn_db.py:188-211, same checks and messages.

    """
    if not isinstance(x, DeviceTable):
        raise KlongDbException(x, "An index may only be created on a table.")
    if x.has_index():
        raise KlongDbException(x, "Table already has an index.")
    if not _is_list(y):
        raise KlongDbException(x, "An index must be a list of column names")
    for q in y:
        if q not in x.columns:
            raise KlongDbException(x, f"An index column {q} not found in table")
    x.set_index(list(y))
    return x.idx_cols or []


def eval_sys_fn_reset_index(x):
    """

        .rindex(x)                                  [Reset-Table-Index]

        sys_fn_db.py:214-228.

    """
    if not isinstance(x, DeviceTable):
        raise KlongDbException(x, "An index may only be created on a table.")
    if x.has_index():
        x.reset_index()
        return 1
    return 0


def eval_sys_fn_schema(x):
    """

        .schema(x)                                        [Table-Schema]

        sys_fn_db.py:231-246 (tables only; there is no device database object).

    """
    x = getattr(x, "a", x) if not isinstance(x, DeviceTable) else x
    if not isinstance(x, DeviceTable):
        raise KlongDbException(x, "A schema is available only for a table or database.")
    return x.schema()


def eval_sys_fn_insert_table(x, y):
    """

        .insert(x, y)                                     [Table-Insert]

        sys_fn_db.py:249-283: one row, or an array of rows.

    """
    if not isinstance(x, DeviceTable):
        raise KlongDbException(x, "Inserts must be applied to a table")
    if not _is_list(y):
        raise KlongDbException(x, "Values to insert must be a list")
    y = np.asarray(y.to_numpy() if isinstance(y, DeviceArray) else y)
    batch = y.ndim > 1
    y_cols = len(y[0]) if batch else len(y)
    if y_cols != len(x.columns):
        raise KlongDbException(x, f"Expected {len(x.columns)} values, received {len(y)}")
    if batch:
        x.insertb(y)
    else:
        x.insert(y)
    return x


def eval_sys_fn_group_stats(x, y, z):
    """

        .gstats(x, y, z)                                  [Table-Group-Stats]

        Grouped statistics of value column "z" by key column "y" of table "x", keys ascending:
        [keys min mean max count] — the query the db / 1brc examples send to DuckDB, on the device.

    """
    if not isinstance(x, DeviceTable):
        raise KlongDbException(x, "Grouped statistics need a table")
    for q in (y, z):
        if q not in x.columns:
            raise KlongDbException(x, f"Column {q} not found in table")
    out = np.empty(5, dtype=object)
    out[:] = list(x.group_stats(y, z))
    return out


def create_system_functions_db():
    """Same discovery rule as sys_fn_db.py:300-313: the Klong name is read from each docstring."""
    def _get_name(s):
        i = s.index(".")
        return s[i: i + s[i:].index("(")]

    return {_get_name(fn.__doc__): fn for name, fn in globals().items() if name.startswith("eval_sys_")}


klongpy_exports = create_system_functions_db()
