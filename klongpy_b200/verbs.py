"""Op-table overrides for the verbs that bypass the backend seam.

`= ? ! & ~ % + ^` (monads) and `? @ , :+` (dyads) call the *global* NumPy module or Python loops directly
(`bknp.unique/where/arange/repeat/...`, SURVEY.md §8b "leaks": monads.py:52,79,202-203,256,275-281,316,442,
dyads.py:176-181,336-342,563-585).  A provider cannot intercept those, so after constructing an interpreter
on the b200 backend `install(klong)` replaces the affected entries of `klong._vm` / `klong._vd` (plain dicts
built at interpreter.py:252-253) with device-aware versions.  Every override handles ONLY device arrays and
defers to the reference implementation for everything else (strings, dicts, host object arrays).
"""
import os
import warnings

import numpy as np
import torch as _torch

from . import ops
from .array import DeviceArray


class HostRoundTripWarning(UserWarning):
    """A structural verb (`:= :- :_ :^ :# $ :$`) moved device operands to the host and back (see install())."""


def _is_dev(x):
    return isinstance(x, DeviceArray)


def _is_num(x):
    return isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, bool)


def group_indices(a):
    """`=a` on a device array: groups in ascending key order, indices ascending inside each group —
    the contract of monads.py:182-204 (np.unique + one np.where per group) in one sort + one compaction."""
    if a.size == 0:
        return a
    order, starts, g = ops.group(a.flatten() if a.ndim != 1 else a)
    st = starts.to_numpy()
    sizes = np.diff(st)
    if g > 0 and (sizes == sizes[0]).all():
        return order.reshape(g, int(sizes[0]))  # numpy would build a 2-d int64 array here too
    out = np.empty(g, dtype=object)
    t = order._t
    for i in range(g):
        out[i] = DeviceArray(t[int(st[i]):int(st[i + 1])])
    return out


def install(klong):
    """Replace the leaking verbs of `klong` (a KlongInterpreter on the b200 backend).  Returns klong."""
    backend = klong._backend
    shim = backend.np
    dev = shim.device
    vm, vd = klong._vm, klong._vd
    ref_m = dict(vm)
    ref_d = dict(vd)

    # ---------------------------------------------------------------- monads
    def enumerate_(a):  # !a  (monads.py:38-52)
        if not backend.is_integer(a):
            raise RuntimeError(f"enumerate: invalid type error: {a}")
        return ops.iota(int(a), dev)

    def expand_where(a):  # &a  (monads.py:55-79)
        if _is_dev(a):
            if a.ndim == 0:
                a = a.reshape(1)
            return ops.expand(a)
        if _is_num(a):
            return ops.full(int(a), 0, np.int64, dev)
        if isinstance(a, (list, np.ndarray)) and not (isinstance(a, np.ndarray) and a.dtype == object):
            return ops.expand(ops.asarray(np.asarray(a, dtype=np.int64), device=dev))
        return ref_m["&"](a)

    def groupby(a):  # =a  (monads.py:182-204)
        if _is_dev(a) and a.dtype.kind in "ifb":
            return group_indices(a)
        return ref_m["="](a)

    def range_(a):  # ?a  (monads.py:260-295)
        if _is_dev(a) and a.ndim == 1 and a.dtype.kind in "ifb":
            return ops.unique_first(a)
        return ref_m["?"](a)

    def not_(a):  # ~a  (monads.py:240-257): 1 for 0, else 0, as int64
        if _is_dev(a) and a.size > 0:
            return ops.run([ops.OP_ARG, 0, ops.OPC["not_b"], ops.OPC["to_i64"]], [a])
        return ref_m["~"](a)

    def reciprocal(a):  # %a  (monads.py:298-316)
        if _is_dev(a) and a.ndim > 0:
            return ops.unary("recip", a)
        return ref_m["%"](a)

    def transpose(a):  # +a  (monads.py:430-442)
        if _is_dev(a):
            return shim.transpose(a) if a.ndim > 1 else a
        return ref_m["+"](a)

    def shape(a):  # ^a  (monads.py:372-405): metadata only, no D2H of the payload
        if _is_dev(a):
            if a.ndim == 0 or (a.ndim == 1 and a.size == 0):
                return 0  # `^[]` is 0 (monads.py:372-405: the empty list is an atom); `^[[]]` is [1 0]
            return np.asarray(a.shape)
        return ref_m["^"](a)

    def reverse(a):  # |a  (monads.py:319-340)
        if _is_dev(a):
            return ops.reverse(a) if a.ndim > 0 else a
        return ref_m["|"](a)

    vm.update({"!": enumerate_, "&": expand_where, "=": groupby, "?": range_, "~": not_, "%": reciprocal,
               "+": transpose, "^": shape, "|": reverse})

    # ---------------------------------------------------------------- dyads
    def find(a, b):  # a?b  (dyads.py:307-342)
        if _is_dev(a) and (_is_num(b) or (_is_dev(b) and b.ndim == 0)):
            return ops.find_equal(a, b)
        if _is_dev(a) and _is_dev(b) and b.ndim >= 1 and a.ndim >= 1:
            # list b (dyads.py:340-341): positions i with kg_equal(a[i], b) — for two device arrays kg_equal is exact
            # array equality (base.py:372-374).  The reference walks `a` in Python; here one compare + row fold + compaction.
            if a.ndim != b.ndim + 1 or a.shape[1:] != b.shape or a.shape[0] == 0:
                return DeviceArray(a._t.new_empty(0, dtype=_torch.int64))   # no element of `a` has b's shape: no match
            rows = a.shape[0]
            eq = ops.binary("eq_b", a.reshape(rows, -1), b.reshape(1, -1))  # [rows, cols] bool
            if eq.shape[1] == 0:
                return ops.iota(rows, dev)
            flags = ops.reduce_axis0([ops.OP_ARG, 0], [DeviceArray(eq._t.T.contiguous())], ops.RED["&"])
            return ops.nonzero(flags)
        return ref_d["?"](a, b)

    def at_index(a, b):  # a@b  (dyads.py:140-191)
        if _is_dev(a):
            if _is_dev(b):
                g = ops.grade_source(b)
                if g is not None and (g[0] is a or (g[0]._tensor is not None and a._tensor is not None and
                                                    g[0]._tensor.data_ptr() == a._tensor.data_ptr() and g[0].shape == a.shape
                                                    and g[0].dtype == a.dtype)):
                    return ops.sort_values(a, descending=g[1])  # a@<a: sorted keys from the last radix pass, no gather
                return a[b] if b.ndim == 0 else (ops.gather(a, b) if b.size else DeviceArray(a._t[:0]))
            if isinstance(b, (int, np.integer)) and not isinstance(b, bool):
                return a[int(b)]
            if isinstance(b, (list, np.ndarray)):
                bb = np.asarray(b)
                if bb.size == 0:
                    return DeviceArray(a._t[:0])
                if bb.dtype.kind in "iu":
                    return ops.gather(a, ops.asarray(bb.astype(np.int64), device=dev))
        elif _is_dev(b) and isinstance(a, (np.ndarray, str, list)):
            b = b.to_numpy() if b.ndim > 0 else b.item()
        return ref_d["@"](a, b)

    def join(a, b):  # a,b  (dyads.py:512-585)
        if _is_dev(a) or _is_dev(b):
            da = a if _is_dev(a) else None
            db = b if _is_dev(b) else None
            if da is not None and db is not None and da.ndim >= 1 and db.ndim >= 1:
                if da.shape[0] == 0:
                    return db
                if da.ndim == db.ndim and da.shape[1:] == db.shape[1:]:
                    return shim.concatenate((da, db))
            elif da is not None and da.ndim == 1 and (_is_num(b) or (db is not None and db.ndim == 0)):
                return shim.concatenate((da, _atom_array(b, da)))
            elif db is not None and db.ndim == 1 and (_is_num(a) or (da is not None and da.ndim == 0)):
                return shim.concatenate((_atom_array(a, db), db))
            elif da is not None and db is not None and da.ndim == 0 and db.ndim == 0:
                return shim.concatenate((da.reshape(1), db.reshape(1)))
        return ref_d[","](a, b)

    def _atom_array(x, like):
        if _is_dev(x):
            return x.reshape(1)
        d = np.result_type(like.dtype, np.float64 if isinstance(x, (float, np.floating)) else np.int64)
        return ops.full(1, x, d, dev)

    def rotate(a, b):  # a:+b  (dyads.py:897-923): np.roll
        if _is_dev(b) and b.ndim >= 1 and b.shape[0] > 0:
            a = int(a)
            n = b.shape[0]
            s = a % n
            if s == 0:
                return b
            idx = ops.run([ops.OP_ARG, 0, ops.OP_ARG, 1, ops.OPC["add"], ops.OP_ARG, 2, ops.OPC["fmod"]],
                          [ops.iota(n, dev), n - s, n])
            return ops.gather(b, idx)
        return ref_d[":+"](a, b)

    # ---------------------------------------------------------------- structural verbs: host round trip
    # `:= :- :_ :^ :#` (amend, amend-in-depth, cut, reshape, split) and monadic `,` build their results with global
    # NumPy calls that need real ndarrays (`numpy.put`, `bknp.array_split`, item assignment on the shape list:
    # dyads.py:24-135, 806-895, 926-973).  They restructure small arrays and are not on the data-parallel path, so device
    # operands make one round trip: to host, through the REFERENCE implementation, numeric results back to the device.
    def _to_host(x):
        if _is_dev(x):
            return x.to_numpy() if x.ndim > 0 else x.item()
        if isinstance(x, np.ndarray) and x.dtype == object:
            out = np.empty(len(x), dtype=object)
            for i, y in enumerate(x):
                out[i] = _to_host(y)
            return out
        if isinstance(x, list):
            return [_to_host(y) for y in x]
        return x

    def _via_host(fn, sym):
        def wrapped(a, b):
            if not (_has_dev(a) or _has_dev(b)):
                return fn(a, b)
            # explicit, never silent: a HostRoundTripWarning names the verb (once per verb), KGB200_STRICT=1 raises
            if os.environ.get("KGB200_STRICT", "0") not in ("", "0"):
                raise NotImplementedError(f"Klong verb `{sym}` has no device kernel (KGB200_STRICT=1 forbids the host round trip)")
            warnings.warn(f"Klong verb `{sym}` restructures its operands on the host: device operands make one round trip "
                          "(structural verb, not on the data-parallel path)", HostRoundTripWarning, stacklevel=2)
            r = fn(_to_host(a), _to_host(b))
            return backend.kg_asarray(r) if isinstance(r, (np.ndarray, list)) else r
        return wrapped

    def _has_dev(x):
        if _is_dev(x):
            return True
        if isinstance(x, np.ndarray) and x.dtype == object or isinstance(x, list):
            return any(_has_dev(y) for y in x)
        return False

    # the reference implementations of these verbs call `backend.np.*` internally, so the host twin must be the op
    # table of an interpreter on the reference's own NumPy backend (one shared instance)
    host_vd = _host_twin()._vd
    for sym in (":=", ":-", ":_", ":^", ":#", "$", ":$"):  # `$` / `:$` format to / from strings: host data either way
        if sym in host_vd:
            vd[sym] = _via_host(host_vd[sym], sym)

    vd.update({"?": find, "@": at_index, ",": join, ":+": rotate})
    klong._b200_installed = True
    return klong


_HOST_TWIN = None
BUILDING_HOST_TWIN = False


def _host_twin():
    """A KlongInterpreter on the reference's NumPy backend whose op table serves the host-round-trip verbs."""
    global _HOST_TWIN, BUILDING_HOST_TWIN
    if _HOST_TWIN is None:
        from klongpy import KlongInterpreter
        BUILDING_HOST_TWIN = True  # lets test harnesses that re-route every interpreter to 'b200' leave this one alone
        try:
            _HOST_TWIN = KlongInterpreter(backend="numpy")
        finally:
            BUILDING_HOST_TWIN = False
    return _HOST_TWIN


def grouped_stats(keys, vals, n_groups):
    """Grouped min/mean/max — the `stats`/`collect`/`merge` pipeline of examples/1brc/worker.kg:3-5 as one
    kernel pass.  Returns device arrays (min, max, sum, count) ordered by ascending station id."""
    return ops.groupagg(keys, vals, n_groups)
