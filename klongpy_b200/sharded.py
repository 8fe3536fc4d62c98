"""Multi-GPU behind the backend seam: ONE interpreter, one process, k devices (SURVEY.md §8e process model).

KlongPy is single-process and constructs one backend instance per interpreter (`interpreter.py:237-249`), so a Klong
program can only use several GPUs if the backend's own array type spans them.  `B200BackendProvider(device="cuda:*")`
(or `"cuda:0,1,2,3"`) block-shards long 1-d vectors over the devices: a sharded vector is a `DeviceArray` whose node is a
`ShardNode` (one `DeviceArray` per device), so it passes every `isinstance(v, backend.np.ndarray)` check of the reference
(`compiler.py:56`), and `ops.run` — the one entry every eager operator and every compiled stage goes through — hands
programs with a sharded operand to `run()` below:

  * MAP (elementwise, fused trees)  : the same program on every device's shard, no exchange;
  * REDUCE `+/ */ |/ &/`             : per-device fused fold -> the k partials meet on device 0 -> the same fixed-order
                                      combine kernel as `dist.sharded_reduce` (8 bytes per device over NVLink);
  * SCAN `+\\ *\\ |\\ &\\`              : per-device fold of the scanned expression -> exclusive offsets on device 0 -> each
                                      device scans its shard seeded with its offset (`KGB_HAS_CARRY`), 24 B/elem/device;
  * grouped min / mean / max         : per-device partial tables -> device 0 -> `kgb_groupagg_fold` in device order.

Launches are asynchronous: the host queues device k's kernels and moves on to device k+1, so the shards run
concurrently.  The exchanges are 8 B .. 320 KB peer copies, so no NCCL communicator is needed inside one process (the
one-process-per-GPU flavour with NCCL collectives is `klongpy_b200.dist`).  Anything without a sharded kernel (sort,
indexing, a fused scan+reduce) touches `._t`, which gathers the vector onto device 0 — correct, single-GPU speed, and
announced by a `ShardGatherWarning`.
"""
import warnings

import numpy as np
import torch

from . import _lib
from .array import DeviceArray, _T2K

ACTIVE = False   # set once a sharded vector exists: ops.run then looks for sharded operands (zero cost otherwise)
SHARD_MIN = 1 << 22  # elements below which kg_asarray keeps a vector on one device


class ShardGatherWarning(UserWarning):
    """A sharded vector was gathered onto one device because the operation has no sharded kernel."""


class ShardNode:
    """Node of a DeviceArray that lives as contiguous blocks on several devices (block r = elements [lo_r, hi_r))."""
    __slots__ = ("shards", "shape", "kdt", "device", "uses", "args", "code")

    def __init__(self, shards):
        self.shards = list(shards)
        self.shape = (sum(s.size for s in self.shards),)
        self.kdt = _T2K[self.shards[0]._t.dtype]
        self.device = self.shards[0]._t.device
        self.uses = 1          # never inlined into a deferred program (klongpy_b200.lazy)
        self.args, self.code = [], ()


def make(shards):
    global ACTIVE
    ACTIVE = True
    return DeviceArray._deferred(ShardNode(shards))


def is_sharded(x):
    return isinstance(x, DeviceArray) and x._tensor is None and isinstance(x._node, ShardNode)


def shards_of(x):
    return x._node.shards


def gather(node):
    """All blocks on the first device, as one tensor (the fallback for operations without a sharded kernel)."""
    warnings.warn("a sharded vector is gathered onto one device: the operation has no multi-GPU kernel", ShardGatherWarning,
                  stacklevel=4)
    dev = node.device
    return torch.cat([s._t.to(dev) for s in node.shards])


def bounds(n, k):
    per = (n + k - 1) // k
    return [(min(r * per, n), min((r + 1) * per, n)) for r in range(k)]


def shard(a, devices):
    """Host array / one-device DeviceArray -> sharded vector over `devices` (contiguous blocks, ceil(n/k) each)."""
    if is_sharded(a):
        return a
    if isinstance(a, DeviceArray):
        t = a._t
        parts = [DeviceArray(t[lo:hi].to(d)) for (lo, hi), d in zip(bounds(t.numel(), len(devices)), devices)]
    else:
        from . import ops
        arr = np.ascontiguousarray(a)
        parts = [ops.asarray(arr[lo:hi], device=d) for (lo, hi), d in zip(bounds(arr.shape[0], len(devices)), devices)]
    return make(parts)


def to_numpy(x):
    return np.concatenate([s.to_numpy() for s in shards_of(x)])


# --------------------------------------------------------------------------------------------- program runner
def _layout(args):
    """(devices, shard sizes) of the first sharded operand; None if two sharded operands are partitioned differently."""
    sizes = devs = None
    for a in args:
        if is_sharded(a):
            sh = shards_of(a)
            s, d = [x.size for x in sh], [x._t.device for x in sh]
            if sizes is None:
                sizes, devs = s, d
            elif s != sizes or d != devs:
                return None
    return devs, sizes


def _local(a, r, dev, sizes):
    """Operand `a` as device r sees it."""
    if is_sharded(a):
        return shards_of(a)[r]
    if isinstance(a, DeviceArray):
        t = a._t
        if t.dim() == 0:
            return a if t.device == dev else DeviceArray(t.to(dev))      # replicate a 0-d value (8-byte peer copy)
        if t.dim() == 1 and t.numel() == sum(sizes):                       # a one-device vector of the same length
            lo = sum(sizes[:r])
            return DeviceArray(t[lo:lo + sizes[r]].to(dev))
        raise ValueError("operand shape does not match the sharded vector")
    return a


def _on_first(parts, dev):
    """k 0-d partials (one per device) -> one 1-d DeviceArray of length k on `dev`."""
    return DeviceArray(torch.stack([p._t.reshape(()).to(dev) for p in parts]))


def run(code, args, mode, red, code2, red2):
    """ops.run for programs with a sharded operand.  Returns NotImplemented when the program has no sharded form (the
    caller then proceeds on the gathered vector)."""
    from . import ops
    lay = _layout(args)
    if lay is None or mode == _lib.MODE_SCANRED:
        return NotImplemented
    devs, sizes = lay
    k = len(devs)
    try:
        local = [[_local(a, r, devs[r], sizes) for a in args] for r in range(k)]
    except ValueError:
        return NotImplemented
    if mode == _lib.MODE_MAP:
        if code2:
            return NotImplemented
        return make([ops.run(code, local[r], _no_defer=True) for r in range(k)])
    if any(s == 0 for s in sizes):
        return NotImplemented   # an empty block has no fold: tiny vectors are not worth sharding anyway
    if mode == _lib.MODE_REDUCE:
        if code2:  # two folds in one pass: handled by reduce2()
            return NotImplemented
        parts = [ops.run(code, local[r], mode=_lib.MODE_REDUCE, red=red, _no_defer=True) for r in range(k)]
        return ops.run([ops.OP_ARG, 0], [_on_first(parts, devs[0])], mode=_lib.MODE_REDUCE, red=red, _no_defer=True)
    if mode == _lib.MODE_SCAN:
        if code2:
            return NotImplemented
        totals = [ops.run(code, local[r], mode=_lib.MODE_REDUCE, red=red, _no_defer=True) for r in range(k)]
        # exclusive offsets: the inclusive scan of the k totals on device 0; device r (> 0) carries element r-1
        incl = ops.run([ops.OP_ARG, 0], [_on_first(totals, devs[0])], mode=_lib.MODE_SCAN, red=red, _no_defer=True)
        outs = []
        for r in range(k):
            if r == 0:
                outs.append(ops.run(code, local[0], mode=_lib.MODE_SCAN, red=red, _no_defer=True))
            else:
                off = DeviceArray(incl._t[r - 1].to(devs[r]))
                outs.append(ops.run(code, list(local[r]) + [off], mode=_lib.MODE_SCAN, red=red,
                                    code2=[ops.OP_ARG, len(local[r])], _no_defer=True))
        return make(outs)
    return NotImplemented


def reduce2(code, red, code2, red2, args):
    """Two folds of the same sharded operands in one pass per device (VWAP `(+/p*s)%+/s`)."""
    from . import ops
    lay = _layout(args)
    if lay is None:
        return NotImplemented
    devs, sizes = lay
    if any(s == 0 for s in sizes):
        return NotImplemented
    k = len(devs)
    pa, pb = [], []
    for r in range(k):
        a, b = ops.reduce2(code, red, code2, red2, [_local(x, r, devs[r], sizes) for x in args])
        pa.append(a)
        pb.append(b)
    return (ops.run([ops.OP_ARG, 0], [_on_first(pa, devs[0])], mode=_lib.MODE_REDUCE, red=red, _no_defer=True),
            ops.run([ops.OP_ARG, 0], [_on_first(pb, devs[0])], mode=_lib.MODE_REDUCE, red=red2, _no_defer=True))


def groupagg(keys, vals, n_groups):
    """Grouped (min, max, sum, count) of sharded rows: per-device partial tables, folded on the first device in device
    order (kgb_groupagg_fold) — the single-process form of dist.sharded_groupagg's all_gather exchange."""
    from . import ops
    lay = _layout([keys, vals])
    if lay is None or not (is_sharded(keys) and is_sharded(vals)):
        return NotImplemented
    devs, sizes = lay
    G, k = int(n_groups), len(devs)
    i64, f64 = torch.int64, torch.float64
    blocks, checks = [], []
    for r in range(k):
        block = torch.empty((4, G), dtype=i64, device=devs[r])
        tabs = (block[0].view(f64), block[1].view(f64), block[2].view(f64), block[3])
        # no wait for the error flag here: device r's kernel runs while device r + 1's is being enqueued
        _, chk = ops.groupagg(shards_of(keys)[r], shards_of(vals)[r], G, tables=tabs, accumulate=False, defer_check=True)
        blocks.append(block)
        checks.append(chk)
    allb = torch.stack([b.to(devs[0]) for b in blocks])   # k x 4 x G lanes on device 0 (320 KB each at 10 k stations)
    out = torch.empty((4, G), dtype=i64, device=devs[0])
    o = (out[0].view(f64), out[1].view(f64), out[2].view(f64), out[3])
    lib, stream = ops._lib_for(devs[0])
    _lib.check(lib.kgb_groupagg_fold(allb.data_ptr(), k, G, 4 * G, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(),
                                     o[3].data_ptr(), stream), "kgb_groupagg_fold")
    for chk in checks:
        chk()
    return tuple(DeviceArray(t) for t in o)
