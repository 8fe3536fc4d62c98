"""Multi-GPU sharding of the hot path: one process per GPU, `torch.distributed` (NCCL over NVLink/NVSwitch)
for the plumbing.

A long vector is block-sharded: rank r holds the contiguous slice [r*n_local, (r+1)*n_local).  Elementwise
expressions need no exchange.  The primitives with a real exchange step:

  * reduce  `+/ */ |/ &/` : per-rank fused fold -> all_gather of `world` scalars (8 B each) -> the SAME
    fixed-order combine kernel on every rank (bit-identical results everywhere; all_reduce's association is
    NCCL's choice, so it is not used for float sums).
  * scan    `+\\ *\\ |\\ &\\` : per-rank fold -> all_gather of totals -> exclusive prefix over lower ranks = the
    shard's carried offset -> ONE local decoupled-look-back scan seeded with that offset (24 B/elem/rank).
  * grouped min/mean/max : per-rank partial G x 4 tables -> hash-partitioned all_to_all (station s owned by rank
    s mod w) -> one rank-ordered fold kernel -> all_gather of the owners' results; small tables (10 k stations) take
    ONE all_gather + one fold kernel instead (sharded_groupagg).

  * find `a?b` : local compaction -> all_gather of (shard length, hit count) -> global positions (sharded_find).
  * autograd : the loss `+/` is a sharded reduce whose tape holds only the local fused node (sharded_loss_sum);
    vector gradients stay sharded, scalar-parameter gradients are summed rank-ordered (combine_param_grads).

Messages are tiny (8 B .. 320 KB): they are latency-, not bandwidth-bound, so each primitive does exactly one
or two collectives.
"""
import numpy as np
import torch
import torch.distributed as td

from . import _lib, ops
from .array import DeviceArray

_IDENT = {_lib.RED_ADD: 0, _lib.RED_MUL: 1}


def world():
    return td.get_world_size() if td.is_available() and td.is_initialized() else 1


def rank():
    return td.get_rank() if td.is_available() and td.is_initialized() else 0


def _gather_scalars(x):
    """all_gather of one 0-d device value per rank -> 1-d DeviceArray of length world (same on every rank)."""
    w = world()
    t = x._t.reshape(1)
    out = torch.empty(w, dtype=t.dtype, device=t.device)
    td.all_gather_into_tensor(out, t.contiguous())
    return DeviceArray(out)


def sharded_reduce(code, args, red):
    """Fold of an expression over a block-sharded vector.  Returns a 0-d DeviceArray replicated on all ranks."""
    red = ops.RED[red] if not isinstance(red, int) else red
    part = ops.run(code, args, mode=_lib.MODE_REDUCE, red=red)
    if world() == 1:
        return part
    allp = _gather_scalars(part)
    return ops.run([ops.OP_ARG, 0], [allp], mode=_lib.MODE_REDUCE, red=red)


def combine_partials(part, red):
    """All ranks' 0-d partial folds -> the global fold (same fixed-order combine on every rank)."""
    red = ops.RED[red] if not isinstance(red, int) else red
    if world() == 1:
        return part
    return ops.run([ops.OP_ARG, 0], [_gather_scalars(part)], mode=_lib.MODE_REDUCE, red=red)


def sharded_reduce_with_total(code, red, args, scan_code, scan_red):
    """One pass over the shard for BOTH the fold of `code` and the shard total of `scan_code` (what a following
    sharded_scan needs for its carried offset): returns (global fold replicated on every rank, this rank's exclusive
    scan offset or None on rank 0).  Saves the separate 8 B/elem reduce pass of `shard_offset`, so a reduce + scan
    step stays at 24 B/elem per rank for any number of ranks."""
    red = ops.RED[red] if not isinstance(red, int) else red
    scan_red = ops.RED[scan_red] if not isinstance(scan_red, int) else scan_red
    part, total = ops.reduce2(code, red, scan_code, scan_red, args)
    return exchange_fold_and_offset(part, total, red, scan_red)


def exchange_fold_and_offset(part, total, red, scan_red):
    """This rank's {partial fold, shard total} (0-d device values) -> (global fold, exclusive scan offset of this rank or
    None on rank 0) with ONE all_gather of 16 bytes per rank and the same fixed-order combine on every rank."""
    red = ops.RED[red] if not isinstance(red, int) else red
    scan_red = ops.RED[scan_red] if not isinstance(scan_red, int) else scan_red
    if world() == 1:
        return part, None
    w = world()
    # raw bytes: the two values may differ in dtype
    pair = torch.cat([part._t.reshape(1).view(torch.uint8), total._t.reshape(1).view(torch.uint8)])
    allp = torch.empty(16 * w, dtype=torch.uint8, device=pair.device)
    td.all_gather_into_tensor(allp, pair)
    allp = allp.reshape(w, 16)
    parts = DeviceArray(allp[:, :8].contiguous().view(part._t.dtype).reshape(w))
    totals = DeviceArray(allp[:, 8:].contiguous().view(total._t.dtype).reshape(w))
    fold = ops.run([ops.OP_ARG, 0], [parts], mode=_lib.MODE_REDUCE, red=red)
    r = rank()
    off = None if r == 0 else ops.run([ops.OP_ARG, 0], [DeviceArray(totals._t[:r])], mode=_lib.MODE_REDUCE, red=scan_red)
    return fold, off


def sharded_scan_with_offset(code, args, red, off):
    """Local scan seeded with a precomputed exclusive offset (from sharded_reduce_with_total)."""
    red = ops.RED[red] if not isinstance(red, int) else red
    if off is None:
        return ops.run(code, args, mode=_lib.MODE_SCAN, red=red)
    return ops.run(code, list(args) + [off], mode=_lib.MODE_SCAN, red=red, code2=[ops.OP_ARG, len(args)])


def shard_offset(code, args, red):
    """Exclusive fold over all lower ranks of the per-rank totals: the offset this shard's scan carries.
    Returns (offset 0-d DeviceArray or None on rank 0, totals)."""
    red = ops.RED[red] if not isinstance(red, int) else red
    total = ops.run(code, args, mode=_lib.MODE_REDUCE, red=red)
    totals = _gather_scalars(total)
    r = rank()
    if r == 0:
        return None, totals
    return ops.run([ops.OP_ARG, 0], [DeviceArray(totals._t[:r])], mode=_lib.MODE_REDUCE, red=red), totals


def sharded_scan(code, args, red):
    """Inclusive scan of an expression over a block-sharded vector; returns this rank's slice of the result."""
    red = ops.RED[red] if not isinstance(red, int) else red
    if world() == 1:
        return ops.run(code, args, mode=_lib.MODE_SCAN, red=red)
    off, _ = shard_offset(code, args, red)
    if off is None:
        return ops.run(code, args, mode=_lib.MODE_SCAN, red=red)
    seed = [ops.OP_ARG, len(args)]
    return ops.run(code, list(args) + [off], mode=_lib.MODE_SCAN, red=red, code2=seed)


ALLGATHER_LIMIT_BYTES = 64 << 20  # total bytes of all ranks' table blocks up to which one all_gather is the exchange


def sharded_groupagg(keys, vals, n_groups, exchange="auto"):
    """Grouped (min, max, sum, count) over block-sharded rows.  Every rank returns the full G-entry tables.

    The per-rank partial tables are one block of 8-byte lanes [4][G] (written by the aggregation kernel in place).
      * "alltoall" — the hash-partitioned exchange north_star names: station s is owned by rank hash(s) = s mod w;
        kgb_groupagg_to_owners packs the block per owner, ONE all_to_all moves the slices, ONE kgb_groupagg_fold merges
        the w slices of an owner in rank order, ONE all_gather returns the owners' results and
        kgb_groupagg_from_owners puts them back in station order.  Per rank it moves 2 blocks whatever w is.
      * "allgather" — for small tables (10 k stations: 320 KB per rank) the exchange is latency-, not bandwidth-bound:
        ONE all_gather of the blocks + ONE kgb_groupagg_fold over the w blocks in rank order (w blocks per rank moved,
        one collective instead of two).
      * "auto" — allgather while w * block stays under ALLGATHER_LIMIT_BYTES, else alltoall.
    Both fold in rank order, so results are bit-identical on every rank and every run."""
    w = world()
    G = int(n_groups)
    dev = keys._t.device
    i64, f64 = torch.int64, torch.float64
    block = torch.empty((4, G), dtype=i64, device=dev)
    tabs = (block[0].view(f64), block[1].view(f64), block[2].view(f64), block[3])
    ops.groupagg(keys, vals, G, tables=tabs, accumulate=False)
    if w == 1:
        return tuple(DeviceArray(t) for t in tabs)
    lib, stream = ops._lib_for(dev)
    if exchange == "auto":
        exchange = "allgather" if w * 32 * G <= ALLGATHER_LIMIT_BYTES else "alltoall"
    out = torch.empty((4, G), dtype=i64, device=dev)
    o = (out[0].view(f64), out[1].view(f64), out[2].view(f64), out[3])
    if exchange == "allgather":
        allb = torch.empty((w, 4, G), dtype=i64, device=dev)
        td.all_gather_into_tensor(allb.view(-1), block.view(-1))
        _lib.check(lib.kgb_groupagg_fold(allb.data_ptr(), w, G, 4 * G, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(),
                                         o[3].data_ptr(), stream), "kgb_groupagg_fold")
    elif exchange == "alltoall":
        B = (G + w - 1) // w
        send = torch.empty((w, 4, B), dtype=i64, device=dev)
        _lib.check(lib.kgb_groupagg_to_owners(block.data_ptr(), G, w, send.data_ptr(), stream), "kgb_groupagg_to_owners")
        recv = torch.empty_like(send)
        td.all_to_all_single(recv.view(-1), send.view(-1))  # recv[k] = rank k's partials for the stations I own
        mine = torch.empty((4, B), dtype=i64, device=dev)
        _lib.check(lib.kgb_groupagg_fold(recv.data_ptr(), w, B, 4 * B, mine[0].data_ptr(), mine[1].data_ptr(),
                                         mine[2].data_ptr(), mine[3].data_ptr(), stream), "kgb_groupagg_fold")
        full = torch.empty((w, 4, B), dtype=i64, device=dev)
        td.all_gather_into_tensor(full.view(-1), mine.view(-1))
        _lib.check(lib.kgb_groupagg_from_owners(full.data_ptr(), G, w, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(),
                                                o[3].data_ptr(), stream), "kgb_groupagg_from_owners")
    else:
        raise ValueError(f"sharded_groupagg: unknown exchange {exchange!r}")
    return tuple(DeviceArray(t) for t in o)


def sharded_find(a, needle, n_global=None):
    """`a?b` over a block-sharded vector: this rank's hits as GLOBAL positions (ascending), plus every rank's hit
    count (1-d host array), so the caller knows where its slice sits in the concatenated answer.  One all_gather
    of `world` counts; no data moves."""
    pos = ops.find_equal(a, needle)
    w = world()
    if w == 1:
        return pos, np.array([pos.size], dtype=np.int64)
    n_local = a.size
    sizes = torch.empty(2 * w, dtype=torch.int64, device=a._t.device)
    mine = torch.tensor([n_local, pos.size], dtype=torch.int64, device=a._t.device)
    td.all_gather_into_tensor(sizes, mine)
    sizes = sizes.reshape(w, 2).cpu().numpy()
    start = int(sizes[:rank(), 0].sum())  # block start from the actual shard lengths (ragged last block included)
    if start:
        pos = ops.binary("add", pos, start)
    return pos, sizes[:, 1].copy()


def sharded_loss_sum(code, args):
    """Differentiable `+/` of an expression over block-sharded operands (autograd row): the scalar loss
    is replicated on every rank, the tape holds only this rank's fused reduce node, so gradients of sharded vectors
    stay sharded and need no exchange.  Gradients of replicated (scalar) parameters are per-rank partial sums:
    finish them with `combine_param_grads`."""
    part = ops.run(code, args, mode=_lib.MODE_REDUCE, red=_lib.RED_ADD)
    if world() == 1:
        return part
    with torch.no_grad():
        allp = _gather_scalars(DeviceArray(part._t.detach()))
        total = ops.run([ops.OP_ARG, 0], [allp], mode=_lib.MODE_REDUCE, red=_lib.RED_ADD)
        rest = total._t - part._t.detach()
    # value = the fixed-order global sum (bit-identical on every rank); d value / d part = 1
    return DeviceArray(part._t - part._t.detach() + (rest + part._t.detach()))


def combine_param_grads(grads):
    """Per-rank partial gradients of replicated scalar parameters -> their global sums, in ONE all_gather and with the
    same rank-ordered summation on every rank."""
    w = world()
    if w == 1:
        return list(grads)
    dev = grads[0]._t.device
    mine = torch.stack([g._t.reshape(()).to(torch.float64) for g in grads])
    allg = torch.empty(w * len(grads), dtype=torch.float64, device=dev)
    td.all_gather_into_tensor(allg, mine.contiguous())
    allg = allg.reshape(w, len(grads))
    out = allg[0].clone()
    for r in range(1, w):
        out = out + allg[r]
    return [DeviceArray(out[i].reshape(())) for i in range(len(grads))]


def shard_bounds(n_global, r=None, w=None):
    """Contiguous block partition of a global index range."""
    r = rank() if r is None else r
    w = world() if w is None else w
    per = (n_global + w - 1) // w
    lo = min(r * per, n_global)
    return lo, min(lo + per, n_global)
