"""Multi-GPU sharding of the hot path: one process per GPU, `torch.distributed` (NCCL over NVLink/NVSwitch)
for the plumbing.

A long vector is block-sharded: rank r holds the contiguous slice [r*n_local, (r+1)*n_local).  Elementwise
expressions need no exchange.  The primitives with a real exchange step:

  * reduce  `+/ */ |/ &/` : per-rank fused fold -> all_gather of `world` scalars (8 B each) -> the SAME
    fixed-order combine kernel on every rank (bit-identical results everywhere; all_reduce's association is
    NCCL's choice, so it is not used for float sums).
  * scan    `+\\ *\\ |\\ &\\` : per-rank fold -> all_gather of totals -> exclusive prefix over lower ranks = the
    shard's carried offset -> ONE local decoupled-look-back scan seeded with that offset (24 B/elem/rank).
  * grouped min/mean/max : per-rank partial G x 4 tables -> all_to_all of the per-owner table slices
    (station block s // B owned by rank s // B) -> merge -> all_gather of the final slices.

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
    if world() == 1:
        return part, None
    w = world()
    # one all_gather of the 16-byte pair {partial fold, shard total} (raw bytes: the two may differ in dtype)
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


def sharded_groupagg(keys, vals, n_groups):
    """Grouped (min, max, sum, count) over block-sharded rows.  Every rank returns the full G-entry tables."""
    w = world()
    mn, mx, sm, ct = ops.groupagg(keys, vals, n_groups)
    if w == 1:
        return mn, mx, sm, ct
    G = int(n_groups)
    B = (G + w - 1) // w  # stations [r*B, (r+1)*B) are owned by rank r
    dev = mn._t.device
    lib, stream = ops._lib_for(dev)
    # One exchange buffer for the four tables (all 8-byte lanes; float64 travels as its bit pattern):
    # send[r, j, :] = table j of this rank's partials for the stations rank r owns.
    i64 = torch.int64
    pad = torch.empty((4, w * B), dtype=i64, device=dev)
    if w * B > G:
        pad[0, G:] = torch.tensor(float("inf"), dtype=torch.float64).view(i64).item()
        pad[1, G:] = torch.tensor(float("-inf"), dtype=torch.float64).view(i64).item()
        pad[2:, G:] = 0
    for j, t in enumerate((mn._t, mx._t, sm._t, ct._t)):
        pad[j, :G].copy_(t.view(i64) if t.dtype != i64 else t)
    send = pad.view(4, w, B).permute(1, 0, 2).contiguous()
    recv = torch.empty_like(send)
    td.all_to_all_single(recv, send)  # recv[k] = rank k's partial tables for MY stations
    acc = recv[0]
    f64 = torch.float64
    for k in range(1, w):  # fixed rank order: identical results on every run
        sl = recv[k]
        _lib.check(lib.kgb_groupagg_merge(acc[0].data_ptr(), acc[1].data_ptr(), acc[2].data_ptr(), acc[3].data_ptr(),
                                          sl[0].data_ptr(), sl[1].data_ptr(), sl[2].data_ptr(), sl[3].data_ptr(), B,
                                          stream), "kgb_groupagg_merge")
    full = torch.empty(w * 4 * B, dtype=i64, device=dev)
    td.all_gather_into_tensor(full, acc.contiguous().view(-1))
    tabs = full.view(w, 4, B).permute(1, 0, 2).reshape(4, w * B)[:, :G].contiguous()
    return (DeviceArray(tabs[0].view(f64)), DeviceArray(tabs[1].view(f64)), DeviceArray(tabs[2].view(f64)),
            DeviceArray(tabs[3] if ct._t.dtype == i64 else tabs[3].view(ct._t.dtype)))


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
