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
  * grade `<a >a` : sample sort — composite (key, position) splitters, one variable-size all_to_all of (key, position)
    pairs, a local stable sort, one all_to_all of positions back to the block partition (sharded_argsort).
  * group `=a` : the same sample sort + boundaries of each rank's sorted piece, merged across pieces (sharded_group).
  * range `?a` : the ranks' own answers concatenated in rank order, then `?` of that (sharded_range).
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
    # raw 8-byte lanes: the two values may differ in dtype.  ops.reduce2 writes both results side by side into one
    # 16-byte buffer: that buffer IS the message (no concatenation kernel in the dependent chain between the fold and
    # the scan); one transposing copy then yields both columns contiguous
    pt, tt = part._t, total._t
    if (pt.untyped_storage().data_ptr() == tt.untyped_storage().data_ptr() and pt.element_size() == 8 and
            tt.element_size() == 8 and tt.storage_offset() == pt.storage_offset() + 1):
        pair = torch.empty(0, dtype=torch.int64, device=pt.device).set_(pt.untyped_storage(), pt.storage_offset(), (2,))
    else:
        pair = torch.cat([pt.reshape(1).view(torch.int64), tt.reshape(1).view(torch.int64)])
    allp = torch.empty(2 * w, dtype=torch.int64, device=pair.device)
    td.all_gather_into_tensor(allp, pair)
    cols = allp.reshape(w, 2).t().contiguous()
    parts = DeviceArray(cols[0].view(pt.dtype))
    totals = DeviceArray(cols[1].view(tt.dtype))
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
    if w == 1:
        ops.groupagg(keys, vals, G, tables=tabs, accumulate=False)
        return tuple(DeviceArray(t) for t in tabs)
    # the aggregation is enqueued WITHOUT waiting for its error flag: the exchange below (host-side NCCL enqueue, two
    # small kernels) is issued while the aggregation kernel is still running; the flag is collected at the end
    _, check = ops.groupagg(keys, vals, G, tables=tabs, accumulate=False, defer_check=True)
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
    check()
    return tuple(DeviceArray(t) for t in o)


def _all_counts(counts_row):
    """all_gather of one row of `world` int64 counts per rank -> [world, world] host matrix (row r = rank r's row)."""
    w = world()
    mat = torch.empty(w * w, dtype=torch.int64, device=counts_row.device)
    td.all_gather_into_tensor(mat, counts_row.contiguous())
    return mat.reshape(w, w).cpu().numpy()


def _exchange(send, send_counts, recv_counts):
    """Variable-size all_to_all of a 1-d tensor (segments in destination-rank order)."""
    recv = torch.empty(int(sum(recv_counts)), dtype=send.dtype, device=send.device)
    td.all_to_all_single(recv, send.contiguous(), [int(c) for c in recv_counts], [int(c) for c in send_counts])
    return recv


def sharded_group(a, samples_per_rank=1024):
    """`=a` of a block-sharded vector (monads.py:182-204) in CSR form: (this rank's block of the
    GLOBAL order — global positions, grouped by ascending key, ascending within a group; partitioned like the input —,
    the group start offsets into the global order, replicated on every rank (G + 1 entries), G).

    The sample sort of `sharded_argsort` leaves every rank with a contiguous piece of the globally sorted sequence; the
    group boundaries inside a piece come from the local `=` of its keys, a piece's first group is merged with the
    previous piece's last one when they carry the same key (one all_gather of each piece's first / last key), and the
    boundary lists meet in one (padded) all_gather."""
    order, info = sharded_argsort(a, samples_per_rank=samples_per_rank, _group=True)
    return order, info[0], info[1]


def sharded_argsort(a, descending=False, samples_per_rank=1024, _group=False):
    """Grade of a block-sharded vector: this rank's block of the GLOBAL stable permutation (global
    positions; the blocks are partitioned like the input).  Sample sort:

      1. every rank contributes `samples_per_rank` (key, global position) pairs; the w-1 splitters are quantiles of the
         sorted samples — COMPOSITE (key, position) splitters, so heavy ties (10 k distinct keys) still balance and the
         stable order is respected across ranks;
      2. elements are bucketed by one fused compare program per splitter group, partitioned stably by bucket (a one-pass
         packed radix sort of the bucket ids) and exchanged with ONE variable-size all_to_all of keys and one of positions
         (16 B per element over NVLink);
      3. the received segments arrive in source-rank order = ascending position among equal keys, so ONE local stable sort by
         key finishes the job; a second all_to_all of positions only (8 B per element) restores the block partition.

    `descending` follows the language contract: the reversed stable ascending permutation (writer.py:97-108)."""
    w, r = world(), rank()
    if w == 1:
        if _group:
            o, st, g = ops.group(a)
            return o, (st, g)
        return ops.argsort(a, descending=descending)
    dev = a._t.device
    n_local = a.size
    sizes = torch.empty(w, dtype=torch.int64, device=dev)
    td.all_gather_into_tensor(sizes, torch.tensor([n_local], dtype=torch.int64, device=dev))
    sizes = sizes.cpu().numpy()
    lo, n_global = int(sizes[:r].sum()), int(sizes.sum())
    gidx = ops.iota(n_local, dev, start=lo)
    keys = a if a._t.dtype != torch.bool else ops.astype(a, np.int64)
    # ---- 1. splitters
    s = samples_per_rank
    if n_local:
        pick = ops.astype(ops.binary("div", ops.binary("mul", ops.iota(s, dev), n_local), s), np.int64)   # floor(j * n / s)
        sk, sp = ops.gather(keys, pick)._t, ops.gather(gidx, pick)._t
    else:
        sk = torch.empty(0, dtype=keys._t.dtype, device=dev)
        sp = torch.empty(0, dtype=torch.int64, device=dev)
    cnt = _all_counts(torch.tensor([sk.numel()] * w, dtype=torch.int64, device=dev))[:, 0]
    pad_k = torch.zeros(s, dtype=keys._t.dtype, device=dev)
    pad_p = torch.zeros(s, dtype=torch.int64, device=dev)
    pad_k[: sk.numel()], pad_p[: sp.numel()] = sk, sp
    all_k = torch.empty(w * s, dtype=keys._t.dtype, device=dev)
    all_p = torch.empty(w * s, dtype=torch.int64, device=dev)
    td.all_gather_into_tensor(all_k, pad_k)
    td.all_gather_into_tensor(all_p, pad_p)
    valid = torch.cat([torch.arange(s, device=dev) < int(c) for c in cnt])
    all_k, all_p = DeviceArray(all_k[valid]), DeviceArray(all_p[valid])
    order = ops.argsort(all_k)                       # samples are in ascending position order per rank: stable = composite
    m = all_k.size
    if m == 0:
        empty = DeviceArray(torch.empty(0, dtype=torch.int64, device=dev))
        return (empty, (DeviceArray(torch.zeros(1, dtype=torch.int64, device=dev)), 0)) if _group else empty
    qs = ops.asarray(np.minimum((np.arange(1, w) * m) // w, m - 1).astype(np.int64), device=dev)
    at = ops.gather(order, qs)
    split_k, split_p = ops.gather(all_k, at).to_numpy(), ops.gather(all_p, at).to_numpy()
    # ---- 2. bucket = number of splitters (k_s, p_s) <= (key, position), lexicographically
    bucket = None
    for j in range(w - 1):
        ks = split_k[j].item()
        ps = int(split_p[j])
        # (key > ks) | ((key == ks) & (pos >= ps))  as 0/1 int64:  gt + eq * ge
        code = [ops.OP_ARG, 0, ops.OP_ARG, 2, ops.OPC["gt"], ops.OP_ARG, 0, ops.OP_ARG, 2, ops.OPC["eq"],
                ops.OP_ARG, 1, ops.OP_ARG, 3, ops.OPC["lt"]] + ops.lit(1) + [ops.OPC["sub"], ops.OPC["neg"], ops.OPC["mul"], ops.OPC["add"]]
        term = ops.run(code, [keys, gidx, ks, ps], _no_defer=True) if n_local else ops.iota(0, dev)
        bucket = term if bucket is None else ops.binary("add", bucket, term)
    if n_local:
        if np.issubdtype(np.asarray(split_k).dtype, np.floating):
            # NaN keys sort last (np.argsort): comparisons with NaN are false, so they would land in bucket 0
            isn = ops.astype(ops.unary("isnan_b", keys), np.int64)
            bucket = ops.binary("add", ops.binary("mul", bucket, ops.binary("sub", 1, isn)), ops.binary("mul", isn, w - 1))
        part = ops.argsort(bucket)                   # stable partition by bucket
        send_k, send_p = ops.gather(keys, part)._t, ops.gather(gidx, part)._t
        counts = torch.stack([ops.reduce("add", ops.astype(ops.binary("eq_b", bucket, b), np.int64))._t for b in range(w)])
    else:
        send_k, send_p = keys._t, gidx._t
        counts = torch.zeros(w, dtype=torch.int64, device=dev)
    mat = _all_counts(counts)                        # mat[src, dst]
    recv_k = _exchange(send_k, mat[r, :], mat[:, r])
    recv_p = _exchange(send_p, mat[r, :], mat[:, r])
    # ---- 3. local stable sort by key, then back to the block partition
    held = mat.sum(axis=0)                           # elements each rank holds after the exchange
    my_off = int(held[:r].sum())
    group_info = None
    if _group:
        # boundaries of this piece (local `=`), shifted to global offsets; the piece's first group continues the previous
        # non-empty piece's last one when the keys agree (NaN keys all fall into one group, like the local kernel)
        if recv_k.numel():
            local, lstarts, lg = ops.group(DeviceArray(recv_k))
            chunk = ops.gather(DeviceArray(recv_p), local)._t
            first_k, last_k = recv_k[local._t[0]].reshape(1), recv_k[local._t[-1]].reshape(1)
            bounds = lstarts._t[:lg] + my_off
        else:
            chunk = recv_p
            first_k = last_k = torch.zeros(1, dtype=recv_k.dtype, device=dev)
            bounds = torch.empty(0, dtype=torch.int64, device=dev)
        ends = torch.empty(2 * w, dtype=recv_k.dtype, device=dev)
        td.all_gather_into_tensor(ends, torch.cat([first_k, last_k]))
        ends = ends.reshape(w, 2).cpu().numpy()
        prev = [q for q in range(r) if held[q] > 0]
        if bounds.numel() and prev:
            pk, fk = ends[prev[-1], 1], ends[r, 0]
            if pk == fk or (pk != pk and fk != fk):
                bounds = bounds[1:]
        bc = _all_counts(torch.tensor([bounds.numel()] * w, dtype=torch.int64, device=dev))[:, 0]
        cap = max(int(bc.max()), 1)
        padb = torch.zeros(cap, dtype=torch.int64, device=dev)
        padb[: bounds.numel()] = bounds
        allb = torch.empty(w * cap, dtype=torch.int64, device=dev)
        td.all_gather_into_tensor(allb, padb)
        validb = torch.cat([torch.arange(cap, device=dev) < int(c) for c in bc])
        starts = torch.cat([allb[validb], torch.tensor([n_global], dtype=torch.int64, device=dev)])
        group_info = (DeviceArray(starts), int(starts.numel()) - 1)
    elif recv_k.numel():
        local = ops.argsort(DeviceArray(recv_k))
        chunk = ops.gather(DeviceArray(recv_p), local)._t
    else:
        chunk = recv_p
    if descending:
        chunk = ops.reverse(DeviceArray(chunk))._t if chunk.numel() else chunk
        my_off = n_global - my_off - int(held[r])    # reversed global order: this rank's chunk sits mirrored
    blocks = np.concatenate([[0], np.cumsum(sizes)])
    send_counts = [max(0, min(my_off + int(held[r]), int(blocks[d + 1])) - max(my_off, int(blocks[d]))) for d in range(w)]
    if descending:
        # destination blocks are visited from the last rank down in the mirrored order; segments must still be laid out
        # in destination-rank order for the collective, which they are (offsets ascend with the destination)
        pass
    mat2 = _all_counts(torch.tensor(send_counts, dtype=torch.int64, device=dev))
    out = _exchange(chunk, mat2[r, :], mat2[:, r])
    if descending:
        # source ranks contribute in ascending rank order, but in the mirrored order higher ranks come first
        segs, o = [], 0
        for src in range(w):
            c = int(mat2[src, r])
            segs.append(out[o:o + c])
            o += c
        out = torch.cat(segs[::-1]) if segs else out
    if _group:
        return DeviceArray(out), group_info
    return DeviceArray(out)


def sharded_range(a):
    """`?a` of a block-sharded vector: the distinct values in order of first appearance, replicated on
    every rank.  The ranks' own answers, concatenated in rank order, list every candidate in ascending position of its first
    LOCAL appearance — so the global answer is simply `?` of that concatenation: one all_gather of the candidate counts, one
    of the (padded) candidates, one small local `?a`."""
    w = world()
    local = ops.unique_first(a)
    if w == 1:
        return local
    dev = a._t.device
    cnt = _all_counts(torch.tensor([local.size] * w, dtype=torch.int64, device=dev))[:, 0]
    cap = int(cnt.max())
    pad = torch.zeros(cap, dtype=local._t.dtype, device=dev)
    pad[: local.size] = local._t
    allc = torch.empty(w * cap, dtype=local._t.dtype, device=dev)
    td.all_gather_into_tensor(allc, pad)
    valid = torch.cat([torch.arange(cap, device=dev) < int(c) for c in cnt])
    return ops.unique_first(DeviceArray(allc[valid]))


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
