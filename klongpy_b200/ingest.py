"""1brc text ingest on the device.

The reference's example reads the file line by line in Python — `location, measurement = line.split(";")`, stations
as strings, `np.asarray(temps, dtype=float)` (examples/1brc/par.py:57-88) — and the Klong worker then groups the rows
by station NAME (`=s`, worker.kg:5).  Here the whole text buffer goes to the GPU once: `parse` returns a dense station
id and a float64 temperature per row plus the station dictionary, `aggregate` feeds them to the grouped
min / mean / max kernel and returns the report of par.kg:4-6 (stations in ascending name order).  The host only
ever touches the dictionary (one name per station), never the rows.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib, ops
from .array import DeviceArray

DICT_CAPACITY = 1 << 17  # distinct station names per call (the table behind it has 2^18 slots)


def _as_device_text(text, device):
    if isinstance(text, DeviceArray):
        text = text._t
    if isinstance(text, torch.Tensor):
        t = text if text.dtype == torch.uint8 else text.view(torch.uint8)
        return t.to(device).contiguous().reshape(-1)
    if isinstance(text, str):
        text = text.encode("utf-8")
    if isinstance(text, (bytes, bytearray, memoryview)):
        arr = np.frombuffer(bytes(text), dtype=np.uint8)
    else:
        arr = np.ascontiguousarray(np.asarray(text, dtype=np.uint8)).reshape(-1)
    if arr.size and arr[-1] != 10:
        arr = np.concatenate([arr, np.array([10], dtype=np.uint8)])  # the last line may lack its newline
    return torch.from_numpy(arr.copy()).to(device)


NAME_BYTES = 256  # one dictionary row (include/kgb200.h: kgb_brc_parse)


def _parse_arrival(text, device, fetch_names=True):
    """-> (ids in first-arrival order (torch int64), temps (torch float64), (names list[bytes] by arrival id, the same
    as a fixed-width numpy bytes array or None)).  fetch_names=False stops after the two C-ABI calls."""
    t = _as_device_text(text, device)
    n = t.numel()
    lib, stream = ops._lib_for(device)
    wsb = ctypes.c_size_t()
    _lib.check(lib.kgb_brc_workspace_bytes(n, ctypes.byref(wsb)), "kgb_brc_workspace_bytes")
    ws = torch.empty(max(wsb.value, 64), dtype=torch.uint8, device=device)
    # one pass over the text: the outputs are sized from an estimate of the row count (three 4 MB windows, + 2 %); a
    # text whose windows were not representative reports how many rows it really holds and is parsed once more
    est = ctypes.c_int64()
    _lib.check(lib.kgb_brc_estimate(t.data_ptr(), n, ctypes.byref(est), ws.data_ptr(), wsb.value, stream), "kgb_brc_estimate")
    cap = est.value
    dnames = torch.zeros(DICT_CAPACITY * NAME_BYTES, dtype=torch.uint8, device=device)
    dlen = torch.empty(DICT_CAPACITY, dtype=torch.int32, device=device)
    rows, n_names = ctypes.c_int64(), ctypes.c_int64()
    while True:
        ids = torch.empty(cap, dtype=torch.int64, device=device)
        temps = torch.empty(cap, dtype=torch.float64, device=device)
        _lib.check(lib.kgb_brc_parse(t.data_ptr(), n, cap, ids.data_ptr(), temps.data_ptr(), dnames.data_ptr(), dlen.data_ptr(),
                                     DICT_CAPACITY, ctypes.byref(rows), ctypes.byref(n_names), ws.data_ptr(), wsb.value, stream),
                   "kgb_brc_parse")
        if rows.value <= cap:
            break
        cap = rows.value
        dnames.zero_()
    ids, temps = ids[:rows.value], temps[:rows.value]
    d = n_names.value
    if d == 0 or not fetch_names:
        return ids, temps, ([], None)
    # the dictionary comes back as one [D, 256] block of zero-padded names: two small copies, no per-row work
    block = dnames[:d * NAME_BYTES].cpu().numpy().reshape(d, NAME_BYTES)
    lens = dlen[:d].cpu().numpy()
    L = int(lens.max())
    fixed = None
    if (block[:, :L] == 0).sum() == int((L - lens).sum()):   # no NUL inside a name: fixed-width bytes strip exactly the padding
        fixed = np.ascontiguousarray(block[:, :L]).view(f"S{L}").reshape(d)
        raw = fixed.tolist()
    else:
        raw = [bytes(block[i, :lens[i]]) for i in range(d)]
    return ids, temps, (raw, fixed)


def _name_order(names):
    """par.kg:4 `s@<s`: ascending names (bytewise).  -> (order: sorted position -> arrival id, rank: the inverse)"""
    raw, fixed = names
    d = len(raw)
    if fixed is not None:
        order = np.argsort(fixed, kind="stable").astype(np.int64)   # unsigned bytewise, as Python compares bytes
    else:
        order = np.asarray(sorted(range(d), key=raw.__getitem__), dtype=np.int64)
    rank = np.empty(d, dtype=np.int64)
    rank[order] = np.arange(d, dtype=np.int64)
    return order, rank


def parse(text, device=None):
    """-> (ids int64 DeviceArray [rows], temps float64 DeviceArray [rows], names list[str] sorted ascending).
    ids index `names`.  `text` is bytes / str / a uint8 array or tensor (host or device) of complete lines."""
    device = torch.device(device) if device is not None else ops.default_device()
    ids, temps, names = _parse_arrival(text, device)
    raw = names[0]
    if not raw:
        return DeviceArray(ids), DeviceArray(temps), []
    order, rank = _name_order(names)
    # arrival-order ids -> name-order ids: a D-entry lookup gathered over the rows (8 B/row in, 8 B/row out)
    lut = ops.asarray(rank, device=device)
    ids = ops.gather(lut, DeviceArray(ids))
    return ids, DeviceArray(temps), [raw[i].decode("utf-8") for i in order]


def aggregate(text, device=None):
    """-> (names sorted, min, mean, max as host float64 arrays, count int64): examples/1brc end to end.
    The rows are aggregated under their arrival-order ids and only the D results are put in name order."""
    device = torch.device(device) if device is not None else ops.default_device()
    ids, temps, names = _parse_arrival(text, device)
    raw = names[0]
    if not raw:
        z = np.zeros(0)
        return [], z, z, z, np.zeros(0, dtype=np.int64)
    order, _ = _name_order(names)
    mn, mx, sm, ct = ops.groupagg(DeviceArray(ids), DeviceArray(temps), len(raw))
    ctn = ct.to_numpy()[order]
    return ([raw[i].decode("utf-8") for i in order], mn.to_numpy()[order], sm.to_numpy()[order] / ctn,
            mx.to_numpy()[order], ctn)


def chunk_bounds(data, parts):
    """Line-aligned byte ranges of `data` (bytes-like), one per part: the chunking of examples/1brc/par.py:14-49
    (`get_file_chunks`: a chunk end falls back to the previous line start; an empty chunk grows to the next line end)
    restated for a buffer.  -> [(start, end)] covering the buffer; parts with nothing left are (n, n)."""
    mv = memoryview(data).cast("B") if not isinstance(data, (bytes, bytearray)) else data
    n = len(mv)
    size = max(1, n // max(int(parts), 1))
    out, start = [], 0
    for r in range(int(parts)):
        if start >= n:
            out.append((n, n))
            continue
        end = n if r == parts - 1 else min(n, start + size)
        while end > start and end < n and mv[end - 1] != 10:
            end -= 1
        if end == start:  # one line longer than the chunk: take it whole
            nl = bytes(mv[start:]).find(b"\n")
            end = n if nl < 0 else start + nl + 1
        out.append((start, end))
        start = end
    return out


def sharded_aggregate(text_chunk, device=None):
    """`aggregate` over a file split across the ranks of torch.distributed (one line-aligned chunk per rank, as
    par.py's worker processes get theirs): every rank parses and aggregates its own chunk on its GPU, then the
    D-entry per-station tables — not the rows — are exchanged and merged by station NAME in rank order with the
    worker's rule (worker.kg:4 `merge`: min&, max|, sum+, count+).  Every rank returns the full result."""
    import torch.distributed as td
    device = torch.device(device) if device is not None else ops.default_device()
    mine = _local_tables(text_chunk, device)
    if td.is_available() and td.is_initialized() and td.get_world_size() > 1:
        parts = [None] * td.get_world_size()
        td.all_gather_object(parts, mine)
    else:
        parts = [mine]
    return _merge_tables(parts)


def _local_tables(text_chunk, device):
    """One chunk of complete lines -> (names by arrival id (bytes), min, max, sum, count) as host arrays."""
    ids, temps, names = _parse_arrival(text_chunk, device)
    raw = names[0]
    if not raw:
        return ([], np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0, dtype=np.int64))
    mn, mx, sm, ct = ops.groupagg(DeviceArray(ids), DeviceArray(temps), len(raw))
    return (raw, mn.to_numpy(), mx.to_numpy(), sm.to_numpy(), ct.to_numpy())


def _merge_tables(parts):
    """Per-chunk station tables -> one, by station NAME, in the order given (worker.kg:4 `merge`: min&, max|, sum+,
    count+).  -> (names ascending, min, mean, max, count)"""
    index, cols = {}, ([], [], [], [])
    for rnames, rmn, rmx, rsm, rct in parts:  # fixed rank order: identical sums on every rank and every run
        for j, s in enumerate(rnames):
            k = index.get(s)
            if k is None:
                index[s] = len(cols[0])
                for c, v in zip(cols, (rmn[j], rmx[j], rsm[j], rct[j])):
                    c.append(v)
            else:
                cols[0][k] = min(cols[0][k], rmn[j])
                cols[1][k] = max(cols[1][k], rmx[j])
                cols[2][k] = cols[2][k] + rsm[j]
                cols[3][k] = cols[3][k] + rct[j]
    order = sorted(index, key=lambda b: b)
    sel = np.asarray([index[s] for s in order], dtype=np.int64)
    mn, mx, sm = (np.asarray(c, dtype=np.float64)[sel] if len(sel) else np.zeros(0) for c in cols[:3])
    ct = np.asarray(cols[3], dtype=np.int64)[sel] if len(sel) else np.zeros(0, dtype=np.int64)
    return [s.decode("utf-8") for s in order], mn, sm / np.maximum(ct, 1), mx, ct


_file_buffers = {}
_readers = None


def _read_parallel(fd, offset, mv, block=16 << 20, threads=16):
    """Fill `mv` from the file at `offset` with positional reads issued from several threads (a single read() copies
    out of the page cache at a few GB/s; pread releases the GIL, so the copies run side by side).  -> bytes read; short
    only at the end of the file."""
    global _readers
    n = len(mv)
    if n <= block:
        return os.preadv(fd, [mv], offset)
    if _readers is None:
        from concurrent.futures import ThreadPoolExecutor
        _readers = ThreadPoolExecutor(max_workers=threads)

    def one(lo):
        done, want = 0, min(block, n - lo)
        while done < want:
            r = os.preadv(fd, [mv[lo + done:lo + want]], offset + lo + done)
            if r <= 0:
                break
            done += r
        return done
    total = 0
    for lo, got in zip(range(0, n, block), _readers.map(one, range(0, n, block))):
        total += got
        if got < min(block, n - lo):
            break                                          # end of file inside this block: later blocks read nothing
    return total


def aggregate_file(path, device=None, chunk_bytes=256 << 20):
    """examples/1brc end to end from a FILE: the measurements file is read in line-aligned chunks of about
    `chunk_bytes` straight into pinned host memory, each chunk's host->device copy runs on its own stream while the
    previous chunk is parsed and aggregated, and the per-chunk station tables are merged by name (the reference
    splits the file the same way across worker processes, par.py:14-49, 92-104).  -> like `aggregate`."""
    device = torch.device(device) if device is not None else ops.default_device()
    cuda = device.type == "cuda"
    chunk_bytes = max(int(chunk_bytes), 1 << 16)
    cap = chunk_bytes + (1 << 16)                          # room for the carried partial line
    key = (str(device), cap)
    if key not in _file_buffers:                           # pinning 2 x 256 MB costs tens of milliseconds: keep them
        _file_buffers.clear()
        _file_buffers[key] = ([torch.empty(cap, dtype=torch.uint8, pin_memory=cuda) for _ in range(2)],
                              [torch.empty(cap, dtype=torch.uint8, device=device) for _ in range(2)] if cuda else None)
    host, devb = _file_buffers[key]
    copy_stream = torch.cuda.Stream(device) if cuda else None
    parts, pending, carry, k, pos = [], None, b"", 0, 0

    def finish(p):
        slot, m, ev = p
        if cuda:
            torch.cuda.current_stream(device).wait_event(ev)
            parts.append(_local_tables(devb[slot][:m], device))
        else:
            parts.append(_local_tables(host[slot][:m].clone(), device))

    with open(path, "rb", buffering=0) as f:
        while True:
            slot = k & 1
            hv = host[slot].numpy()
            c = len(carry)
            if c > cap - chunk_bytes:
                raise _lib.KgbError("aggregate_file: a line longer than 65536 bytes")
            hv[:c] = np.frombuffer(carry, dtype=np.uint8)
            got = _read_parallel(f.fileno(), pos, memoryview(hv)[c:c + chunk_bytes])
            pos += got
            m = c + (got or 0)
            if m == 0:
                break
            eof = not got
            if eof:                                        # the last line may lack its newline
                if hv[m - 1] != 10:
                    hv[m] = 10
                    m += 1
                cut = m
            else:
                lo = max(0, m - (1 << 16))
                nl = np.flatnonzero(hv[lo:m] == 10)
                if nl.size:
                    cut = lo + int(nl[-1]) + 1
                elif got < chunk_bytes:
                    cut = 0                                # a short read without a newline: keep it all for the next round
                else:
                    raise _lib.KgbError("aggregate_file: a line longer than 65536 bytes")
            carry = bytes(hv[cut:m])
            ev = None
            if cuda:
                with torch.cuda.stream(copy_stream):
                    devb[slot][:cut].copy_(host[slot][:cut], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
            if pending is not None:
                finish(pending)                            # parse chunk k-1 while chunk k is on the wire
                if cuda:
                    copy_stream.wait_stream(torch.cuda.current_stream(device))   # its buffers are free for chunk k+1
            pending = (slot, cut, ev) if cut else None
            k += 1
            if eof:
                break
    if pending is not None:
        finish(pending)
    if not parts:
        z = np.zeros(0)
        return [], z, z, z, np.zeros(0, dtype=np.int64)
    return _merge_tables(parts)


def report(text, device=None):
    """{name: 'min/mean/max'} with one decimal each (par.kg:5 `result`)."""
    names, mn, mean, mx, _ = aggregate(text, device)
    return {s: f"{a:.1f}/{b:.1f}/{c:.1f}" for s, a, b, c in zip(names, mn, mean, mx)}
