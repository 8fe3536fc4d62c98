"""1brc text ingest on the device.

The reference's example reads the file line by line in Python — `location, measurement = line.split(";")`, stations
as strings, `np.asarray(temps, dtype=float)` (examples/1brc/par.py:57-88) — and the Klong worker then groups the rows
by station NAME (`=s`, worker.kg:5).  Here the whole text buffer goes to the GPU once: `parse` returns a dense station
id and a float64 temperature per row plus the station dictionary, `aggregate` feeds them to the grouped
min / mean / max kernel and returns the report of par.kg:4-6 (stations in ascending name order).  The host only
ever touches the dictionary (one name per station), never the rows.
"""
import ctypes

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
    rows = ctypes.c_int64()
    _lib.check(lib.kgb_brc_scan(t.data_ptr(), n, ctypes.byref(rows), ws.data_ptr(), wsb.value, stream), "kgb_brc_scan")
    r = rows.value
    ids = torch.empty(r, dtype=torch.int64, device=device)
    temps = torch.empty(r, dtype=torch.float64, device=device)
    dnames = torch.zeros(DICT_CAPACITY * NAME_BYTES, dtype=torch.uint8, device=device)
    dlen = torch.empty(DICT_CAPACITY, dtype=torch.int32, device=device)
    n_names = ctypes.c_int64()
    _lib.check(lib.kgb_brc_parse(t.data_ptr(), n, r, ids.data_ptr(), temps.data_ptr(), dnames.data_ptr(), dlen.data_ptr(),
                                 DICT_CAPACITY, ctypes.byref(n_names), ws.data_ptr(), wsb.value, stream), "kgb_brc_parse")
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
    ids, temps, names = _parse_arrival(text_chunk, device)
    raw = names[0]
    if raw:
        mn, mx, sm, ct = ops.groupagg(DeviceArray(ids), DeviceArray(temps), len(raw))
        mine = (raw, mn.to_numpy(), mx.to_numpy(), sm.to_numpy(), ct.to_numpy())
    else:
        mine = ([], np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0, dtype=np.int64))
    if td.is_available() and td.is_initialized() and td.get_world_size() > 1:
        parts = [None] * td.get_world_size()
        td.all_gather_object(parts, mine)
    else:
        parts = [mine]
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


def report(text, device=None):
    """{name: 'min/mean/max'} with one decimal each (par.kg:5 `result`)."""
    names, mn, mean, mx, _ = aggregate(text, device)
    return {s: f"{a:.1f}/{b:.1f}/{c:.1f}" for s, a, b, c in zip(names, mn, mean, mx)}
