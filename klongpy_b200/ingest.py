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


def report(text, device=None):
    """{name: 'min/mean/max'} with one decimal each (par.kg:5 `result`)."""
    names, mn, mean, mx, _ = aggregate(text, device)
    return {s: f"{a:.1f}/{b:.1f}/{c:.1f}" for s, a, b, c in zip(names, mn, mean, mx)}
