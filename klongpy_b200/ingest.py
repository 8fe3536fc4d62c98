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


def parse(text, device=None):
    """-> (ids int64 DeviceArray [rows], temps float64 DeviceArray [rows], names list[str] sorted ascending).
    ids index `names`.  `text` is bytes / str / a uint8 array or tensor (host or device) of complete lines."""
    device = torch.device(device) if device is not None else ops.default_device()
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
    doff = torch.empty(DICT_CAPACITY, dtype=torch.int64, device=device)
    dlen = torch.empty(DICT_CAPACITY, dtype=torch.int32, device=device)
    n_names = ctypes.c_int64()
    _lib.check(lib.kgb_brc_parse(t.data_ptr(), n, r, ids.data_ptr(), temps.data_ptr(), doff.data_ptr(), dlen.data_ptr(),
                                 DICT_CAPACITY, ctypes.byref(n_names), ws.data_ptr(), wsb.value, stream), "kgb_brc_parse")
    d = n_names.value
    if d == 0:
        return DeviceArray(ids), DeviceArray(temps), []
    # the dictionary: D names of <= 255 bytes, fetched as one padded [D, L] block (a gather on the device, one copy)
    off = doff[:d]
    ln = dlen[:d].to(torch.int64)
    L = int(ln.max().item())
    cols = torch.arange(L, device=device, dtype=torch.int64)
    idx = (off[:, None] + cols[None, :]).clamp_(max=n - 1)
    block = t[idx.reshape(-1)].reshape(d, L).cpu().numpy()
    lens = ln.cpu().numpy()
    raw = [bytes(block[i, :lens[i]]) for i in range(d)]
    order = sorted(range(d), key=lambda i: raw[i])           # par.kg:4 `s@<s`: ascending names
    rank = np.empty(d, dtype=np.int64)
    rank[np.asarray(order, dtype=np.int64)] = np.arange(d, dtype=np.int64)
    # arrival-order ids -> name-order ids: a D-entry lookup gathered over the rows (8 B/row in, 8 B/row out)
    lut = ops.asarray(rank, device=device)
    ids = ops.gather(lut, DeviceArray(ids))
    return ids, DeviceArray(temps), [raw[i].decode("utf-8") for i in order]


def aggregate(text, device=None):
    """-> (names sorted, min, mean, max as host float64 arrays, count int64): examples/1brc end to end."""
    ids, temps, names = parse(text, device)
    if not names:
        z = np.zeros(0)
        return [], z, z, z, np.zeros(0, dtype=np.int64)
    mn, mx, sm, ct = ops.groupagg(ids, temps, len(names))
    ctn = ct.to_numpy()
    return names, mn.to_numpy(), sm.to_numpy() / ctn, mx.to_numpy(), ctn


def report(text, device=None):
    """{name: 'min/mean/max'} with one decimal each (par.kg:5 `result`)."""
    names, mn, mean, mx, _ = aggregate(text, device)
    return {s: f"{a:.1f}/{b:.1f}/{c:.1f}" for s, a, b, c in zip(names, mn, mean, mx)}
