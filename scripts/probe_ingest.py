"""Throughput of the 1brc ingest: a 1e6-row synthetic file tiled to ~1.4 GB on the device."""
import os
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
from klongpy_b200 import ingest, ops  # noqa: E402

dev = torch.device("cuda:0")
ops.set_default_device(dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
t0 = time.perf_counter()
from bench import synthetic_1brc_block  # noqa: E402
block, names, ids, tenths = synthetic_1brc_block()
print(f"host text build {time.perf_counter() - t0:.1f} s, {len(block) / 1e6:.1f} MB per 1e6 rows", flush=True)
text = torch.from_numpy(np.frombuffer(block, dtype=np.uint8).copy()).to(dev).repeat(reps)
n = text.numel()
rows = ids.size * reps


def timeit(fn, k=3):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(k):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / k


i2, t2, nm = ingest.parse(text)
assert len(nm) == len(names) and i2.size == rows
ref_ids = torch.from_numpy(ids).to(dev).repeat(reps)
ref_t = torch.from_numpy(tenths / 10.0).to(dev).repeat(reps)
assert nm == sorted(names, key=lambda s: s.encode()) and torch.equal(i2.tensor, ref_ids) and torch.equal(t2.tensor, ref_t)
del i2, t2, ref_ids, ref_t
# the two C-ABI calls on their own (device time of everything they enqueue; each ends with a synchronising read)
import ctypes
from klongpy_b200 import _lib
lib, stream = ops._lib_for(dev)
wsb = ctypes.c_size_t()
lib.kgb_brc_workspace_bytes(n, ctypes.byref(wsb))
ws = torch.empty(wsb.value, dtype=torch.uint8, device=dev)
r = ctypes.c_int64()
o_ids = torch.empty(rows, dtype=torch.int64, device=dev)
o_t = torch.empty(rows, dtype=torch.float64, device=dev)
dn = torch.zeros(ingest.DICT_CAPACITY * 256, dtype=torch.uint8, device=dev)
dl = torch.empty(ingest.DICT_CAPACITY, dtype=torch.int32, device=dev)
nn = ctypes.c_int64()
scan = lambda: _lib.check(lib.kgb_brc_scan(text.data_ptr(), n, ctypes.byref(r), ws.data_ptr(), wsb.value, stream), "scan")
rr = ctypes.c_int64()
prs = lambda: _lib.check(lib.kgb_brc_parse(text.data_ptr(), n, rows, o_ids.data_ptr(), o_t.data_ptr(), dn.data_ptr(), dl.data_ptr(),
                                           ingest.DICT_CAPACITY, ctypes.byref(rr), ctypes.byref(nn), ws.data_ptr(), wsb.value, stream), "parse")
est = lambda: _lib.check(lib.kgb_brc_estimate(text.data_ptr(), n, ctypes.byref(r), ws.data_ptr(), wsb.value, stream), "estimate")
ms_s = timeit(scan, 5)
ms_e = timeit(est, 5)
ms_p = timeit(prs, 5)
assert rr.value == rows and nn.value == len(names), (rr.value, nn.value)
alg = n + 16 * rows
print(f"kgb_brc_scan (exact count) {ms_s:.3f} ms ({n / ms_s / 1e6:.0f} GB/s of text)   kgb_brc_estimate {ms_e:.3f} ms   kgb_brc_parse {ms_p:.3f} ms   "
      f"estimate + parse {ms_e + ms_p:.3f} ms = {alg / (ms_e + ms_p) / 1e6:.0f} GB/s algorithmic ({alg / rows:.1f} B/row: text + 16 B out)", flush=True)
ms = timeit(lambda: ingest.parse(text))
print(f"parse     rows={rows} bytes={n}: {ms:.2f} ms  {rows / ms / 1e6:.2f} Grows/s  {n / ms / 1e6:.0f} GB/s of text", flush=True)
ms = timeit(lambda: ingest.aggregate(text))
print(f"aggregate rows={rows}: {ms:.2f} ms  {rows / ms / 1e6:.2f} Grows/s", flush=True)
# from a FILE: line-aligned chunks through pinned memory, H2D of chunk k+1 under the parse of chunk k
import tempfile
path = os.path.join(tempfile.gettempdir(), "kgb200_probe_measurements.txt")
with open(path, "wb") as f:
    for _ in range(reps):
        f.write(block)
fsize = os.path.getsize(path)
for cb in (64 << 20, 256 << 20):
    res = ingest.aggregate_file(path, chunk_bytes=cb)   # warm: page cache, pinned buffers
    assert len(res[0]) == len(names) and int(res[4].sum()) == rows
    t0 = time.perf_counter()
    k = 3
    for _ in range(k):
        ingest.aggregate_file(path, chunk_bytes=cb)
    s = (time.perf_counter() - t0) / k
    print(f"aggregate_file {fsize / 1e9:.2f} GB, {cb >> 20} MB chunks: {s * 1e3:.0f} ms  {rows / s / 1e9:.2f} Grows/s  {fsize / s / 1e9:.1f} GB/s of file", flush=True)
os.remove(path)
