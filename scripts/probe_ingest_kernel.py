"""kgb_brc_parse alone (for ncu): a 1e6-row synthetic file tiled `reps` times on the device."""
import ctypes
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from klongpy_b200 import _lib, ingest, ops  # noqa: E402
from bench import synthetic_1brc_block  # noqa: E402

dev = torch.device("cuda:0")
ops.set_default_device(dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
block, names, ids, tenths = synthetic_1brc_block()
text = torch.from_numpy(np.frombuffer(block, dtype=np.uint8).copy()).to(dev).repeat(reps)
n, rows = text.numel(), ids.size * reps
lib, stream = ops._lib_for(dev)
wsb = ctypes.c_size_t()
lib.kgb_brc_workspace_bytes(n, ctypes.byref(wsb))
ws = torch.empty(wsb.value, dtype=torch.uint8, device=dev)
o_ids = torch.empty(rows, dtype=torch.int64, device=dev)
o_t = torch.empty(rows, dtype=torch.float64, device=dev)
dn = torch.zeros(ingest.DICT_CAPACITY * 256, dtype=torch.uint8, device=dev)
dl = torch.empty(ingest.DICT_CAPACITY, dtype=torch.int32, device=dev)
nn, rr = ctypes.c_int64(), ctypes.c_int64()
for _ in range(iters):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    _lib.check(lib.kgb_brc_parse(text.data_ptr(), n, rows, o_ids.data_ptr(), o_t.data_ptr(), dn.data_ptr(), dl.data_ptr(),
                                 ingest.DICT_CAPACITY, ctypes.byref(rr), ctypes.byref(nn), ws.data_ptr(), wsb.value, stream), "parse")
    e.record()
    torch.cuda.synchronize()
    print(f"rows {rr.value} names {nn.value}: {s.elapsed_time(e):.3f} ms", flush=True)
