"""End-to-end streamed step (host -> reduce + scan -> host) for a few chunk sizes / depths."""
import sys
import torch
sys.path.insert(0, ".")
from klongpy_b200 import ops, stream
from klongpy_b200.ops import OPC, OP_ARG, lit
dev = torch.device("cuda:0")
ops.set_default_device(dev)
n = 1_000_000_000
hx = torch.empty(n, dtype=torch.float64).pin_memory()
hx.copy_(torch.randint(0, 256, (n,), dtype=torch.int64).to(torch.float64) / 256)
hs = torch.empty(n, dtype=torch.float64).pin_memory()
RED = [OP_ARG, 0] + lit(2) + [OPC["mul"]] + lit(1) + [OPC["add"]] + lit(2) + [OPC["pow"]]
for chunk, depth in ((1 << 24, 3), (1 << 26, 3), (1 << 27, 3), (1 << 26, 4), (1 << 25, 6)):
    st = stream.HostStream(dev, chunk=chunk, depth=depth)
    st.reduce_and_scan(hx, RED, "+", [OP_ARG, 0], "+", hs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        st.reduce_and_scan(hx, RED, "+", [OP_ARG, 0], "+", hs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"chunk={chunk} depth={depth}: {ms:.1f} ms/step  {2 * n / ms / 1e6:.2f} Gelem/s  {8 * n / ms / 1e6:.1f} GB/s per direction", flush=True)
    del st
    torch.cuda.empty_cache()
