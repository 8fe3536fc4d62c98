"""Time the fused scan+reduce (max drawdown `|/(|\\ts)-ts`) and the running-sum variant at 1e8 float64."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import klongpy_b200  # noqa: E402
from klongpy_b200 import DeviceArray  # noqa: E402
be = klongpy_b200.B200BackendProvider(device="cuda:0")
n = 100_000_000
g = torch.Generator(device="cuda").manual_seed(1)
ts = DeviceArray(torch.rand(n, dtype=torch.float64, device="cuda", generator=g))
dd = ("reduce", "|", ("binop", "-", ("scan", "|", ("var", "_v0")), ("var", "_v0")))
sm = ("reduce", "|", ("binop", "-", ("scan", "+", ("var", "_v0")), ("var", "_v0")))
for name, ir in (("|/(|\\ts)-ts", dd), ("|/(+\\ts)-ts", sm)):
    fn, _ = be.compile_expr_ir(ir, ["ts"])
    r = fn(ts).item()
    h = ts.to_numpy()
    ref = np.max((np.maximum.accumulate(h) if "|\\" in name else np.cumsum(h)) - h)
    assert abs(r - ref) <= 1e-9 * abs(ref), (name, r, ref)
    for _ in range(3):
        fn(ts)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        fn(ts)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f"{name}: {ms:.3f} ms  {8 * n / ms / 1e6:.0f} GB/s  frac {8 * n / ms / 1e6 / 6546.2:.3f}", flush=True)
