"""Time the kernels that fold with float64 max: |/x, |\\x and |/(|\\x)-x at 1e8."""
import sys
import torch
sys.path.insert(0, ".")
import klongpy_b200 as kb  # noqa: E402
from klongpy_b200 import ops, DeviceArray  # noqa: E402

dev = torch.device("cuda:0")
ops.set_default_device(dev)
be = kb.B200BackendProvider(device="cuda:0")
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
x = DeviceArray(torch.randn(n, dtype=torch.float64, device=dev))


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


dd, _ = be.compile_expr_ir(("reduce", "|", ("binop", "-", ("scan", "|", ("var", "_v0")), ("var", "_v0"))), ["x"])
ref = float((torch.cummax(x.tensor, 0).values - x.tensor).max())
assert dd(x).item() == ref
for name, fn, b in (("|/x", lambda: ops.reduce("max", x), 8), ("|\\x", lambda: ops.scan("max", x), 16),
                    ("|/(|\\x)-x", lambda: dd(x), 8)):
    ms = timeit(fn)
    print(f"{name:12s} n={n}: {ms*1e3:8.1f} us  frac {b*n/(ms*1e-3)/1e9/6546.2:.3f}", flush=True)
