"""Time the running-sum / fused scan-reduce kernels for one tile geometry (KGB200_SCAN_BLOCK / KGB200_SCAN_ITEMS
are read at JIT time, so each geometry is its own process)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from klongpy_b200 import _lib, ops  # noqa: E402
from klongpy_b200.ops import OPC, OP_ARG  # noqa: E402

dev = torch.device("cuda:0")
ops.set_default_device(dev)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 28
peak = 6546.2


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


ki = torch.randint(0, 256, (n,), dtype=torch.int64, device=dev)
x = ops.asarray(ki.to(torch.float64) / 256)
r = ops.scan("add", x)
chk = torch.cumsum(ki, 0).to(torch.float64) / 256  # exact: every partial sum of k/256 is representable
bad = (r._t != chk).nonzero()
assert bad.numel() == 0, f"scan mismatch at {bad[:4].flatten().tolist()} of {n}"
del r, chk, ki, bad
for m in (n - 1, n - 4097, 4096 * 148 * 4 + 5):  # ragged tails through the same path
    xs = ops.asarray(x._t[:m])
    rs = ops.scan("add", xs)
    assert torch.equal(rs._t, torch.cumsum(xs._t, 0)), f"scan mismatch at n={m}"
    del xs, rs
mx = ops.scan("max", x)
assert torch.equal(mx._t, torch.cummax(x._t, 0).values), "running max mismatch"
del mx
SR = dict(mode=_lib.MODE_SCANRED, red=_lib.RED_MAX, code2=[4, OP_ARG, 0, OPC["sub"]], red2=_lib.RED_MAX)
for m in (n, n - 4097, 4096 * 148 * 4 + 5, 100_003):
    xs = ops.asarray(x._t[:m])
    got = ops.run([OP_ARG, 0], [xs], **SR).item()
    ref = (torch.cummax(xs._t, 0).values - xs._t).max().item()
    assert got == ref, f"max-drawdown mismatch at n={m}: {got} vs {ref}"
    got = ops.run([OP_ARG, 0], [xs], mode=_lib.MODE_SCANRED, red=_lib.RED_ADD, code2=[4, OP_ARG, 0, OPC["mul"]], red2=_lib.RED_ADD).item()
    ref = (torch.cumsum(xs._t[: min(m, 1 << 20)], 0) * xs._t[: min(m, 1 << 20)]).sum().item() if m <= (1 << 20) else None
    if ref is not None:
        assert abs(got - ref) <= 1e-12 * abs(ref), f"sum(cumsum*x) mismatch at n={m}: {got} vs {ref}"
    del xs
ms = timeit(lambda: ops.scan("add", x))
ms2 = timeit(lambda: ops.run([OP_ARG, 0], [x], mode=_lib.MODE_SCANRED, red=_lib.RED_MAX, code2=[4, OP_ARG, 0, OPC["sub"]],
                             red2=_lib.RED_MAX))
xi = ops.asarray(torch.randint(-1000, 1000, (n,), dtype=torch.int64, device=dev))
ms3 = timeit(lambda: ops.scan("add", xi))
print(f"pipe={os.environ.get('KGB200_PIPE', 'dflt')} nopipe={os.environ.get('KGB200_NO_PIPE', '0')} n={n}: "
      f"scan f64 {ms:.3f} ms {16 * n / ms / 1e6:.0f} GB/s frac {16 * n / ms / 1e6 / peak:.3f} | "
      f"scanred {ms2:.3f} ms {8 * n / ms2 / 1e6:.0f} GB/s frac {8 * n / ms2 / 1e6 / peak:.3f} | "
      f"scan i64 {ms3:.3f} ms frac {16 * n / ms3 / 1e6 / peak:.3f}", flush=True)
