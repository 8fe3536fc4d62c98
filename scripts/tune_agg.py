"""Time grouped min/max/sum/count for one kernel variant (KGB200_AGG_VARIANT is read per call)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from klongpy_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
ops.set_default_device(dev)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
peak = 6546.2
g = torch.Generator(device=dev).manual_seed(5)
k = torch.randint(0, G, (n,), dtype=torch.int64, device=dev, generator=g)
v = (torch.randn(n, dtype=torch.float64, device=dev, generator=g) * 15 + 10).mul_(10).round_().div_(10).clamp_(-99.9, 99.9)
K, V = ops.asarray(k), ops.asarray(v)
mn, mx, sm, ct = ops.groupagg(K, V, G)
rc = torch.bincount(k, minlength=G)
assert torch.equal(ct._t, rc), "count mismatch"
rs = torch.zeros(G, dtype=torch.float64, device=dev).index_add_(0, k, v)
assert torch.allclose(sm._t, rs, rtol=1e-12, atol=1e-9), "sum mismatch"
rmn = torch.full((G,), float("inf"), dtype=torch.float64, device=dev).scatter_reduce_(0, k, v, "amin")
rmx = torch.full((G,), float("-inf"), dtype=torch.float64, device=dev).scatter_reduce_(0, k, v, "amax")
assert torch.equal(mn._t, rmn) and torch.equal(mx._t, rmx), "min/max mismatch"


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


ms = timeit(lambda: ops.groupagg(K, V, G))
print(f"variant={os.environ.get('KGB200_AGG_VARIANT', 'dflt')} n={n} G={G}: {ms:.3f} ms  {n / ms / 1e6:.1f} Grows/s  "
      f"{16 * n / ms / 1e6:.0f} GB/s frac {16 * n / ms / 1e6 / peak:.3f}", flush=True)
