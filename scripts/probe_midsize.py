"""Per-call time of the non-streaming primitives at mid sizes (1e5 .. 1e7): where fixed costs show."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
import klongpy_b200  # noqa: E402
from klongpy_b200 import DeviceArray, ops  # noqa: E402

dev = torch.device("cuda:0")
be = klongpy_b200.B200BackendProvider(device="cuda:0")
g = torch.Generator(device=dev).manual_seed(1)


def per_call(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


for n in (100_000, 1_000_000, 10_000_000):
    perm = DeviceArray(torch.randperm(n, device=dev, generator=g))
    k10 = DeviceArray(torch.randint(0, 10_000, (n,), dtype=torch.int64, device=dev, generator=g))
    v = DeviceArray(torch.randn(n, dtype=torch.float64, device=dev, generator=g))
    hp, hk = perm.to_numpy(), k10.to_numpy()
    rows = {
        "grade perm": (lambda: ops.argsort(perm), lambda: np.argsort(hp, kind="stable")),
        "grade 10k": (lambda: ops.argsort(k10), lambda: np.argsort(hk, kind="stable")),
        "group 10k": (lambda: ops.group(k10), None),
        "find": (lambda: ops.find_equal(k10, 77), lambda: np.where(hk == 77)[0]),
        "range 10k": (lambda: ops.unique_first(k10), None),
        "groupagg 10k": (lambda: ops.groupagg(k10, v, 10_000), None),
        "gather": (lambda: ops.gather(v, perm), None),
    }
    out = []
    for name, (fn, ref) in rows.items():
        us = per_call(fn)
        rs = ""
        if ref is not None and n <= 1_000_000:
            t0 = time.perf_counter()
            ref()
            rs = f" (numpy {1e6 * (time.perf_counter() - t0):.0f})"
        out.append(f"{name} {us:.0f}{rs}")
    print(f"n={n}: " + " | ".join(out) + "  [us per call]", flush=True)
