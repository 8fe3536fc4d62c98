"""Grade at small / medium n: packed path (sample + verify + 16 B passes, two host syncs) against the general pass."""
import os
import sys

import torch

sys.path.insert(0, ".")
import klongpy_b200  # noqa: E402
from klongpy_b200 import DeviceArray, ops  # noqa: E402

be = klongpy_b200.B200BackendProvider(device="cuda:0")
g = torch.Generator(device="cuda").manual_seed(1)


def t(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for logn in (14, 16, 17, 18, 19, 20, 22, 24):
    n = 1 << logn
    perm = DeviceArray(torch.randperm(n, device="cuda", generator=g))
    k10 = DeviceArray(torch.randint(0, 10_000, (n,), dtype=torch.int64, device="cuda", generator=g))
    row = [f"n=2^{logn}"]
    for name, keys in (("perm", perm), ("10k", k10)):
        os.environ["KGB200_SORT_PACKED"] = "0"
        a = t(lambda: ops.argsort(keys))
        os.environ["KGB200_SORT_PACKED"] = "1"
        os.environ["KGB200_SORT_PACKED_MIN"] = "1024"
        b = t(lambda: ops.argsort(keys))
        del os.environ["KGB200_SORT_PACKED_MIN"]
        row.append(f"{name}: general {a:7.1f} us  packed {b:7.1f} us  torch {t(lambda: torch.argsort(keys.tensor, stable=True)):7.1f} us")
    print("  ".join(row), flush=True)
