"""Time DeviceTable operations on cuda:0 (CUDA events on the current stream, best of 3 after a warm-up)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import klongpy_b200
from klongpy_b200 import ops
from klongpy_b200.array import DeviceArray
from klongpy_b200.table import DeviceTable


def timed(fn, reps=3):
    fn()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


be = klongpy_b200.B200BackendProvider(device="cuda:0")
g = torch.Generator(device="cuda:0").manual_seed(0)
for n in (100_000_000,):
    k = DeviceArray(torch.randint(0, 10_000, (n,), device="cuda:0", generator=g))
    v = DeviceArray(torch.round(torch.randn(n, device="cuda:0", dtype=torch.float64, generator=g) * 150) / 10)
    t = DeviceTable({"k": k, "v": v})
    ms = timed(lambda: t.group_stats("k", "v"))
    print(f"group_stats dense ids   rows={n}: {ms:.2f} ms  {n / ms / 1e6:.1f} Grows/s  ({16 * n / ms / 1e6:.0f} GB/s of 16 B/row)")
    ks = DeviceArray(k._t * 1_000_003 - 17)
    t2 = DeviceTable({"k": ks, "v": v})
    ms = timed(lambda: t2.group_stats("k", "v"))
    print(f"group_stats sparse keys rows={n}: {ms:.2f} ms  {n / ms / 1e6:.1f} Grows/s  (radix-sort dictionary encode first)")
n = 20_000_000
a = DeviceArray(torch.randint(0, 1_000_000, (n,), device="cuda:0", generator=g))
b = DeviceArray(torch.randint(0, 100, (n,), device="cuda:0", generator=g))
c = DeviceArray(torch.randint(0, 2, (n,), device="cuda:0", generator=g))


def idx():
    t = DeviceTable({"a": a, "b": b, "c": c})
    t.set_index(["a", "b"])
    return t


ms = timed(idx)
t = idx()
print(f"set_index [a b] rows={n}: {ms:.2f} ms -> {len(t)} rows")
up = np.stack([np.random.default_rng(0).integers(0, 1_000_000, 100_000), np.random.default_rng(1).integers(0, 100, 100_000),
               np.full(100_000, 7)], 1)


def upsert():
    t.insertb(up)
    t.commit()


ms = timed(upsert, reps=2)
print(f"indexed upsert of 1e5 rows into {len(t)}: {ms:.2f} ms")
