"""range (?a): hash fast path vs sort path — correctness against a torch construction and timing."""
import os
import sys
import torch
sys.path.insert(0, ".")
from klongpy_b200 import ops, DeviceArray  # noqa: E402

dev = torch.device("cuda:0")
ops.set_default_device(dev)
g = torch.Generator(device=dev).manual_seed(21)


def ref_unique_first(x):
    bits = x.view(torch.int64) if x.dtype == torch.float64 else x
    u, inv = torch.unique(bits, return_inverse=True)
    first = torch.full((u.numel(),), x.numel(), dtype=torch.int64, device=x.device)
    first.scatter_reduce_(0, inv, torch.arange(x.numel(), device=x.device), "amin")
    return x[torch.sort(first).values]


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


n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
cases = {
    "10k int64": torch.randint(0, 10_000, (n,), dtype=torch.int64, device=dev, generator=g),
    "with -1 keys": torch.randint(-3, 50, (n,), dtype=torch.int64, device=dev, generator=g),
    "500k distinct": torch.randint(0, 500_000, (n,), dtype=torch.int64, device=dev, generator=g),
    "600k distinct (fallback)": torch.randint(0, 600_000, (n,), dtype=torch.int64, device=dev, generator=g),
    "permutation (fallback)": torch.randperm(n, device=dev, generator=g),
}
f = torch.randint(0, 1000, (n,), device=dev, generator=g).to(torch.float64) / 8
f[::1001] = float("nan")
f[5::997] = -0.0
f[7::991] = 0.0
cases["float64 nan/-0/+0"] = f
for name, x in cases.items():
    r = ops.unique_first(DeviceArray(x)).tensor
    want = ref_unique_first(x)
    a = r.view(torch.int64) if r.dtype == torch.float64 else r
    b = want.view(torch.int64) if want.dtype == torch.float64 else want
    assert a.shape == b.shape and torch.equal(a, b), name
    ms = timeit(lambda: ops.unique_first(DeviceArray(x)))
    print(f"{name:28s} n={n} distinct={r.numel():9d}: {ms:7.3f} ms", flush=True)
    del want, r
