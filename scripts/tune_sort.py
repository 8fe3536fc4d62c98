"""Time grade-up on the three key distributions of config 3 (permutation / 10k stations / full 64-bit) and check it."""
import sys

import torch

sys.path.insert(0, ".")
from klongpy_b200 import ops, DeviceArray  # noqa: E402

dev = torch.device("cuda:0")
ops.set_default_device(dev)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
tag = sys.argv[2] if len(sys.argv) > 2 else ""
g = torch.Generator(device=dev).manual_seed(11)


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


cases = {
    "perm": lambda: torch.randperm(n, device=dev, generator=g),
    "10k": lambda: torch.randint(0, 10_000, (n,), dtype=torch.int64, device=dev, generator=g),
    "full64": lambda: torch.randint(-2**62, 2**62, (n,), dtype=torch.int64, device=dev, generator=g),
}
out = []
for name, mk in cases.items():
    k = mk()
    K = DeviceArray(k)
    r = ops.argsort(K)
    ref = torch.argsort(k, stable=True)
    assert torch.equal(r.tensor, ref), f"{name}: mismatch"
    del ref, r
    out.append(f"{name} {timeit(lambda: ops.argsort(K)):.3f} ms")
    del K, k
print(f"sort[{tag}] n={n}: " + " | ".join(out), flush=True)
