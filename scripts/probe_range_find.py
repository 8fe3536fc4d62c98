"""range / find at 1e8 (10 k keys): one call each after warm-up (for an ncu launch list) + wall time per call."""
import sys
import time
import torch
sys.path.insert(0, ".")
import klongpy_b200  # noqa: E402
from klongpy_b200 import DeviceArray, ops  # noqa: E402
dev = torch.device("cuda:0")
be = klongpy_b200.B200BackendProvider(device="cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
k = DeviceArray(torch.randint(0, 10_000, (100_000_000,), dtype=torch.int64, device=dev, generator=g))
for name, fn in (("range", lambda: ops.unique_first(k)), ("find", lambda: ops.find_equal(k, 77))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms per call", flush=True)
