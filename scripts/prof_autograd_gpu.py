"""Host-side profile of compute_multi_autograd on the sigmoid-NN loss (config 5) at 1e7 elements on the device."""
import cProfile
import pstats
import sys
import time
import torch
sys.path.insert(0, ".")
import klongpy_b200  # noqa: E402
from klongpy_b200 import DeviceArray, ops  # noqa: E402

be = klongpy_b200.B200BackendProvider(device="cuda:0")
n = 10_000_000
g = torch.Generator(device="cuda").manual_seed(3)
X = DeviceArray(torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2)
Y = DeviceArray(torch.rand(n, dtype=torch.float64, device="cuda", generator=g))


def loss(params):
    w, b = params
    s = 1 / (1 + ops.unary("exp", 0 - ((w * X) + b)))
    return ops.reduce("add", (s - Y) ** 2)


for _ in range(20):
    be.compute_multi_autograd(loss, [0.1, 0.1])
torch.cuda.synchronize()
K = 300
t0 = time.perf_counter()
for _ in range(K):
    gw, gb = be.compute_multi_autograd(loss, [0.1, 0.1])
torch.cuda.synchronize()
print(f"{(time.perf_counter() - t0) / K * 1e3:.3f} ms per call")
pr = cProfile.Profile()
pr.enable()
for _ in range(K):
    be.compute_multi_autograd(loss, [0.1, 0.1])
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
