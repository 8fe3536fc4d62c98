"""cProfile of the two config-5 gradient steps at 1e7 elements (host-side time per step is what bounds them)."""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
import klongpy_b200  # noqa: E402
from klongpy_b200 import DeviceArray, ops  # noqa: E402

be = klongpy_b200.B200BackendProvider(device="cuda:0")
dev = torch.device("cuda:0")
m = 10_000_000
g = torch.Generator(device=dev).manual_seed(1)
X = DeviceArray(torch.rand(m, dtype=torch.float64, device=dev, generator=g) * 2)
Y = DeviceArray(torch.rand(m, dtype=torch.float64, device=dev, generator=g))
R = DeviceArray(torch.rand(m, dtype=torch.float64, device=dev, generator=g) * 0.14 + 0.01)
V = DeviceArray(torch.rand(m, dtype=torch.float64, device=dev, generator=g) * 0.35 + 0.05)
W = DeviceArray(torch.full((m,), 1.0 / m, dtype=torch.float64, device=dev))


def nn_loss(params):
    w, b = params
    s = 1 / (1 + ops.unary("exp", 0 - ((w * X) + b)))
    return ops.reduce("add", (s - Y) ** 2)


sharpe_ir = ("binop", "%", ("reduce", "+", ("binop", "*", ("var", "_v0"), ("var", "_v1"))),
             ("binop", "^", ("reduce", "+", ("binop", "*", ("binop", "^", ("var", "_v0"), ("literal", 2)),
                                             ("binop", "^", ("var", "_v2"), ("literal", 2)))), ("literal", 0.5)))
sharpe, _ = be.compile_expr_ir(sharpe_ir, ["x", "returns", "vols"])
steps = {"sigmoid_nn": lambda: be.compute_multi_autograd(nn_loss, [0.1, 0.1]),
         "sharpe": lambda: be.compute_autograd(lambda x: sharpe(x, R, V), W)}
for name, fn in steps.items():
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"== {name}: host {1e3 * (t1 - t0) / 50:.3f} ms per step, with final sync {1e3 * (t2 - t0) / 50:.3f} ms")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(50):
        fn()
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)
