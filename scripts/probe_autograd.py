"""Kernel-by-kernel view of the config-5 autograd paths (torch.profiler / CUPTI): what launches, how long, and the wall
time of one gradient."""
import sys
import time
import torch
from torch.profiler import profile, ProfilerActivity

sys.path.insert(0, ".")
import klongpy_b200 as kb  # noqa: E402
from klongpy_b200 import ops, DeviceArray, lazy  # noqa: E402

dev = torch.device("cuda:0")
ops.set_default_device(dev)
be = kb.B200BackendProvider(device="cuda:0")
m = 10_000_000
g = torch.Generator(device=dev).manual_seed(3)
R = DeviceArray(torch.rand(m, dtype=torch.float64, device=dev, generator=g) * 0.14 + 0.01)
V = DeviceArray(torch.rand(m, dtype=torch.float64, device=dev, generator=g) * 0.35 + 0.05)
W = DeviceArray(torch.full((m,), 1.0 / m, dtype=torch.float64, device=dev))
X = DeviceArray(torch.rand(m, dtype=torch.float64, device=dev, generator=g) * 2)
Y = DeviceArray(torch.rand(m, dtype=torch.float64, device=dev, generator=g))
sharpe_ir = ("binop", "%", ("reduce", "+", ("binop", "*", ("var", "_v0"), ("var", "_v1"))),
             ("binop", "^", ("reduce", "+", ("binop", "*", ("binop", "^", ("var", "_v0"), ("literal", 2)),
                                             ("binop", "^", ("var", "_v2"), ("literal", 2)))), ("literal", 0.5)))
sharpe, _ = be.compile_expr_ir(sharpe_ir, ["x", "returns", "vols"])


def nn_loss(params):
    w, b = params
    s = 1 / (1 + ops.unary("exp", 0 - ((w * X) + b)))
    return ops.reduce("add", (s - Y) ** 2)


def run(name, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    print(f"== {name}: {(time.perf_counter() - t0) * 1e5:.1f} us wall per gradient")
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        fn()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    tot = 0.0
    for e in evs:
        tot += e.device_time
        print(f"   {e.device_time:8.1f} us  {e.name[:90]}")
    print(f"   kernels {len(evs)}  device total {tot:.1f} us", flush=True)


run("sharpe grad", lambda: be.compute_autograd(lambda x: sharpe(x, R, V), W))
run("nn grad eager", lambda: be.compute_multi_autograd(nn_loss, [0.1, 0.1]))
lazy.enable(True)
run("nn grad lazy", lambda: be.compute_multi_autograd(nn_loss, [0.1, 0.1]))
