"""Time the elementwise / reduce kernels (KGB200_NO_WIDE=1 selects 128-bit instead of 256-bit global accesses)."""
import os
import sys

import torch

sys.path.insert(0, ".")
import klongpy_b200  # noqa: E402
from klongpy_b200 import _lib, ops  # noqa: E402
from klongpy_b200.ops import OPC, OP_ARG, lit  # noqa: E402

dev = torch.device("cuda:0")
be = klongpy_b200.B200BackendProvider(device="cuda:0")
peak = 6546.2


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


for n in (100_000_000, 1 << 28):
    a = ops.asarray(torch.rand(n, dtype=torch.float64, device=dev))
    b = ops.asarray(torch.rand(n, dtype=torch.float64, device=dev))
    c = ops.asarray(torch.rand(n, dtype=torch.float64, device=dev))
    assert torch.equal(ops.binary("add", a, b).tensor, a.tensor + b.tensor)
    assert torch.equal(ops.binary("add", a[1:], b[1:]).tensor, a.tensor[1:] + b.tensor[1:])  # 8-byte aligned views
    ms = timeit(lambda: ops.binary("add", a, b))
    tms = timeit(lambda: torch.add(a.tensor, b.tensor))
    fn, _ = be.compile_expr_ir(("binop", "%", ("binop", "*", ("var", "_v0"), ("binop", "+", ("literal", 2), ("binop", "*", ("var", "_v1"), ("literal", 3)))),
                                ("binop", "-", ("binop", "+", ("var", "_v0"), ("literal", 1)), ("binop", "*", ("var", "_v1"), ("var", "_v2")))), ["a", "b", "c"])
    ms2 = timeit(lambda: fn(a, b, c))
    red = [OP_ARG, 0] + lit(2) + [OPC["mul"]] + lit(1) + [OPC["add"]] + lit(2) + [OPC["pow"]]
    ms3 = timeit(lambda: ops.run(red, [a], mode=_lib.MODE_REDUCE, red=_lib.RED_ADD))
    ms4 = timeit(lambda: ops.unary("neg", a))
    print(f"nowide={os.environ.get('KGB200_NO_WIDE', '0')} n={n}: a+b {ms:.3f} ms frac {24 * n / ms / 1e6 / peak:.3f} (torch {tms:.3f}) | "
          f"expr3 {ms2:.3f} ms frac {32 * n / ms2 / 1e6 / peak:.3f} | reduce {ms3:.3f} ms frac {8 * n / ms3 / 1e6 / peak:.3f} | "
          f"-a {ms4:.3f} ms frac {16 * n / ms4 / 1e6 / peak:.3f}", flush=True)
    del a, b, c
