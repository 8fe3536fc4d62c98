"""Per-call latency of the backend at bench_compiler.kg sizes (config 1): µs per call, N calls then one sync."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import klongpy_b200  # noqa: E402
from klongpy_b200 import ops  # noqa: E402

be = klongpy_b200.B200BackendProvider(device="cuda:0")
rng = np.random.default_rng(0)
N = 2000


def per_call(fn):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(N):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / N * 1e6


def per_call_np(fn):
    for _ in range(20):
        fn()
    t0 = time.perf_counter()
    for _ in range(N):
        fn()
    return (time.perf_counter() - t0) / N * 1e6


for n in (1000, 100_000):
    a, b = rng.random(n), rng.random(n)
    A, B = be.kg_asarray(a), be.kg_asarray(b)
    add_ir = ("binop", "+", ("var", "_v0"), ("var", "_v1"))
    fadd, _ = be.compile_expr_ir(add_ir, ["a", "b"])
    red_ir = ("reduce", "+", ("binop", "*", ("var", "_v0"), ("var", "_v1")))
    fred, _ = be.compile_expr_ir(red_ir, ["a", "b"])
    scan_ir = ("scan", "+", ("var", "_v0"))
    fscan, _ = be.compile_expr_ir(scan_ir, ["a"])
    print(f"n={n}: a+b eager {per_call(lambda: ops.binary('add', A, B).tensor):.1f} us | compiled {per_call(lambda: fadd(A, B).tensor):.1f} us "
          f"(numpy {per_call_np(lambda: a + b):.1f}) | +/a*b {per_call(lambda: fred(A, B)):.1f} us (numpy {per_call_np(lambda: np.add.reduce(a * b)):.1f}) | "
          f"+\\a {per_call(lambda: fscan(A)):.1f} us (numpy {per_call_np(lambda: np.cumsum(a)):.1f})", flush=True)
