"""One launch of each kernel family at a size well beyond L2 — the command profiled under ncu.

    python scripts/ncu_targets.py [--n 67108864] [--only scan,sort,...]

Run plain first (must exit 0), then under `ncu --set full ... -k regex:...`.  Not a benchmark: numbers printed
under a profiler are never reported.
"""
import argparse
import sys

import torch

sys.path.insert(0, ".")
from klongpy_b200 import _lib, ops  # noqa: E402
from klongpy_b200.ops import OPC, OP_ARG, lit  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 26)
ap.add_argument("--only", default="")
ap.add_argument("--reps", type=int, default=1)
args = ap.parse_args()
only = set(args.only.split(",")) if args.only else None
dev = torch.device("cuda:0")
ops.set_default_device(dev)
n = args.n
g = torch.Generator(device=dev).manual_seed(7)


def want(name):
    return only is None or name in only


x = ops.asarray(torch.rand(n, dtype=torch.float64, device=dev, generator=g))
RED = [OP_ARG, 0] + lit(2) + [OPC["mul"]] + lit(1) + [OPC["add"]] + lit(2) + [OPC["pow"]]
for _ in range(args.reps):
    if want("reduce"):
        ops.run(RED, [x], mode=_lib.MODE_REDUCE, red=_lib.RED_ADD)
    if want("scan"):
        ops.scan("add", x)
    if want("map"):
        y = ops.asarray(torch.rand(n, dtype=torch.float64, device=dev, generator=g))
        ops.binary("add", x, y)
        del y
    if want("scanred"):
        ops.run([OP_ARG, 0], [x], mode=_lib.MODE_SCANRED, red=_lib.RED_MAX, code2=[4, OP_ARG, 0, OPC["sub"]],
                red2=_lib.RED_MAX)
    if want("sort"):
        k = ops.asarray(torch.randint(-2**62, 2**62, (n,), dtype=torch.int64, device=dev, generator=g))
        ops.argsort(k)
        del k
    k10 = ops.asarray(torch.randint(0, 10_000, (n,), dtype=torch.int64, device=dev, generator=g))
    if want("sort10k"):
        ops.argsort(k10)
    if want("group"):
        ops.group(k10)
    if want("find"):
        ops.find_equal(k10, 77)
    if want("unique"):
        ops.unique_first(k10)
    if want("groupagg"):
        v = ops.asarray(torch.randn(n, dtype=torch.float64, device=dev, generator=g))
        ops.groupagg(k10, v, 10_000)
        del v
torch.cuda.synchronize()
print("ncu_targets ok")
