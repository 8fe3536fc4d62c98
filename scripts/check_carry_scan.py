"""Pipelined scan with a carried seed and the two-output reduce at sizes that take the persistent kernels, against torch
(the check that exposed the stage hand-over race of round 1; kept as a quick standalone probe)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from klongpy_b200 import _lib, ops
from klongpy_b200.array import DeviceArray
dev = torch.device("cuda:0")
ops.set_default_device(dev)
for n in (1_000_003, 20_000_000, 20_000_001, 1 << 26):
    ki = torch.randint(0, 256, (n,), dtype=torch.int64, device=dev)
    x = ops.asarray(ki.to(torch.float64) / 256)
    carry = ops.asarray(torch.tensor(12345.5, dtype=torch.float64, device=dev))
    s = ops.run([ops.OP_ARG, 0], [x, carry], mode=_lib.MODE_SCAN, red=_lib.RED_ADD, code2=[ops.OP_ARG, 1])
    ref = torch.cumsum(ki, 0).to(torch.float64) / 256 + 12345.5
    bad = (s.tensor != ref).nonzero()
    print(n, "scan+carry mismatches:", bad.numel(), bad[:3].flatten().tolist())
    red = [ops.OP_ARG, 0] + ops.lit(2) + [ops.OPC["mul"]] + ops.lit(1) + [ops.OPC["add"]] + ops.lit(2) + [ops.OPC["pow"]]
    a, b = ops.reduce2(red, "+", [ops.OP_ARG, 0], "+", [x])
    ra = (((x.tensor * 2) + 1) ** 2).sum().item()
    rb = ki.sum().item() / 256.0
    print(n, "reduce2:", a.item() == ra, b.item() == rb, a.item(), ra, b.item(), rb)
