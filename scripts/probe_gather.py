"""Random gather at 1e8 elements: kernel geometry sweep (KGB200_GATHER_U x KGB200_GATHER_CTAS) against torch indexing."""
import os
import sys

import torch

sys.path.insert(0, ".")
import klongpy_b200  # noqa: E402
from klongpy_b200 import DeviceArray, ops  # noqa: E402

be = klongpy_b200.B200BackendProvider(device="cuda:0")
n = 100_000_000
g = torch.Generator(device="cuda").manual_seed(1)
a = DeviceArray(torch.rand(n, dtype=torch.float64, device="cuda", generator=g))
p = DeviceArray(torch.randperm(n, device="cuda", generator=g))


def t(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print(f"torch a[p]: {t(lambda: a.tensor[p.tensor]):.3f} ms   torch.gather: {t(lambda: torch.gather(a.tensor, 0, p.tensor)):.3f} ms")
ref = a.tensor[p.tensor]
for U, block, ctas in ((4, 128, 0), (2, 128, 0), (8, 128, 0), (4, 256, 0), (8, 256, 0), (4, 128, 64), (4, 256, 32), (4, 128, 128)):
    os.environ["KGB200_GATHER_U"], os.environ["KGB200_GATHER_BLOCK"], os.environ["KGB200_GATHER_CTAS"] = str(U), str(block), str(ctas)
    ms = t(lambda: ops.gather(a, p))
    ok = torch.equal(ops.gather(a, p).tensor, ref)
    print(f"U={U:2d} block={block} ctas/SM={ctas or 'all chunks'}: {ms:.3f} ms  {'ok' if ok else 'MISMATCH'}", flush=True)
