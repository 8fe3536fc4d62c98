import sys, time, torch
sys.path.insert(0, ".")
from klongpy_b200 import ops, DeviceArray
dev = torch.device("cuda:0"); ops.set_default_device(dev)
for n in (100_000_000, 1_000_000_000):
    k = DeviceArray(torch.randint(0, 10_000, (n,), dtype=torch.int64, device=dev))
    for _ in range(3): r = ops.find_equal(k, 77)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): r = ops.find_equal(k, 77)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    ref = torch.nonzero(k.tensor == 77).flatten()
    assert torch.equal(ref, r.tensor)
    print(f"find n={n}: {ms*1e3:.1f} us  frac={8*n/(ms*1e-3)/1e9/6546.2:.3f} count={r.shape}")
