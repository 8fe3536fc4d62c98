"""Throughput of the 1brc ingest: a 1e6-row synthetic file tiled to ~1.4 GB on the device."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
from klongpy_b200 import ingest, ops  # noqa: E402

dev = torch.device("cuda:0")
ops.set_default_device(dev)
rng = np.random.default_rng(5)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
names = sorted({"".join(rng.choice(list("abcdefghijklmnopqrstuvwxyz ABCDEFGHIJK"), int(rng.integers(3, 20)))).strip() or "q"
                for _ in range(12_000)})[:10_000]
ids = rng.integers(0, len(names), 1_000_000)
tenths = np.clip(np.round(rng.normal(100, 250, ids.size)), -999, 999).astype(np.int64)
t0 = time.perf_counter()
block = "".join(f"{names[i]};{'-' if t < 0 else ''}{abs(int(t)) // 10}.{abs(int(t)) % 10}\n" for i, t in zip(ids, tenths)).encode()
print(f"host text build {time.perf_counter() - t0:.1f} s, {len(block) / 1e6:.1f} MB per 1e6 rows", flush=True)
text = torch.from_numpy(np.frombuffer(block, dtype=np.uint8).copy()).to(dev).repeat(reps)
n = text.numel()
rows = ids.size * reps


def timeit(fn, k=3):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(k):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / k


i2, t2, nm = ingest.parse(text)
assert len(nm) == len(names) and i2.size == rows
ref_ids = torch.from_numpy(ids).to(dev).repeat(reps)
ref_t = torch.from_numpy(tenths / 10.0).to(dev).repeat(reps)
assert nm == sorted(names, key=lambda s: s.encode()) and torch.equal(i2.tensor, ref_ids) and torch.equal(t2.tensor, ref_t)
ms = timeit(lambda: ingest.parse(text))
print(f"parse     rows={rows} bytes={n}: {ms:.2f} ms  {rows / ms / 1e6:.2f} Grows/s  {n / ms / 1e6:.0f} GB/s of text", flush=True)
ms = timeit(lambda: ingest.aggregate(text))
print(f"aggregate rows={rows}: {ms:.2f} ms  {rows / ms / 1e6:.2f} Grows/s", flush=True)
