"""Latency of the grouped-aggregation exchanges under torchrun: all_gather vs hash-partitioned all_to_all, piece by piece."""
import os
import sys
import torch
import torch.distributed as td
sys.path.insert(0, ".")
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
import klongpy_b200  # noqa: E402
from klongpy_b200 import DeviceArray, dist, ops  # noqa: E402
dev = torch.device("cuda", local)
G = 10_000
n = 20_000_000
g = torch.Generator(device=dev).manual_seed(rank)
K = DeviceArray(torch.randint(0, G, (n,), dtype=torch.int64, device=dev, generator=g))
V = DeviceArray(torch.randn(n, dtype=torch.float64, device=dev, generator=g))


def timeit(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    td.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / reps], device=dev)
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return t.item()


B = (G + world - 1) // world
send = torch.zeros((world, 4, B), dtype=torch.int64, device=dev)
recv = torch.empty_like(send)
mine = torch.zeros((4, B), dtype=torch.int64, device=dev)
full = torch.empty((world, 4, B), dtype=torch.int64, device=dev)
block = torch.zeros((4, G), dtype=torch.int64, device=dev)
allb = torch.empty((world, 4, G), dtype=torch.int64, device=dev)
res = {
    "local aggregation only (20M rows)": timeit(lambda: ops.groupagg(K, V, G)),
    "all_to_all_single of the owner slices": timeit(lambda: td.all_to_all_single(recv.view(-1), send.view(-1))),
    "all_gather of the owner results": timeit(lambda: td.all_gather_into_tensor(full.view(-1), mine.view(-1))),
    "all_gather of the whole blocks": timeit(lambda: td.all_gather_into_tensor(allb.view(-1), block.view(-1))),
    "sharded_groupagg allgather": timeit(lambda: dist.sharded_groupagg(K, V, G, exchange="allgather")),
    "sharded_groupagg alltoall": timeit(lambda: dist.sharded_groupagg(K, V, G, exchange="alltoall")),
}
if rank == 0:
    for k, v in res.items():
        print(f"N={world} {k}: {v * 1e3:.1f} us", flush=True)
td.barrier()
td.destroy_process_group()
