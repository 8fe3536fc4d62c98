"""Multi-GPU parity under torchrun (NCCL): sharded reduce / scan / grouped aggregation vs the oracle on the full data.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/nccl_dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

sys.path.insert(0, ".")
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
import klongpy_b200  # noqa: E402
from klongpy_b200 import dist, ops  # noqa: E402
from oracle import kg_oracle as O  # noqa: E402

be = klongpy_b200.B200BackendProvider(device=f"cuda:{local}")
ok = True
for n in (1000, 1_000_003, 40_000_001):
    rng = np.random.default_rng(n)           # same data on every rank
    x = rng.integers(0, 256, n) / 256.0
    lo, hi = dist.shard_bounds(n)
    X = be.kg_asarray(x[lo:hi])
    red = [ops.OP_ARG, 0] + ops.lit(2) + [ops.OPC["mul"]] + ops.lit(1) + [ops.OPC["add"]] + ops.lit(2) + [ops.OPC["pow"]]
    r = dist.sharded_reduce(red, [X], "+").item()
    ref = O.eval_ir(("reduce", "+", ("binop", "^", ("binop", "+", ("binop", "*", ("var", "_v0"), ("literal", 2)), ("literal", 1)),
                                    ("literal", 2))), {"_v0": x})
    s = dist.sharded_scan([ops.OP_ARG, 0], [X], "+").to_numpy()
    sref = np.cumsum(x)[lo:hi]
    mx = dist.sharded_scan([ops.OP_ARG, 0], [X], "|").to_numpy()
    fold, off = dist.sharded_reduce_with_total(red, "+", [X], [ops.OP_ARG, 0], "+")   # fused pass (two-output reduce)
    s2 = dist.sharded_scan_with_offset([ops.OP_ARG, 0], [X], "+", off).to_numpy()
    assert fold.item() == ref and np.array_equal(s2, sref), "fused reduce+total path"
    G = 1000
    k = rng.integers(0, G, n)
    v = np.round(rng.normal(10, 15, n), 1)
    mn, mxg, sm, ct = (t.to_numpy() for t in dist.sharded_groupagg(be.kg_asarray(k[lo:hi]), be.kg_asarray(v[lo:hi]), G))
    rmn, rmx, rsm, rct = O.grouped_stats(k, v, G)
    for exchange in ("allgather", "alltoall"):  # both exchanges: same tables (each call re-aggregates the shard, and the
        # kernel's shared-memory float adds associate differently from run to run, so sums are compared to 1e-12)
        t2 = [t.to_numpy() for t in dist.sharded_groupagg(be.kg_asarray(k[lo:hi]), be.kg_asarray(v[lo:hi]), G, exchange=exchange)]
        assert np.array_equal(t2[0], mn) and np.array_equal(t2[1], mxg) and np.array_equal(t2[3], ct), exchange
        assert np.allclose(t2[2], sm, rtol=1e-12, atol=1e-9), exchange
    # multi-GPU grade (sample sort, composite splitters, two all_to_alls) and range — SURVEY.md §8e rows 5-6
    for kk in (rng.permutation(n), k, np.round(v, 0)):
        gref = np.argsort(kk, kind="stable")
        up = dist.sharded_argsort(be.kg_asarray(kk[lo:hi])).to_numpy()
        down = dist.sharded_argsort(be.kg_asarray(kk[lo:hi]), descending=True).to_numpy()
        assert np.array_equal(up, gref[lo:hi]) and np.array_equal(down, gref[::-1][lo:hi]), "sharded_argsort"
    assert np.array_equal(dist.sharded_range(be.kg_asarray(k[lo:hi])).to_numpy(), O.range_unique(k)), "sharded_range"
    go, gs, gg = dist.sharded_group(be.kg_asarray(k[lo:hi]))       # `=a`: block of the global order + replicated starts
    ginv = np.argsort(k, kind="stable")
    gstarts = np.concatenate([[0], np.cumsum(np.bincount(k, minlength=G))])
    gstarts = gstarts[np.concatenate([[True], np.diff(gstarts) > 0])] if n else gstarts[:1]
    assert gg == len(gstarts) - 1 and np.array_equal(gs.to_numpy(), gstarts) and np.array_equal(go.to_numpy(), ginv[lo:hi]), "sharded_group"
    pos, counts = dist.sharded_find(be.kg_asarray(k[lo:hi]), 7)
    want = np.nonzero(k == 7)[0]
    find_ok = counts.sum() == len(want) and np.array_equal(pos.to_numpy(), want[(want >= lo) & (want < hi)])
    y = rng.random(n)
    Y = be.kg_asarray(y[lo:hi])
    lcode = [ops.OP_ARG, 0, ops.OP_ARG, 1, ops.OPC["mul"], ops.OP_ARG, 2, ops.OPC["add"], ops.OP_ARG, 3, ops.OPC["sub"], ops.OPC["square"]]
    gw, gb = dist.combine_param_grads(be.compute_multi_autograd(
        lambda p: dist.sharded_loss_sum(lcode, [p[0], X, p[1], Y]), [0.3, -0.2]))
    d = 2 * (0.3 * x - 0.2 - y)
    grad_ok = np.allclose([gw.item(), gb.item()], [np.sum(d * x), np.sum(d)], rtol=1e-10)
    good = (find_ok and grad_ok and r == ref and np.array_equal(s, sref) and np.array_equal(mx, np.maximum.accumulate(x)[lo:hi]) and
            np.array_equal(mn, rmn) and np.array_equal(mxg, rmx) and np.array_equal(ct, rct) and
            np.allclose(sm, rsm, rtol=1e-12, atol=1e-9))
    # host-resident shard streamed through the same step (stream.HostStream.sharded_reduce_and_scan): chunked H2D +
    # two-output reduce on arrival, one all_gather, seeded scan chunks overlapped with D2H
    from klongpy_b200 import stream  # noqa: E402
    hx = torch.from_numpy(x[lo:hi].copy()).pin_memory()
    hs = torch.empty(hi - lo, dtype=torch.float64).pin_memory()
    hstream = stream.HostStream(f"cuda:{local}", chunk=1 << 16)
    for _ in range(2):
        f2 = hstream.sharded_reduce_and_scan(hx, red, "+", [ops.OP_ARG, 0], "+", hs)
        torch.cuda.synchronize()
        good = good and f2.item() == ref and np.array_equal(hs.numpy(), sref)
        hs.zero_()
    print(f"rank {rank}/{world} n={n}: {'ok' if good else 'MISMATCH'}", flush=True)
    ok &= good
# 1brc ingest: the text split into line-aligned chunks, one per rank; per-station tables merged by name
from klongpy_b200 import ingest  # noqa: E402
srng = np.random.default_rng(77)
snames = sorted({"".join(srng.choice(list("abcdefghijklmnopqrstuvwxyz éü'"), int(srng.integers(1, 30)))).strip() or "q" for _ in range(3000)})
rows = [f"{snames[j]};{t / 10:.1f}" for j, t in zip(srng.integers(0, len(snames), 400_000), srng.integers(-999, 1000, 400_000))]
text = ("\n".join(rows) + "\n").encode("utf-8")
s0, e0 = ingest.chunk_bounds(text, world)[rank]
gn, gmn, gmean, gmx, gct = ingest.sharded_aggregate(text[s0:e0], device=f"cuda:{local}")
rn, ri, rt = O.parse_1brc(text)
rmn, rmx, rsm, rct = O.grouped_stats(ri, rt, len(rn))
good = gn == rn and np.array_equal(gmn, rmn) and np.array_equal(gmx, rmx) and np.array_equal(gct, rct) and \
    np.allclose(gmean, rsm / rct, rtol=1e-12, atol=1e-12)
print(f"rank {rank}/{world} ingest 400k rows / {len(rn)} stations: {'ok' if good else 'MISMATCH'}", flush=True)
ok &= good
flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
td.all_reduce(flag, op=td.ReduceOp.MIN)
td.barrier()
td.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
