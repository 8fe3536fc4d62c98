"""Multi-process path on CPU: world_size 2 over gloo, device library replaced by the NumPy test double.
Checks the exchange logic of klongpy_b200/dist.py (all_gather of partials, carried scan offset, all_to_all of
per-owner table slices)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        import torch.distributed as td
        td.init_process_group("gloo", rank=rank, world_size=world)
        import fake_lib
        dev = fake_lib.install()
        from klongpy_b200 import B200BackendProvider, dist, ops
        be = B200BackendProvider(device=dev)
        rng = np.random.default_rng(123)
        n = 20_001
        x = rng.integers(0, 256, n) / 256.0
        i = rng.integers(-100, 100, n)
        lo, hi = dist.shard_bounds(n)
        X, I = be.kg_asarray(x[lo:hi]), be.kg_asarray(i[lo:hi])
        code = [ops.OP_ARG, 0] + ops.lit(2) + [ops.OPC["mul"]] + ops.lit(1) + [ops.OPC["add"]] + ops.lit(2) + [ops.OPC["pow"]]
        r = dist.sharded_reduce(code, [X], "+").item()
        assert r == np.add.reduce(((x * 2) + 1) ** 2), (r, rank)
        assert dist.sharded_reduce([ops.OP_ARG, 0], [I], "|").item() == i.max()
        s = dist.sharded_scan([ops.OP_ARG, 0], [X], "+").to_numpy()
        assert np.array_equal(s, np.cumsum(x)[lo:hi]), rank
        s = dist.sharded_scan([ops.OP_ARG, 0], [I], "|").to_numpy()
        assert np.array_equal(s, np.maximum.accumulate(i)[lo:hi]), rank
        s = dist.sharded_scan([ops.OP_ARG, 0], [I], "+").to_numpy()
        assert np.array_equal(s, np.cumsum(i)[lo:hi]), rank
        # fused pass: global fold of one expression + this shard's scan offset from the SAME read of the shard
        fold, off = dist.sharded_reduce_with_total(code, "+", [X], [ops.OP_ARG, 0], "+")
        assert fold.item() == np.add.reduce(((x * 2) + 1) ** 2), rank
        s = dist.sharded_scan_with_offset([ops.OP_ARG, 0], [X], "+", off).to_numpy()
        assert np.array_equal(s, np.cumsum(x)[lo:hi]), rank
        G = 37
        k = rng.integers(0, G, n)
        v = np.round(rng.normal(10, 15, n), 1)
        from oracle import kg_oracle as O
        rmn, rmx, rsm, rct = O.grouped_stats(k, v, G)
        ref_sum = None
        for exchange in ("auto", "allgather", "alltoall"):  # one all_gather + fold | hash-partitioned all_to_all
            mn, mx, sm, ct = (t.to_numpy() for t in dist.sharded_groupagg(be.kg_asarray(k[lo:hi]), be.kg_asarray(v[lo:hi]), G,
                                                                          exchange=exchange))
            assert np.array_equal(mn, rmn) and np.array_equal(mx, rmx) and np.array_equal(ct, rct), (exchange, rank)
            assert np.allclose(sm, rsm, rtol=1e-12, atol=1e-9), (exchange, rank)
            # both exchanges fold the ranks' partial sums in rank order: bit-identical to one another and on every rank
            assert ref_sum is None or np.array_equal(sm, ref_sum), (exchange, rank)
            ref_sum = sm
        # a station count that is not a multiple of the world size, and one smaller than it (empty owner slices)
        for G2 in (world * 5 + 1, 1):
            k2 = rng.integers(0, G2, n)
            got = [t.to_numpy() for t in dist.sharded_groupagg(be.kg_asarray(k2[lo:hi]), be.kg_asarray(v[lo:hi]), G2, exchange="alltoall")]
            want = O.grouped_stats(k2, v, G2)
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[3], want[3]), (G2, rank)
            assert np.allclose(got[2], want[2], rtol=1e-12, atol=1e-9), (G2, rank)
        # grade of a sharded vector (sample sort with composite splitters) and range (?a): exact against the global oracle
        for name, kk in (("perm", rng.permutation(n)), ("ties", rng.integers(0, 7, n)), ("neg", rng.integers(-50, 50, n)),
                         ("f64", np.round(rng.standard_normal(n), 1)), ("const", np.full(n, 3))):
            if name == "f64":
                kk = kk.copy()
                kk[rng.integers(0, n, 50)] = np.nan
                kk[rng.integers(0, n, 50)] = -0.0
            ref = np.argsort(kk, kind="stable")
            up = dist.sharded_argsort(be.kg_asarray(kk[lo:hi])).to_numpy()
            assert np.array_equal(up, ref[lo:hi]), (name, rank)
            down = dist.sharded_argsort(be.kg_asarray(kk[lo:hi]), descending=True).to_numpy()
            assert np.array_equal(down, ref[::-1][lo:hi]), (name, rank)
            u = dist.sharded_range(be.kg_asarray(kk[lo:hi])).to_numpy()
            want_u = O.range_unique(kk)
            assert len(u) == len(want_u) and np.array_equal(u, want_u, equal_nan=True), (name, rank)
            # group `=a` in CSR form: this rank's block of the global order + the replicated group starts
            go, gs, gg = dist.sharded_group(be.kg_asarray(kk[lo:hi]))
            want_o, want_s = O.group_csr(kk)
            assert gg == len(want_s) - 1 and np.array_equal(gs.to_numpy(), want_s), (name, rank)
            assert np.array_equal(go.to_numpy(), want_o[lo:hi]), (name, rank)
        # find: global positions of this rank's hits + every rank's count
        pos, counts = dist.sharded_find(I, 7)
        want = np.nonzero(i == 7)[0]
        assert counts.sum() == len(want) and counts[rank] == pos.size
        assert np.array_equal(pos.to_numpy(), want[(want >= lo) & (want < hi)]), rank
        # autograd over shards: loss = +/(w*x + b - y)^2 with replicated scalars w, b and a sharded weight vector u
        y = rng.random(n)
        Y = be.kg_asarray(y[lo:hi])
        lcode = [ops.OP_ARG, 0, ops.OP_ARG, 1, ops.OPC["mul"], ops.OP_ARG, 2, ops.OPC["add"], ops.OP_ARG, 3, ops.OPC["sub"],
                 ops.OPC["square"]]
        def loss(params):
            return dist.sharded_loss_sum(lcode, [params[0], X, params[1], Y])
        gw, gb = dist.combine_param_grads(be.compute_multi_autograd(loss, [0.3, -0.2]))
        d = 2 * (0.3 * x - 0.2 - y)
        assert np.allclose([float(gw.to_numpy()), float(gb.to_numpy())], [np.sum(d * x), np.sum(d)], rtol=1e-10), rank
        gu = be.compute_autograd(lambda u: dist.sharded_loss_sum([ops.OP_ARG, 0, ops.OP_ARG, 1, ops.OPC["mul"], ops.OPC["square"]],
                                                                 [u, X]), Y).to_numpy()
        assert np.allclose(gu, 2 * y[lo:hi] * x[lo:hi] ** 2, rtol=1e-12), rank   # a sharded vector's gradient stays sharded
        lv = dist.sharded_loss_sum([ops.OP_ARG, 0, ops.OPC["square"]], [X]).item()
        assert lv == np.add.reduce(x * x) or abs(lv - np.sum(x * x)) <= 1e-12 * np.sum(x * x)
        # 1brc ingest over a file split into line-aligned chunks, one per rank; per-station tables merged by name
        from klongpy_b200 import ingest
        srng = np.random.default_rng(77)
        snames = ["Abha", "Zürich", "St. John's", "a", "Ab", "x" * 40, "Ürümqi"] + [f"s{j}" for j in range(50)]
        rows = [f"{snames[j]};{t / 10:.1f}" for j, t in zip(srng.integers(0, len(snames), 4000), srng.integers(-999, 1000, 4000))]
        rows[:3] = ["only-in-first-chunk;1.0"] * 3
        text = ("\n".join(rows) + "\n").encode("utf-8")
        bounds = ingest.chunk_bounds(text, world)
        assert bounds[0][0] == 0 and bounds[-1][1] == len(text) and all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
        assert all(e == s_ or text[e - 1:e] == b"\n" for s_, e in bounds)
        s0, e0 = bounds[rank]
        gn, gmn, gmean, gmx, gct = ingest.sharded_aggregate(text[s0:e0], device=dev)
        rn, ri, rt = O.parse_1brc(text)
        rmn, rmx, rsm, rct = O.grouped_stats(ri, rt, len(rn))
        assert gn == rn and np.array_equal(gmn, rmn) and np.array_equal(gmx, rmx) and np.array_equal(gct, rct), rank
        assert np.allclose(gmean, rsm / rct, rtol=1e-12, atol=1e-12), rank
        td.barrier()
        td.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL " + traceback.format_exc()))


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_primitives_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r, status in res:
        assert status == "ok", f"rank {r}: {status}"


def test_shard_bounds_cover_range():
    from klongpy_b200 import dist
    for n in (0, 1, 7, 1000, 10**9):
        for w in (1, 2, 3, 8):
            spans = [dist.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
