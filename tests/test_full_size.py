"""BASELINE.json's full sizes (1e9 keys / rows), checked through size-independent properties on the device — the oracle
cannot run at these sizes in seconds (np.argsort: 8 s per 1e8 keys; the reference's `=x` is O(G*N)).  torch is only
the checker here."""
import pytest


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["perm", "10k", "full64", "f64"])
def test_grade_up_1e9_properties(cuda_backend, kind):
    import torch
    from klongpy_b200 import DeviceArray, ops
    dev = torch.device("cuda:0")
    n = 1_000_000_000
    g = torch.Generator(device=dev).manual_seed(5)
    if kind == "perm":
        keys = torch.randperm(n, device=dev, generator=g)
    elif kind == "10k":
        keys = torch.randint(0, 10_000, (n,), dtype=torch.int64, device=dev, generator=g)
    elif kind == "full64":
        keys = torch.randint(-2**62, 2**62, (n,), dtype=torch.int64, device=dev, generator=g)
    else:
        keys = torch.randn(n, dtype=torch.float64, device=dev, generator=g).round_(decimals=3)
    perm = ops.argsort(DeviceArray(keys)).tensor
    # a permutation: the counting checksum of bincount is too big at 1e9, so use sum + sum of squares (mod 2^64) and range
    assert int(perm.min()) == 0 and int(perm.max()) == n - 1
    assert int(perm.sum()) == n * (n - 1) // 2
    ar = torch.arange(n, device=dev)
    assert int((perm * perm).sum()) == int((ar * ar).sum())  # wraps identically mod 2^64
    del ar
    sk = keys[perm]
    assert bool((sk[1:] >= sk[:-1]).all())                               # sorted
    ties = sk[1:] == sk[:-1]
    assert bool((perm[1:][ties] > perm[:-1][ties]).all())                # stable: ties keep ascending positions
    if kind == "perm":
        assert torch.equal(sk, torch.arange(n, device=dev))              # distinct keys: the permutation is unique
    del sk, ties
    # grade-down is the reversed stable ascending permutation (writer.py:97-108)
    down = ops.argsort(DeviceArray(keys), descending=True).tensor
    assert torch.equal(down, perm.flip(0))


@pytest.mark.gpu
def test_group_and_grouped_aggregation_1e9_properties(cuda_backend):
    import torch
    from klongpy_b200 import DeviceArray, ops
    dev = torch.device("cuda:0")
    n, G = 1_000_000_000, 10_000
    g = torch.Generator(device=dev).manual_seed(6)
    st = torch.randint(0, G, (n,), dtype=torch.int64, device=dev, generator=g)
    tp = (torch.randn(n, dtype=torch.float64, device=dev, generator=g) * 15 + 10).mul_(10).round_().div_(10).clamp_(-99.9, 99.9)
    mn, mx, sm, ct = (t.tensor for t in ops.groupagg(DeviceArray(st), DeviceArray(tp), G))
    ref_ct = torch.bincount(st, minlength=G)
    assert torch.equal(ct, ref_ct)                                                  # counts exact
    assert torch.equal(mn, torch.full((G,), float("inf"), dtype=torch.float64, device=dev).scatter_reduce_(0, st, tp, "amin"))
    assert torch.equal(mx, torch.full((G,), float("-inf"), dtype=torch.float64, device=dev).scatter_reduce_(0, st, tp, "amax"))
    # values are k/10 with |k| <= 999: the exact per-station sum is an integer / 10
    exact = torch.zeros(G, dtype=torch.int64, device=dev).index_add_(0, st, (tp * 10).round().to(torch.int64)).to(torch.float64) / 10
    assert float(((sm - exact).abs() / exact.abs().clamp_min(1.0)).max()) <= 1e-12
    del tp, mn, mx, sm
    # group: CSR of the stable grade; starts are the cumulative counts; every group holds one key
    order, starts, ng = ops.group(DeviceArray(st))
    assert ng == G
    assert torch.equal(starts.tensor, torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), ref_ct.cumsum(0)]))
    sk = st[order.tensor]
    assert bool((sk[1:] >= sk[:-1]).all())
    ties = sk[1:] == sk[:-1]
    assert bool((order.tensor[1:][ties] > order.tensor[:-1][ties]).all())


@pytest.mark.gpu
def test_ingest_1e9_rows_properties(cuda_backend):
    """1e9 text rows (17 GB): a 1e6-row block whose parse is checked against the oracle, tiled 1000x on the device —
    the parse of the whole buffer must be the block's parse repeated (ids index the same ascending name list), and the
    aggregation must be the block's count x 1000, same min / max, sum x 1000 within 1e-12."""
    import sys
    import os
    import numpy as np
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import synthetic_1brc_block
    from klongpy_b200 import ingest
    from oracle import kg_oracle as O
    dev = torch.device("cuda:0")
    block, names, _, _ = synthetic_1brc_block(rows=1_000_000, n_names=10_000, seed=9)
    rn, ri, rt = O.parse_1brc(block)
    reps = 1000
    text = torch.from_numpy(np.frombuffer(block, dtype=np.uint8).copy()).to(dev).repeat(reps)
    ids, temps, got_names = ingest.parse(text)
    assert got_names == rn and ids.size == reps * len(ri)
    ref_i, ref_t = torch.from_numpy(ri).to(dev), torch.from_numpy(rt).to(dev)
    assert bool((ids.tensor.view(reps, -1) == ref_i[None, :]).all())
    assert bool((temps.tensor.view(reps, -1).view(torch.int64) == ref_t.view(torch.int64)[None, :]).all())   # bit-exact, signed zeros too
    del ids, temps
    an, mn, mean, mx, ct = ingest.aggregate(text)
    rmn, rmx, rsm, rct = O.grouped_stats(ri, rt, len(rn))
    assert an == rn and np.array_equal(ct, rct * reps) and np.array_equal(mn, rmn) and np.array_equal(mx, rmx)
    assert np.allclose(mean, rsm / rct, rtol=1e-12, atol=1e-12)
