"""1brc text ingest: the oracle's `parse_1brc` against the fixture produced by the REAL
reference example (tests/golden/make_golden_1brc.py), the b200 pipeline on the CPU test double, and — on the GPU —
the kernels against the oracle on synthetic files with the edge cases of the format."""
import json
import os

import numpy as np
import pytest

from oracle import kg_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(HERE, "golden", "golden_1brc.json"), encoding="utf-8"))


def _table(names, ids, temps):
    mn, mx, sm, ct = O.grouped_stats(ids, temps, len(names))
    return {s: [mn[i], mx[i], sm[i], int(ct[i])] for i, s in enumerate(names)}


def _check_table(got, want):
    assert sorted(got) == sorted(want)
    for s, (mn, mx, sm, ct) in want.items():
        g = got[s]
        assert g[0] == mn and g[1] == mx and g[3] == ct, s
        assert abs(g[2] - sm) <= 1e-12 * max(1.0, abs(sm)), s


def test_oracle_matches_reference_example(golden):
    names, ids, temps = O.parse_1brc(golden["text"].encode("utf-8"))
    assert names == sorted(names, key=lambda s: s.encode("utf-8")) and len(ids) == golden["text"].count("\n")
    _check_table(_table(names, ids, temps), golden["table_min_max_sum_count"])
    assert np.signbit(temps[list(golden["text"].split("\n")).index("a;-0.0")])  # float("-0.0") keeps its sign


def test_b200_pipeline_on_cpu_double(golden):
    import fake_lib
    dev = fake_lib.install()
    from klongpy_b200 import ingest
    text = golden["text"].encode("utf-8")
    ids, temps, names = ingest.parse(text, device=dev)
    rn, ri, rt = O.parse_1brc(text)
    assert names == rn and np.array_equal(ids.to_numpy(), ri) and np.array_equal(temps.to_numpy(), rt)
    names, mn, mean, mx, ct = ingest.aggregate(text, device=dev)
    want = golden["table_min_max_sum_count"]
    for i, s in enumerate(names):
        assert mn[i] == want[s][0] and mx[i] == want[s][1] and ct[i] == want[s][3]
        assert abs(mean[i] - want[s][2] / want[s][3]) <= 1e-12 * max(1.0, abs(mean[i]))
    assert ingest.parse(text[:-1], device=dev)[2] == rn          # a missing final newline is tolerated by the wrapper
    from klongpy_b200 import _lib
    with pytest.raises(_lib.KgbError):
        ingest.parse(b"Abha;12.34\n", device=dev)                # not the 1brc format


def _synthetic(rng, rows, n_names):
    alphabet = list("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ éüñ'-.,")
    names = set()
    while len(names) < n_names:
        k = int(rng.integers(1, 27))
        names.add("".join(rng.choice(alphabet, k)).strip() or "q")
    names = sorted(names)
    ids = rng.integers(0, len(names), rows)
    tenths = np.clip(np.round(rng.normal(100, 250, rows)), -999, 999).astype(np.int64)
    parts = []
    for i, t in zip(ids, tenths):
        a = abs(int(t))
        parts.append(f"{names[i]};{'-' if t < 0 else ''}{a // 10}.{a % 10}\n")
    return "".join(parts).encode("utf-8")


@pytest.mark.gpu
@pytest.mark.parametrize("rows,n_names", [(1, 1), (7, 3), (70_001, 413), (600_000, 10_000)])
def test_gpu_ingest_vs_oracle(rows, n_names):
    import torch
    from klongpy_b200 import ingest
    rng = np.random.default_rng(rows)
    text = _synthetic(rng, rows, n_names)
    ids, temps, names = ingest.parse(text, device="cuda:0")
    rn, ri, rt = O.parse_1brc(text)
    assert names == rn
    assert np.array_equal(ids.to_numpy(), ri)
    t = temps.to_numpy()
    assert np.array_equal(t, rt) and np.array_equal(np.signbit(t), np.signbit(rt))
    rep = ingest.report(text, device="cuda:0")
    mn, mx, sm, ct = O.grouped_stats(ri, rt, len(rn))
    ref = {rn[i]: v for i, v in O.report_1brc(mn, mx, sm, ct).items()}
    bad = [s for s in ref if rep[s] != ref[s]]
    for s in bad:  # a mean that sits on a rounding boundary may print one digit apart when sums differ in the last ulp
        a, b = rep[s].split("/"), ref[s].split("/")
        assert a[0] == b[0] and a[2] == b[2] and abs(float(a[1]) - float(b[1])) <= 0.1 + 1e-9, s
    assert torch.cuda.is_available()


@pytest.mark.gpu
def test_gpu_ingest_fixture_and_errors(golden):
    from klongpy_b200 import _lib, ingest
    text = golden["text"].encode("utf-8")
    ids, temps, names = ingest.parse(text, device="cuda:0")
    _check_table(_table(names, ids.to_numpy(), temps.to_numpy()), golden["table_min_max_sum_count"])
    for bad in (b"Abha;12.34\n", b"Abha 12.3\n", b";1.0\n", b"Abha;1,0\n"):
        with pytest.raises(_lib.KgbError):
            ingest.parse(b"ok;1.0\n" + bad, device="cuda:0")
    assert ingest.parse(b"", device="cuda:0")[2] == []


@pytest.mark.gpu
def test_gpu_ingest_alignment_long_names_tile_edges():
    """Names of every length 1..255 (rows of the dictionary are compared a word at a time), a text buffer that starts
    at each of the four byte alignments, lines straddling the 65536-byte tile edges, and a 256-byte name rejected."""
    import torch
    from klongpy_b200 import _lib, ingest
    rng = np.random.default_rng(11)
    names = ["".join(rng.choice(list("abcdefghijklmnopqrstuvwxyz"), k)) for k in range(1, 256)]
    names += [names[40][:-1] + "#", names[40][:-2] + "##", names[200][:-1] + "#"]  # same length, differ in the last bytes
    rows = 40_000                                                                   # ~5 MB: ~80 tiles
    ids = rng.integers(0, len(names), rows)
    tenths = rng.integers(-999, 1000, rows)
    text = "".join(f"{names[i]};{'-' if t < 0 else ''}{abs(int(t)) // 10}.{abs(int(t)) % 10}\n" for i, t in zip(ids, tenths)).encode()
    rn, ri, rt = O.parse_1brc(text)
    buf = torch.zeros(len(text) + 8, dtype=torch.uint8, device="cuda:0")
    for shift in range(4):
        view = buf[shift:shift + len(text)]
        view.copy_(torch.from_numpy(np.frombuffer(text, dtype=np.uint8).copy()))
        gi, gt, gn = ingest.parse(view, device="cuda:0")
        assert gn == rn and np.array_equal(gi.to_numpy(), ri) and np.array_equal(gt.to_numpy(), rt), shift
    with pytest.raises(_lib.KgbError):
        ingest.parse(("x" * 256 + ";1.0\n").encode(), device="cuda:0")
    assert ingest.parse(("x" * 255 + ";1.0\n").encode(), device="cuda:0")[2] == ["x" * 255]


@pytest.mark.gpu
def test_gpu_ingest_every_measurement_value_and_row_estimate_miss():
    """(1) Every measurement -99.9 .. 99.9 in every spelling the format allows (d.d, dd.d, -d.d, -dd.d, 0d.d) equals
    Python's float() bit for bit — the kernel computes tenths * 0.1 corrected by its exact residual instead of a table.
    (2) A text whose three sample windows hold long lines while the rest holds short ones: kgb_brc_estimate comes out
    too low, kgb_brc_parse reports the real row count and ingest.parse repeats the pass — same result as the oracle.
    (3) rows_capacity 0 with null outputs counts the rows and builds the dictionary."""
    import ctypes
    import torch
    from klongpy_b200 import _lib, ingest, ops
    vals = [f"{s}{a}.{b}" for s in ("", "-") for a in list(range(100)) + ["00", "01", "05", "09"] for b in range(10)]
    text = "".join(f"st{k % 7};{v}\n" for k, v in enumerate(vals)).encode()
    ids, temps, names = ingest.parse(text, device="cuda:0")
    want = np.array([float(v) for v in vals])
    got = temps.to_numpy()
    assert got.shape == want.shape and np.array_equal(got.view(np.int64), want.view(np.int64))   # incl. the sign of -0.0
    rn, ri, rt = O.parse_1brc(text)
    assert names == rn and np.array_equal(ids.to_numpy(), ri)
    # (2)
    rng = np.random.default_rng(4)
    long_names = ["L" + "".join(rng.choice(list("abcdefghij"), 40)) for _ in range(50)]
    short_names = ["s%d" % k for k in range(90)]
    win = 4 << 20

    def block(names_, nbytes):
        out, size = [], 0
        while size < nbytes:
            line = f"{names_[int(rng.integers(0, len(names_)))]};{int(rng.integers(-99, 100))}.{int(rng.integers(0, 10))}\n"
            out.append(line)
            size += len(line)
        return "".join(out)
    text = (block(long_names, win + 4096) + block(short_names, 3 * win) + block(long_names, win + 4096) +
            block(short_names, 3 * win) + block(long_names, win + 4096)).encode()
    dev = torch.device("cuda:0")
    t = torch.from_numpy(np.frombuffer(text, dtype=np.uint8).copy()).to(dev)
    lib, stream = ops._lib_for(dev)
    wsb = ctypes.c_size_t()
    _lib.check(lib.kgb_brc_workspace_bytes(t.numel(), ctypes.byref(wsb)), "ws")
    ws = torch.empty(wsb.value, dtype=torch.uint8, device=dev)
    est, exact = ctypes.c_int64(), ctypes.c_int64()
    _lib.check(lib.kgb_brc_estimate(t.data_ptr(), t.numel(), ctypes.byref(est), ws.data_ptr(), wsb.value, stream), "estimate")
    _lib.check(lib.kgb_brc_scan(t.data_ptr(), t.numel(), ctypes.byref(exact), ws.data_ptr(), wsb.value, stream), "scan")
    rows = text.count(b"\n")
    assert exact.value == rows and est.value < rows          # the windows were not representative
    gi, gt, gn = ingest.parse(t, device=dev)
    rn, ri, rt = O.parse_1brc(text)
    assert gn == rn and np.array_equal(gi.to_numpy(), ri) and np.array_equal(gt.to_numpy(), rt)
    # (3)
    dn = torch.zeros(ingest.DICT_CAPACITY * 256, dtype=torch.uint8, device=dev)
    dl = torch.empty(ingest.DICT_CAPACITY, dtype=torch.int32, device=dev)
    r, nn = ctypes.c_int64(), ctypes.c_int64()
    _lib.check(lib.kgb_brc_parse(t.data_ptr(), t.numel(), 0, None, None, dn.data_ptr(), dl.data_ptr(), ingest.DICT_CAPACITY,
                                 ctypes.byref(r), ctypes.byref(nn), ws.data_ptr(), wsb.value, stream), "count only")
    assert r.value == rows and nn.value == len(rn)
    with pytest.raises(_lib.KgbError):                       # no final newline
        lib_text = t[:-1].contiguous()
        _lib.check(lib.kgb_brc_parse(lib_text.data_ptr(), lib_text.numel(), 0, None, None, dn.data_ptr(), dl.data_ptr(),
                                     ingest.DICT_CAPACITY, ctypes.byref(r), ctypes.byref(nn), ws.data_ptr(), wsb.value, stream), "parse")


@pytest.mark.gpu
def test_gpu_ingest_text_lengths_around_tile_edges():
    """Texts whose LAST newline falls exactly on, one before and one after a 65536-byte tile edge (the one-pass kernel's
    tiles, look-back chain and the final-newline check all key on that position), for one to three tiles, with short
    and near-maximal names around the edge."""
    from klongpy_b200 import ingest
    rng = np.random.default_rng(17)
    names = ["n%d" % k for k in range(50)] + ["".join(rng.choice(list("abcdefgh"), int(k))) for k in rng.integers(100, 250, 20)]
    for target in (65535, 65536, 65537, 131071, 131072, 131073, 196607, 196608, 196609):
        out, size = [], 0
        while size < target - 300:
            line = f"{names[int(rng.integers(0, len(names)))]};{int(rng.integers(-99, 100))}.{int(rng.integers(0, 10))}\n"
            out.append(line)
            size += len(line)
        while target - size > 260:                       # leave a last line of 6 .. 260 bytes
            line = f"pad;{int(rng.integers(0, 10))}.{int(rng.integers(0, 10))}\n"
            out.append(line)
            size += len(line)
        rest = target - size                             # the last line: name of rest - 5 bytes + ";1.5\n"
        out.append("z" * (rest - 5) + ";1.5\n")
        text = "".join(out).encode()
        assert len(text) == target
        gi, gt, gn = ingest.parse(text, device="cuda:0")
        rn, ri, rt = O.parse_1brc(text)
        assert gn == rn and np.array_equal(gi.to_numpy(), ri) and np.array_equal(gt.to_numpy(), rt), target


def _file_case(tmp_path, rows, n_names, tail_newline=True, seed=3):
    rng = np.random.default_rng(seed)
    text = _synthetic(rng, rows, n_names)
    if not tail_newline:
        text = text[:-1]
    path = tmp_path / "measurements.txt"
    path.write_bytes(text)
    rn, ri, rt = O.parse_1brc(text)
    return str(path), rn, O.grouped_stats(ri, rt, len(rn))


def _check_file_result(res, rn, stats):
    names, mn, mean, mx, ct = res
    rmn, rmx, rsm, rct = stats
    assert names == rn and np.array_equal(mn, rmn) and np.array_equal(mx, rmx) and np.array_equal(ct, rct)
    assert np.allclose(mean, rsm / rct, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("tail_newline", [True, False])
def test_aggregate_file_on_cpu_double(tmp_path, tail_newline):
    """File -> line-aligned chunks -> per-chunk tables merged by name: host logic on the CPU test double, chunks of
    64 KB so the 25 k-row file takes several rounds and partial lines are carried across them."""
    import fake_lib
    dev = fake_lib.install()
    from klongpy_b200 import ingest
    path, rn, stats = _file_case(tmp_path, 25_000, 300, tail_newline)
    _check_file_result(ingest.aggregate_file(path, device=dev, chunk_bytes=1 << 16), rn, stats)
    empty = tmp_path / "empty.txt"
    empty.write_bytes(b"")
    assert ingest.aggregate_file(str(empty), device=dev)[0] == []
    one = tmp_path / "one.txt"
    one.write_bytes(b"a;1.0")
    assert ingest.aggregate_file(str(one), device=dev)[0] == ["a"]


@pytest.mark.gpu
def test_gpu_aggregate_file(tmp_path):
    from klongpy_b200 import ingest
    path, rn, stats = _file_case(tmp_path, 400_000, 5_000, tail_newline=False, seed=8)
    _check_file_result(ingest.aggregate_file(path, device="cuda:0", chunk_bytes=1 << 20), rn, stats)   # ~6 chunks
    _check_file_result(ingest.aggregate_file(path, device="cuda:0"), rn, stats)                          # one chunk
