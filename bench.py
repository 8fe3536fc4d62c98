#!/usr/bin/env python
"""bench.py — the headline measurement (BASELINE.json config 2) plus a per-primitive table.

    python bench.py --gpus N --steps K --warmup W            # this framework (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (NumPy oracle port), rank 0 only

A *step* is one pass of the hot path over one resident batch: the fused reduce `{+/((x*2)+1)^2}` followed by the
running sum `+\\x` over a float64 vector of --n elements per GPU (default 1e9 = 8 GB, so every pass streams from
HBM: inputs are 60x larger than L2).  value = elements processed per second = 2*n*N / step time (each primitive
sweeps the n-element shard once; weak scaling: per-GPU work is fixed, the global vector is N*n).  Timed with
CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.

Extra keys: `roofline` (dominant kernel = the scan, algorithmic 16 B/elem), `cpu_baseline` (oracle port timed on
the host, bounded sample), `e2e` (same step with HOST buffers: pinned H2D of x and D2H of both results inside
the timed region), `primitives` (grade / group / find / range / grouped aggregation / elementwise at their
BASELINE sizes), `clocks`, `gpu_launches`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "elements_per_sec"
UNIT = "elements/s"
WORKLOAD = "config2: fused reduce {+/((x*2)+1)^2} + running sum +\\x over float64"


def scan_traffic(n):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu --set full capture
    (profiles/r01_scan_traffic.json), per launch; only quoted when it was captured at this problem size."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_scan_traffic.json")) as f:
            d = json.load(f)
        return float(d["traffic_bytes"]) if int(d["n_elements"]) == int(n) else None
    except Exception:
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
            "clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def reference_step(x):
    """(step, kind, how): one pass of the workload on the host array `x` through the reference's CPU implementation —
    the UNMODIFIED reference (klongpy from the git-ignored baseline/_ref install) when it travelled to
    this box, else the oracle port of its NumPy-backend calls."""
    from oracle import kg_oracle as O
    ir = ("reduce", "+", ("binop", "^", ("binop", "+", ("binop", "*", ("var", "_v0"), ("literal", 2)), ("literal", 1)), ("literal", 2)))
    sc = ("scan", "+", ("var", "_v0"))

    def port_step():
        O.eval_ir(ir, {"_v0": x})
        O.eval_ir(sc, {"_v0": x})

    how = "oracle/kg_oracle.eval_ir = numpy_backend.py:116-178"
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref_dir, "klongpy")):
        try:
            if ref_dir not in sys.path:
                sys.path.insert(0, ref_dir)
            from klongpy import KlongInterpreter
            klong = KlongInterpreter(backend="numpy")
            klong["x"] = x
            assert float(klong("+/((x*2)+1)^2")) == float(O.eval_ir(ir, {"_v0": x}))

            def ref_step():
                klong("+/((x*2)+1)^2")
                klong("+\\x")

            return ref_step, "reference", ("klongpy 0.7.0 from baseline/_ref: KlongInterpreter(backend='numpy') on "
                                          "'+/((x*2)+1)^2' and '+\\x'")
        except Exception as ex:  # fall back to the port, say why
            how += f"; baseline/_ref not usable: {type(ex).__name__}: {ex}"
    return port_step, "port", how


def run_reference(args):
    """The reference's own CPU implementation of the path = its NumPy backend calls, restated in oracle/."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = int(args.ref_n)
    rng = np.random.default_rng(0)
    x = rng.integers(0, 256, n) / 256.0
    step, kind, how = reference_step(x)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = 2 * n / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_per_step": n, "note": "bounded sample of the 1e9-element workload; "
                   "NumPy executes these primitives single-threaded"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
                         "sample": f"{n} float64 elements per step ({how})"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cpus": os.cpu_count(),
    }
    print(json.dumps(line), file=_claim_stdout(), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
_RESULT_OUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line to stdout when
    the box sets NCCL_DEBUG), so everything else this process and its libraries print is sent to stderr and only the
    result line goes to the real stdout."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    return _RESULT_OUT


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=float, default=1e9, help="float64 elements per GPU")
    ap.add_argument("--ref-n", type=float, default=1e8, help="elements per step of the CPU reference arm")
    ap.add_argument("--cpu-n", type=float, default=1e8, help="elements of the cpu_baseline sample")
    ap.add_argument("--no-primitives", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as td
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    import klongpy_b200
    from klongpy_b200 import DeviceArray, _lib, dist, ops
    be = klongpy_b200.B200BackendProvider(device=f"cuda:{local}")
    dev = torch.device("cuda", local)
    n = int(args.n)
    peak, peak_src = measured_peak()

    # synthetic shard: dyadic float64 (k/256) so the parity assertion below is bit-exact
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    ki = torch.randint(0, 256, (n,), dtype=torch.int64, device=dev, generator=g)
    X = DeviceArray(ki) / 256
    # exact witnesses in integer arithmetic: x = k/256, so +\x ends at (sum k)/256 and ((x*2)+1)^2 = (k+128)^2 / 2^14 —
    # every term and every partial sum is a dyadic rational below 2^53 * 2^-14, i.e. exact in float64 in ANY summation
    # order, on any number of ranks
    isum = int(ops.reduce("add", DeviceArray(ki)).item())
    isq = int(ops.reduce("add", (DeviceArray(ki) + 128) * (DeviceArray(ki) + 128)).item())
    del ki
    wit = torch.tensor([isum, isq], dtype=torch.int64, device=dev)
    if world > 1:
        allw = torch.empty(2 * world, dtype=torch.int64, device=dev)
        td.all_gather_into_tensor(allw, wit)
        allw = allw.reshape(world, 2).cpu().numpy()
    else:
        allw = wit.reshape(1, 2).cpu().numpy()
    want_scan_last = float(int(allw[: rank + 1, 0].sum())) / 256.0     # last element of THIS rank's slice of +\x
    want_fold = float(int(allw[:, 1].sum())) / 16384.0                 # +/((x*2)+1)^2 over the global vector
    RED_CODE = [ops.OP_ARG, 0] + ops.lit(2) + [ops.OPC["mul"]] + ops.lit(1) + [ops.OPC["add"]] + ops.lit(2) + [ops.OPC["pow"]]
    ID_CODE = [ops.OP_ARG, 0]
    # N=1 goes through the provider API the interpreter calls (compile_expr_ir -> compiled callable); N>1 through dist
    RED_IR = ("reduce", "+", ("binop", "^", ("binop", "+", ("binop", "*", ("var", "_v0"), ("literal", 2)), ("literal", 1)), ("literal", 2)))
    fn_red, _ = be.compile_expr_ir(RED_IR, ["x"])
    fn_scan, _ = be.compile_expr_ir(("scan", "+", ("var", "_v0")), ["x"])

    def step():
        if world == 1:
            r = fn_red(X)
            s = fn_scan(X)
        else:
            # the shard total the scan must carry comes out of the SAME pass as the fused reduce (two-output
            # reduce kernel), so a rank still streams 24 B/elem: reduce 8 + scan 16
            r, off = dist.sharded_reduce_with_total(RED_CODE, "+", [X], ID_CODE, "+")
            s = dist.sharded_scan_with_offset(ID_CODE, [X], "+", off)
        return r, s

    # kernels of ours per step: N=1 reduce + scan; N>1 two-output reduce, combine, (offset fold on ranks > 0), scan
    per_step_launches = 2 if world == 1 else (3 + (1 if rank > 0 else 0))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        r, s = step()
    sync_all()
    # parity inside the bench, on EVERY rank of every world size: the fold and the carried scan offset are exact
    tot = DeviceArray(s.tensor[-1]).item()
    fold = r.item()
    checks = {"scan_last_exact": tot == want_scan_last, "fold_exact": fold == want_fold}
    assert all(checks.values()), (rank, checks, tot, want_scan_last, fold, want_fold)
    okflag = torch.tensor([1], dtype=torch.int64, device=dev)
    if world > 1:
        td.all_reduce(okflag, op=td.ReduceOp.MIN)
    checks["ranks_checked"] = world if int(okflag.item()) == 1 else 0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        r, s = step()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    del r, s

    # dominant kernel (the scan) and the fused reduce timed alone, on the launching stream
    def time_kernel(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    scan_ms = time_kernel(lambda: fn_scan(X), max(3, args.steps))
    red_ms = time_kernel(lambda: fn_red(X), max(3, args.steps))

    tms = torch.tensor([ms, scan_ms, red_ms], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(tms, op=td.ReduceOp.MAX)
    ms, scan_ms, red_ms = (float(v) for v in tms.tolist())
    value = 2.0 * n * world / (ms * 1e-3)

    # ---- e2e: host buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        try:
            hx = torch.empty(n, dtype=torch.float64).pin_memory()
            hx.copy_(X.tensor)
            hs = torch.empty(n, dtype=torch.float64).pin_memory()
            hr = torch.empty(1, dtype=torch.float64).pin_memory()
            tmp = X
            del X
            torch.cuda.empty_cache()

            from klongpy_b200 import stream
            hs_stream = stream.HostStream(dev, chunk=1 << 24)

            def e2e_step():
                if world == 1:
                    # host vector -> chunked H2D | fused reduce + carried scan | D2H, overlapped on three streams
                    rr = hs_stream.reduce_and_scan(hx, RED_CODE, "+", ID_CODE, "+", hs)
                else:
                    # sharded: chunked H2D with the two-output reduce on arrival, ONE all_gather, then seeded scan chunks
                    # overlapped with D2H (the shard stays resident between the two phases)
                    rr = hs_stream.sharded_reduce_and_scan(hx, RED_CODE, "+", ID_CODE, "+", hs)
                hr.copy_(rr.tensor.reshape(1), non_blocking=True)

            X = tmp
            hs.zero_()
            e2e_step()
            sync_all()
            ok = hs[-1].item() == want_scan_last and hr[0].item() == want_fold and \
                hs[n // 2].item() == (hs[n // 2 - 1].item() + hx[n // 2].item())
            e2e_check = ("fold, scan[-1] (incl. the rank's carried offset) and a mid-shard adjacent difference exact on dyadic data, every rank"
                         if ok else "MISMATCH")
            assert ok, (rank, hs[-1].item(), want_scan_last, hr[0].item(), want_fold)
            # the host link this path is bound by: all ranks copy 1 GiB pinned buffers at the same time
            link = {}
            try:
                nb = 1 << 27
                d_a, d_b = torch.empty(nb, dtype=torch.float64, device=dev), torch.empty(nb, dtype=torch.float64, device=dev)
                s2 = torch.cuda.Stream(dev)

                def timed(fn):
                    sync_all()
                    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a_.record()
                    fn()
                    b_.record()
                    sync_all()
                    t_ = torch.tensor([a_.elapsed_time(b_)], dtype=torch.float64, device=dev)
                    if world > 1:
                        td.all_reduce(t_, op=td.ReduceOp.MAX)
                    return float(t_.item())

                def duplex():
                    s2.wait_stream(torch.cuda.current_stream(dev))
                    d_a.copy_(hx[:nb], non_blocking=True)
                    with torch.cuda.stream(s2):
                        hs[:nb].copy_(d_b, non_blocking=True)
                    torch.cuda.current_stream(dev).wait_stream(s2)

                gib = 8.0 * nb * world / 1e9
                link = {"h2d_gbs": gib / (timed(lambda: d_a.copy_(hx[:nb], non_blocking=True)) * 1e-3),
                        "d2h_gbs": gib / (timed(lambda: hs[:nb].copy_(d_b, non_blocking=True)) * 1e-3),
                        "duplex_gbs_each_way": gib / (timed(duplex) * 1e-3),
                        "note": f"aggregate over {world} rank(s), 1 GiB pinned copies issued by all ranks at once"}
                # e2e roofline: N=1 overlaps H2D with D2H (full duplex); N>1 must finish H2D before the seeded scan can start
                t_floor = (8.0 * n * world / 1e9) / link["duplex_gbs_each_way"] if world == 1 else \
                    (8.0 * n * world / 1e9) / link["h2d_gbs"] + (8.0 * n * world / 1e9) / link["d2h_gbs"]
                link["step_floor_ms"] = t_floor * 1e3
                del d_a, d_b
            except Exception as ex:  # noqa: BLE001
                link = {"error": f"{type(ex).__name__}: {ex}"}
            k = max(1, min(3, args.steps))
            e0.record()
            for _ in range(k):
                e2e_step()
            e1.record()
            sync_all()
            e_ms = torch.tensor([e0.elapsed_time(e1) / k], dtype=torch.float64, device=dev)
            if world > 1:
                td.all_reduce(e_ms, op=td.ReduceOp.MAX)
            e2e = {"value": 2.0 * n * world / (float(e_ms.item()) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * n * world,
                   "d2h_bytes_per_step": (8 * n + 8) * world, "ms_per_step": float(e_ms.item()), "steps": k,
                   "path": ("pinned host float64 -> klongpy_b200.stream.HostStream (16 Mi-element chunks: H2D | fused reduce + "
                            "carried scan | D2H overlapped on three streams) -> pinned host") if world == 1 else
                   ("pinned host shard -> HostStream.sharded_reduce_and_scan (16 Mi-element chunks: H2D + two-output reduce on "
                    "arrival | one all_gather | seeded scan chunks + D2H) -> pinned host"), "check": e2e_check,
                   "host_link": link}
            if link.get("step_floor_ms"):
                e2e["frac_of_host_link_floor"] = link["step_floor_ms"] / float(e_ms.item())
            del hx, hs, hr
        except Exception as ex:  # noqa: BLE001
            e2e = {"value": None, "unit": UNIT, "error": f"{type(ex).__name__}: {ex}"}

    # ---- per-primitive table (N=1 sizes from BASELINE.json configs 1/3/4)
    prims = {}
    if not args.no_primitives:
        xt = X.tensor
        prims["fused_reduce"] = prim_row(n, red_ms, 8, peak,
                                         torch_eager_ms=time_kernel(lambda: (((xt * 2) + 1) ** 2).sum(0), 3))
        prims["running_sum"] = prim_row(n, scan_ms, 16, peak, torch_eager_ms=time_kernel(lambda: xt.cumsum(0), 3))
        del xt
        try:
            prims.update(primitive_table(be, ops, dist, time_kernel, dev, peak, world, rank))
        except Exception as ex:  # noqa: BLE001
            prims["error"] = f"{type(ex).__name__}: {ex}"

    # ---- CPU baseline (oracle port), rank 0, bounded sample
    cpu = None
    if rank == 0:
        m = int(args.cpu_n)
        hxs = np.random.default_rng(0).integers(0, 256, m) / 256.0
        cpu_step, cpu_kind, cpu_how = reference_step(hxs)
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            cpu_step()
            best = min(best, time.perf_counter() - t0)
        cpu = {"value": 2 * m / best, "unit": UNIT, "cores": 1, "kind": cpu_kind,
               "sample": f"{m} float64 elements, best of 3 ({cpu_how}; NumPy runs these ops on one core; host has {os.cpu_count()} cpus)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_per_gpu": n, "global_n": n * world, "sharding": f"block x{world}",
                       "l2": "inputs (8 GB/GPU) are larger than L2; no flush needed",
                       "exchange": "none" if world == 1 else "one all_gather of 16 B per rank: {partial fold, shard total}; the scan carries the exclusive offset"},
            "roofline": {"bound": "hbm", "achieved": 16.0 * n / (scan_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": 16.0 * n / (scan_ms * 1e-3) / 1e9 / peak, "traffic": scan_traffic(n), "kernel": "kgb_scan_*_p (running sum +\\x, persistent pipelined scan)",
                         "algorithmic_bytes_per_launch": 16 * n, "ms_per_launch": scan_ms, "peak_source": peak_src},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": per_step_launches * args.steps, "clocks": clocks,
            "checks": checks,
            "primitives": prims,
        }
        print(json.dumps(line), file=_claim_stdout(), flush=True)
    if world > 1:
        td.barrier()
        td.destroy_process_group()


def prim_row(n, ms, bytes_per_elem, peak, **extra):
    gbs = bytes_per_elem * n / (ms * 1e-3) / 1e9
    row = {"n": n, "ms": ms, "elements_per_sec": n / (ms * 1e-3), "algorithmic_bytes_per_elem": bytes_per_elem,
           "achieved_gbs": gbs, "roofline_frac": gbs / peak}
    row.update(extra)
    if "model_bytes_per_elem" in extra:  # multi-pass algorithms: fraction of the roofline of the pass model's bytes
        row["model_roofline_frac"] = extra["model_bytes_per_elem"] * n / (ms * 1e-3) / 1e9 / peak
    return row


def synthetic_1brc_block(rows=1_000_000, n_names=10_000, seed=5):
    """`name;[-]d[d].d\n` lines in the shape of examples/1brc: n_names distinct stations of 3-19 bytes, temperatures
    N(10, 25) clipped to +-99.9.  -> (bytes, names list, station index per row, tenths per row)"""
    rng = np.random.default_rng(seed)
    letters = list("abcdefghijklmnopqrstuvwxyz ABCDEFGHIJK")
    names = set()
    while len(names) < n_names:
        names.add("".join(rng.choice(letters, int(rng.integers(3, 20)))).strip() or "q")
    names = sorted(names)
    ids = rng.integers(0, n_names, rows)
    tenths = np.clip(np.round(rng.normal(100, 250, rows)), -999, 999).astype(np.int64)
    block = "".join(f"{names[i]};{'-' if t < 0 else ''}{abs(int(t)) // 10}.{abs(int(t)) % 10}\n" for i, t in zip(ids, tenths)).encode()
    return block, names, ids, tenths


def primitive_table(be, ops, dist, time_kernel, dev, peak, world, rank):
    """Per-primitive elements/s and roofline fraction at the BASELINE.json sizes (1 GPU each rank)."""
    import torch
    from klongpy_b200 import DeviceArray
    out = {}
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    n = 100_000_000
    a = DeviceArray(torch.rand(n, dtype=torch.float64, device=dev, generator=g))
    b = DeviceArray(torch.rand(n, dtype=torch.float64, device=dev, generator=g))
    # `torch_eager_ms`: the same primitive through the ATen calls the reference's torch backend issues
    # (torch_backend.py:413-459, 612-616, 1058-1082) on the same GPU in the same run — the prior-art GPU path.
    # (`.tensor` launches a deferred elementwise result: with lazy evaluation on, a discarded `a+b` would never run)
    out["a+b"] = prim_row(n, time_kernel(lambda: ops.binary("add", a, b).tensor, 5), 24, peak,
                          torch_eager_ms=time_kernel(lambda: torch.add(a.tensor, b.tensor), 5))
    c = DeviceArray(torch.rand(n, dtype=torch.float64, device=dev, generator=g))
    fn, _ = be.compile_expr_ir(("binop", "%", ("binop", "*", ("var", "_v0"), ("binop", "+", ("literal", 2), ("binop", "*", ("var", "_v1"), ("literal", 3)))),
                                ("binop", "-", ("binop", "+", ("var", "_v0"), ("literal", 1)), ("binop", "*", ("var", "_v1"), ("var", "_v2")))), ["a", "b", "c"])
    ta, tb, tc = a.tensor, b.tensor, c.tensor
    out["(a*2+b*3)%(a+1)-(b*c)"] = prim_row(n, time_kernel(lambda: fn(a, b, c).tensor, 5), 32, peak,
                                            torch_eager_ms=time_kernel(lambda: (ta * (2 + (tb * 3))) / ((ta + 1) - (tb * tc)), 5))
    fn2, _ = be.compile_expr_ir(("reduce", "|", ("binop", "-", ("scan", "|", ("var", "_v0")), ("var", "_v0"))), ["a"])
    out["|/(|\\ts)-ts"] = prim_row(n, time_kernel(lambda: fn2(a), 5), 8, peak,
                                  torch_eager_ms=time_kernel(lambda: (ta.cummax(0).values - ta).amax(0), 5))
    del c
    keys_perm = DeviceArray(torch.randperm(n, device=dev, generator=g))
    # model bytes of the packed radix path (csrc/kgb_sort_packed.cuh): analysis read 8 + histogram read 8 + 16 per pass
    # (27 live bits -> 3 passes of 9 bits); the I/O floor stays 16 B/key (read keys, write int64 positions)
    out["grade_up_perm_1e8"] = prim_row(n, time_kernel(lambda: ops.argsort(keys_perm), 3), 16, peak,
                                        model_bytes_per_elem=16 + 16 * 3, passes=3,
                                        torch_eager_ms=time_kernel(lambda: torch.argsort(keys_perm.tensor, stable=True), 3))
    # (f) rank 1 of SURVEY.md §8f: gather `a@b` and join `a,b` on device arrays (what `x@<x` spends its time in once
    # the sort is fast).  A random 8-byte read moves a 32-byte sector: algorithmic 24 B/elem, sector traffic 48 B/elem.
    out["gather_perm_1e8"] = prim_row(n, time_kernel(lambda: ops.gather(a, keys_perm), 3), 24, peak, sector_bytes_per_elem=48,
                                      torch_eager_ms=time_kernel(lambda: a.tensor[keys_perm.tensor], 3))
    out["join_1e8+1e8"] = prim_row(2 * n, time_kernel(lambda: be.np.concatenate((a, b)), 3), 16, peak)
    # `x@<x` (sort by grade): the last radix pass emits the sorted keys itself, so there is no gather at all
    out["sort_by_grade_x@<x_perm_1e8"] = prim_row(n, time_kernel(lambda: ops.sort_values(keys_perm), 3), 16, peak,
                                                 torch_eager_ms=time_kernel(lambda: torch.sort(keys_perm.tensor, stable=True), 3))
    del keys_perm
    keys_10k = DeviceArray(torch.randint(0, 10_000, (n,), dtype=torch.int64, device=dev, generator=g))
    out["grade_up_10k_1e8"] = prim_row(n, time_kernel(lambda: ops.argsort(keys_10k), 3), 16, peak,
                                       model_bytes_per_elem=16 + 16 * 2, passes=2,
                                       torch_eager_ms=time_kernel(lambda: torch.argsort(keys_10k.tensor, stable=True), 3))
    out["group_10k_1e8"] = prim_row(n, time_kernel(lambda: ops.group(keys_10k), 3), 16, peak)
    out["find_1e8"] = prim_row(n, time_kernel(lambda: ops.find_equal(keys_10k, 77), 3), 8, peak,
                               torch_eager_ms=time_kernel(lambda: torch.nonzero(keys_10k.tensor == 77), 3))
    # `?a` reads 8 B per row and writes U values: 8 B/elem algorithmic (round 1 charged 16)
    out["range_10k_1e8"] = prim_row(n, time_kernel(lambda: ops.unique_first(keys_10k), 3), 8, peak,
                                    torch_eager_ms=time_kernel(lambda: torch.unique(keys_10k.tensor), 3),
                                    note="torch.unique returns sorted values, not first-appearance order")
    temp = DeviceArray(torch.randn(n, dtype=torch.float64, device=dev, generator=g))
    def check_groupagg(keys, vals, G):
        """The sharded tables against torch scatter-reduce / bincount of the same shard set, all-reduced: counts, minima
        and maxima exact, sums to 1e-12 — asserted on every rank of every world size."""
        mn, mx, sm, ct = (t.tensor for t in dist.sharded_groupagg(keys, vals, G))
        kt, vt = keys.tensor, vals.tensor
        rc = torch.bincount(kt, minlength=G)
        rmn = torch.full((G,), float("inf"), dtype=torch.float64, device=dev).scatter_reduce_(0, kt, vt, "amin")
        rmx = torch.full((G,), float("-inf"), dtype=torch.float64, device=dev).scatter_reduce_(0, kt, vt, "amax")
        rsm = torch.zeros(G, dtype=torch.float64, device=dev).scatter_add_(0, kt, vt)
        if world > 1:
            td_ = torch.distributed
            td_.all_reduce(rc)
            td_.all_reduce(rmn, op=td_.ReduceOp.MIN)
            td_.all_reduce(rmx, op=td_.ReduceOp.MAX)
            td_.all_reduce(rsm)
        ok = bool(torch.equal(ct, rc) and torch.equal(mn, rmn) and torch.equal(mx, rmx) and
                  torch.allclose(sm, rsm, rtol=1e-12, atol=1e-6))
        assert ok, f"rank {rank}: sharded grouped aggregation differs from the scatter-reduce witness"
        return "count/min/max exact, sum 1e-12 vs torch scatter-reduce of all shards, every rank"

    def over_ranks(ms):  # a sharded primitive is as slow as its slowest rank
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    # rows block-sharded over the ranks (n rows PER GPU, weak scaling); at N>1 the per-rank partial tables meet in one
    # owner-partitioned all-to-all + merge + all-gather (dist.sharded_groupagg); totals and peak are whole-job
    out["grouped_min_mean_max_10k_1e8"] = prim_row(
        n * world, over_ranks(time_kernel(lambda: dist.sharded_groupagg(keys_10k, temp, 10_000), 3)), 16, peak * world,
        n_gpus=world, rows_per_gpu=n, check=check_groupagg(keys_10k, temp, 10_000))
    if world > 1:
        out["grouped_min_mean_max_10k_1e8_alltoall"] = prim_row(
            n * world, over_ranks(time_kernel(lambda: dist.sharded_groupagg(keys_10k, temp, 10_000, exchange="alltoall"), 3)),
            16, peak * world, n_gpus=world, rows_per_gpu=n, note="hash-partitioned all_to_all exchange forced (auto picks one all_gather for 10 k stations)")
    if world == 1:
        # the same rows as two columns of a DeviceTable; `.gstats` = key range in one two-output
        # reduce (+8 B/row), the aggregation kernel, compaction of the live stations, mean = sum / count
        from klongpy_b200.table import DeviceTable
        tbl = DeviceTable({"station": keys_10k, "temp": temp})
        out["table_gstats_10k_1e8"] = prim_row(n, time_kernel(lambda: tbl.group_stats("station", "temp"), 3), 24, peak,
                                               note="DeviceTable.group_stats: select station, min, avg, max, count group by station")
        del tbl
    del temp, a, b
    # config 4 at its full size: 1e9 synthetic rows PER GPU, 10k stations (16 GB of input per GPU)
    nb = 1_000_000_000
    st_b = DeviceArray(torch.randint(0, 10_000, (nb,), dtype=torch.int64, device=dev, generator=g))
    tp_b = DeviceArray(torch.randn(nb, dtype=torch.float64, device=dev, generator=g))
    out["grouped_min_mean_max_10k_1e9"] = prim_row(
        nb * world, over_ranks(time_kernel(lambda: dist.sharded_groupagg(st_b, tp_b, 10_000), 3)), 16, peak * world,
        n_gpus=world, rows_per_gpu=nb, check=check_groupagg(st_b, tp_b, 10_000))
    if world > 1:
        # config 4 in BASELINE.json's own words: the hash-partitioned all_to_all exchange across the GPUs
        out["grouped_min_mean_max_10k_1e9_alltoall"] = prim_row(
            nb * world, over_ranks(time_kernel(lambda: dist.sharded_groupagg(st_b, tp_b, 10_000, exchange="alltoall"), 3)),
            16, peak * world, n_gpus=world, rows_per_gpu=nb,
            note="station s owned by rank s mod w: to_owners, ONE all_to_all, fold in rank order, ONE all_gather, from_owners")
    del tp_b
    if world == 1:
        # config 3 at its upper size: 1e9 int64 keys
        out["group_10k_1e9"] = prim_row(nb, time_kernel(lambda: ops.group(st_b), 2), 16, peak)
        out["find_1e9"] = prim_row(nb, time_kernel(lambda: ops.find_equal(st_b, 77), 3), 8, peak)
        out["range_10k_1e9"] = prim_row(nb, time_kernel(lambda: ops.unique_first(st_b), 2), 8, peak)
        del st_b
        kp = DeviceArray(torch.randperm(nb, device=dev, generator=g))
        out["grade_up_perm_1e9"] = prim_row(nb, time_kernel(lambda: ops.argsort(kp), 2), 16, peak,
                                            model_bytes_per_elem=16 + 16 * 4, passes=4)  # 30 live bits: 9+7+7+7
        del kp
    else:
        del st_b
    # config 1: examples/bench_compiler.kg expressions at the file's own size (100 K float64): µs per call through the
    # compiled callable (1000 back-to-back calls, one sync), next to the reference's NumPy codegen on this box's CPU
    import time as _time
    from oracle import kg_oracle as O
    rngc = np.random.default_rng(0)
    hc = {k: rngc.random(100_000) for k in ("a", "b", "c")}
    dc = {k: be.kg_asarray(v) for k, v in hc.items()}
    V = lambda k: ("var", k)  # noqa: E731
    exprs = {
        "a+b": (("binop", "+", V("_v0"), V("_v1")), ["a", "b"]),
        "a*2+b": (("binop", "*", V("_v0"), ("binop", "+", ("literal", 2), V("_v1"))), ["a", "b"]),
        "(a+b)*(a-b)": (("binop", "*", ("binop", "+", V("_v0"), V("_v1")), ("binop", "-", V("_v0"), V("_v1"))), ["a", "b"]),
        "+/a": (("reduce", "+", V("_v0")), ["a"]),
        "+/a*b": (("reduce", "+", ("binop", "*", V("_v0"), V("_v1"))), ["a", "b"]),
        "+\\a": (("scan", "+", V("_v0")), ["a"]),
        "|/(|\\a)-a": (("reduce", "|", ("binop", "-", ("scan", "|", V("_v0")), V("_v0"))), ["a"]),
        "(+/a*b)%+/b": (("binop", "%", ("reduce", "+", ("binop", "*", V("_v0"), V("_v1"))), ("reduce", "+", V("_v1"))), ["a", "b"]),
    }
    cfg1 = {}
    for name, (ir, params) in exprs.items():
        fnc, _ = be.compile_expr_ir(ir, params)
        dargs = [dc[k] for k in params]
        # (`.tensor` launches a deferred elementwise result: with lazy evaluation on, a discarded `a+b` would never run)
        for _ in range(20):
            fnc(*dargs).tensor
        torch.cuda.synchronize()
        t0 = _time.perf_counter()
        for _ in range(1000):
            fnc(*dargs).tensor
        torch.cuda.synchronize()
        us = (_time.perf_counter() - t0) * 1e3
        env = {f"_v{i}": hc[k] for i, k in enumerate(params)}
        O.eval_ir(ir, env)
        t0 = _time.perf_counter()
        for _ in range(50):
            O.eval_ir(ir, env)
        cfg1[name] = {"us_per_call": us, "numpy_us_per_call": (_time.perf_counter() - t0) / 50 * 1e6}
    out["config1_bench_compiler_100k"] = cfg1
    del dc
    # config 5: autograd at a 1e7-element batch (fused forward + fused backward kernels, float64)
    m = 10_000_000
    X = DeviceArray(torch.rand(m, dtype=torch.float64, device=dev, generator=g) * 2)
    Y = DeviceArray(torch.rand(m, dtype=torch.float64, device=dev, generator=g))

    def nn_loss(params):
        w, b = params
        s = 1 / (1 + ops.unary("exp", 0 - ((w * X) + b)))
        return ops.reduce("add", (s - Y) ** 2)

    from klongpy_b200 import lazy
    # default path: eager operators are deferred and fused (klongpy_b200/lazy.py, on by default) and the backend's own tape
    # records the launches — fused forward reduce + ONE two-output backward reduce for dL/dw1 and dL/db1
    out["autograd_sigmoid_nn_1e7"] = prim_row(m, time_kernel(lambda: be.compute_multi_autograd(nn_loss, [0.1, 0.1]), 5), 32, peak,
                                              note="operators deferred + fused (default); float64")
    lazy.enable(False)
    try:
        out["autograd_sigmoid_nn_1e7_eager"] = prim_row(m, time_kernel(lambda: be.compute_multi_autograd(nn_loss, [0.1, 0.1]), 3), 32, peak,
                                                        note="KGB200_LAZY=0: operator-by-operator tape, one kernel per Klong operator")
    finally:
        lazy.enable(True)
    R = DeviceArray(torch.rand(m, dtype=torch.float64, device=dev, generator=g) * 0.14 + 0.01)
    V = DeviceArray(torch.rand(m, dtype=torch.float64, device=dev, generator=g) * 0.35 + 0.05)
    W = DeviceArray(torch.full((m,), 1.0 / m, dtype=torch.float64, device=dev))
    sharpe_ir = ("binop", "%", ("reduce", "+", ("binop", "*", ("var", "_v0"), ("var", "_v1"))),
                 ("binop", "^", ("reduce", "+", ("binop", "*", ("binop", "^", ("var", "_v0"), ("literal", 2)),
                                                 ("binop", "^", ("var", "_v2"), ("literal", 2)))), ("literal", 0.5)))
    sharpe, _ = be.compile_expr_ir(sharpe_ir, ["x", "returns", "vols"])
    out["autograd_sharpe_grad_1e7"] = prim_row(m, time_kernel(lambda: be.compute_autograd(lambda x: sharpe(x, R, V), W), 5), 56, peak,
                                               note="compiled IR tree: fused forward and fused backward kernels")
    del X, Y, R, V, W
    keys_full = DeviceArray(torch.randint(-2**62, 2**62, (n,), dtype=torch.int64, device=dev, generator=g))
    # 64 live bits do not fit beside a 27-bit position: 4 packed passes over the top 37 bits + the prefix fix-up (16 B/key)
    out["grade_up_full64_1e8"] = prim_row(n, time_kernel(lambda: ops.argsort(keys_full), 3), 16, peak,
                                          model_bytes_per_elem=16 + 16 * 4 + 16, passes=4)
    del keys_full
    if world == 1:
        # 1brc text -> (station id, float64 temperature) columns + dictionary on the device.  A
        # 1e6-row block tiled 100x (1e8 rows, ~1.7 GB of text); algorithmic bytes per row = its text + 16 B of output.
        from klongpy_b200 import ingest
        block, names, _, _ = synthetic_1brc_block()
        reps = 100
        text = torch.from_numpy(np.frombuffer(block, dtype=np.uint8).copy()).to(dev).repeat(reps)
        rows = 1_000_000 * reps
        bpr = text.numel() / rows + 16
        got = ingest.parse(text)
        assert got[0].size == rows and got[2] == sorted(names, key=lambda s: s.encode())
        del got
        ms_kernels = time_kernel(lambda: ingest._parse_arrival(text, dev, fetch_names=False), 5)
        # the reference's line loop (oracle.parse_1brc = par.py:57-88 restated) on this box's CPU, 200k-row sample
        sample = block[:block.index(b"\n", len(block) // 5) + 1]
        t0 = _time.perf_counter()
        O.parse_1brc(sample)
        cpu_rows_per_sec = sample.count(b"\n") / (_time.perf_counter() - t0)
        out["ingest_1brc_parse_1e8"] = prim_row(rows, time_kernel(lambda: ingest.parse(text), 3), bpr, peak,
                                                device_ms=ms_kernels, device_roofline_frac=bpr * rows / (ms_kernels * 1e-3) / 1e9 / peak,
                                                cpu_port_rows_per_sec=cpu_rows_per_sec,
                                                note="ms: ingest.parse() incl. the host-side dictionary (10k names fetched, sorted, decoded) "
                                                     "and the id remap gather; device_ms: kgb_brc_scan + kgb_brc_parse only")
        out["ingest_1brc_aggregate_1e8"] = prim_row(rows, time_kernel(lambda: ingest.aggregate(text), 3), bpr, peak,
                                                    note="text -> min/mean/max per station, names ascending (examples/1brc end to end)")
        del text
        # the same file shape at config 4's row count: 1e9 rows, 17 GB of text (the dictionary work on the host is per
        # station, not per row, so the API-level number converges to the device number)
        reps = 1000
        text = torch.from_numpy(np.frombuffer(block, dtype=np.uint8).copy()).to(dev).repeat(reps)
        rows = 1_000_000 * reps
        ms_kernels = time_kernel(lambda: ingest._parse_arrival(text, dev, fetch_names=False), 3)
        out["ingest_1brc_parse_1e9"] = prim_row(rows, time_kernel(lambda: ingest.parse(text), 2), bpr, peak,
                                                device_ms=ms_kernels, device_roofline_frac=bpr * rows / (ms_kernels * 1e-3) / 1e9 / peak)
        out["ingest_1brc_aggregate_1e9"] = prim_row(rows, time_kernel(lambda: ingest.aggregate(text), 2), bpr, peak)
        del text
    return out


if __name__ == "__main__":
    main()
