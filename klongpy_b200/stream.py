"""Host-resident vectors streamed through the fused kernels (the end-to-end path for data that starts and ends in
host memory — the situation of every KlongPy program whose arrays come from `.pyf`/NumPy, `tests/perf_vector.py`).

The vector is cut into chunks; chunk k+1 is copied host→device while chunk k is reduced / scanned and the scan of
chunk k-1 is copied device→host, on three CUDA streams, so a step costs max(H2D, compute, D2H) instead of their
sum (PCIe is full duplex).  The running sum carries across chunks exactly like it carries across GPUs in
`dist.sharded_scan`: the last element of chunk k's scan is the seed (`KGB_HAS_CARRY`) of chunk k+1, read on the
device — no host round trip inside the loop.  Reductions fold per-chunk partials on the device in chunk order
(deterministic).
"""
import torch

from . import _lib, ops
from .array import DeviceArray

_RED_NAME = {_lib.RED_ADD: "add", _lib.RED_MUL: "mul", _lib.RED_MAX: "max", _lib.RED_MIN: "min"}


class HostStream:
    """Reusable device staging (3 input + 3 output chunk buffers) for streaming one dtype / chunk size."""

    def __init__(self, device, chunk=1 << 26, dtype=torch.float64, depth=3):
        self.device = torch.device(device)
        self.chunk = int(chunk)
        self.depth = depth
        self.inb = [torch.empty(self.chunk, dtype=dtype, device=self.device) for _ in range(depth)]
        self.s_in = torch.cuda.Stream(self.device)
        self.s_cmp = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)

    def reduce_and_scan(self, host_x, reduce_code, reduce_red, scan_code, scan_red, host_scan_out):
        """One pass over pinned `host_x`: returns the 0-d device fold of `reduce_code` and writes the inclusive scan of
        `scan_code` into pinned `host_scan_out`.  Both programs take the chunk as their only array argument."""
        n = host_x.numel()
        red = ops.RED[reduce_red] if not isinstance(reduce_red, int) else reduce_red
        sred = ops.RED[scan_red] if not isinstance(scan_red, int) else scan_red
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        nchunks = (n + self.chunk - 1) // self.chunk
        in_free = [None] * self.depth   # event: compute finished reading input buffer b
        acc = None
        carry = None
        for k in range(nchunks):
            b = k % self.depth
            lo, hi = k * self.chunk, min(n, (k + 1) * self.chunk)
            with torch.cuda.stream(self.s_in):
                if in_free[b] is not None:
                    self.s_in.wait_event(in_free[b])
                xin = self.inb[b][: hi - lo]
                xin.copy_(host_x[lo:hi], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self.s_in)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(ev_in)
                X = DeviceArray(xin)
                part = ops.run(reduce_code, [X], mode=_lib.MODE_REDUCE, red=red)
                acc = part if acc is None else ops.binary(_RED_NAME[red], acc, part)
                if carry is None:
                    sc = ops.run(scan_code, [X], mode=_lib.MODE_SCAN, red=sred)
                else:
                    sc = ops.run(scan_code, [X, carry], mode=_lib.MODE_SCAN, red=sred, code2=[ops.OP_ARG, 1])
                carry = DeviceArray(sc._t[-1])
                ev_c = torch.cuda.Event()
                ev_c.record(self.s_cmp)
                in_free[b] = ev_c
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_c)
                host_scan_out[lo:hi].copy_(sc._t, non_blocking=True)
                sc._t.record_stream(self.s_out)  # the caching allocator must not recycle it before the copy is done
        for s in (self.s_in, self.s_cmp, self.s_out):
            cur.wait_stream(s)
        return acc

    def sharded_reduce_and_scan(self, host_x, reduce_code, reduce_red, scan_code, scan_red, host_scan_out):
        """The same step for one rank's shard of a vector block-sharded over `torch.distributed` ranks (pinned host in,
        pinned host out).  The scan of a shard needs every lower rank's total first, so the step has two phases:
          1. chunks go host -> device into a RESIDENT copy of the shard; each chunk is folded as it lands by ONE
             two-output reduce (partial of `reduce_code` + total of `scan_code`), partials combine on the device in
             chunk order; then ONE all_gather of 16 bytes per rank (dist.exchange_fold_and_offset);
          2. the resident shard is scanned chunk by chunk, seeded with the rank's exclusive offset and then with each
             chunk's last element, while the previous chunk's scan goes device -> host.
        Returns the global fold (0-d device value, identical on every rank)."""
        from . import dist
        n = host_x.numel()
        red = ops.RED[reduce_red] if not isinstance(reduce_red, int) else reduce_red
        sred = ops.RED[scan_red] if not isinstance(scan_red, int) else scan_red
        if getattr(self, "resident", None) is None or self.resident.numel() < n or self.resident.dtype != host_x.dtype:
            self.resident = torch.empty(n, dtype=host_x.dtype, device=self.device)
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        nchunks = (n + self.chunk - 1) // self.chunk
        acc = tot = None
        for k in range(nchunks):
            lo, hi = k * self.chunk, min(n, (k + 1) * self.chunk)
            with torch.cuda.stream(self.s_in):
                xin = self.resident[lo:hi]
                xin.copy_(host_x[lo:hi], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self.s_in)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(ev_in)
                part, total = ops.reduce2(reduce_code, red, scan_code, sred, [DeviceArray(xin)])
                acc = part if acc is None else ops.binary(_RED_NAME[red], acc, part)
                tot = total if tot is None else ops.binary(_RED_NAME[sred], tot, total)
        with torch.cuda.stream(self.s_cmp):
            fold, carry = dist.exchange_fold_and_offset(acc, tot, red, sred)  # NCCL runs on the compute stream
            for k in range(nchunks):
                lo, hi = k * self.chunk, min(n, (k + 1) * self.chunk)
                X = DeviceArray(self.resident[lo:hi])
                if carry is None:
                    sc = ops.run(scan_code, [X], mode=_lib.MODE_SCAN, red=sred)
                else:
                    sc = ops.run(scan_code, [X, carry], mode=_lib.MODE_SCAN, red=sred, code2=[ops.OP_ARG, 1])
                carry = DeviceArray(sc._t[-1])
                ev_c = torch.cuda.Event()
                ev_c.record(self.s_cmp)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(ev_c)
                    host_scan_out[lo:hi].copy_(sc._t, non_blocking=True)
                    sc._t.record_stream(self.s_out)
        for s in (self.s_in, self.s_cmp, self.s_out):
            cur.wait_stream(s)
        return fold
