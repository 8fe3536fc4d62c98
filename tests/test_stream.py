"""Host-resident vectors streamed through the fused kernels (klongpy_b200/stream.py): chunked H2D | reduce + carried
scan | D2H must equal the oracle's one-shot result bit for bit on dyadic data, for any chunking."""
import numpy as np
import pytest

from oracle import kg_oracle as O

RED_IR = ("reduce", "+", ("binop", "^", ("binop", "+", ("binop", "*", ("var", "_v0"), ("literal", 2)), ("literal", 1)),
                          ("literal", 2)))
SCAN_IR = ("scan", "+", ("var", "_v0"))


@pytest.mark.gpu
@pytest.mark.parametrize("n,chunk", [(1000, 4096), (100_003, 4096), (1_000_003, 65_536), (3_000_000, 1 << 20)])
def test_streamed_reduce_and_scan_matches_oracle(cuda_backend, n, chunk):
    import torch
    from klongpy_b200 import ops, stream
    rng = np.random.default_rng(n)
    x = rng.integers(0, 256, n) / 256.0
    hx = torch.from_numpy(x).pin_memory()
    hs = torch.empty(n, dtype=torch.float64).pin_memory()
    red_code = [ops.OP_ARG, 0] + ops.lit(2) + [ops.OPC["mul"]] + ops.lit(1) + [ops.OPC["add"]] + ops.lit(2) + [ops.OPC["pow"]]
    st = stream.HostStream("cuda:0", chunk=chunk)
    for _ in range(2):  # second call reuses the staging buffers
        r = st.reduce_and_scan(hx, red_code, "+", [ops.OP_ARG, 0], "+", hs)
        torch.cuda.synchronize()
        assert r.item() == O.eval_ir(RED_IR, {"_v0": x})
        assert np.array_equal(hs.numpy(), O.eval_ir(SCAN_IR, {"_v0": x}))
        hs.zero_()
