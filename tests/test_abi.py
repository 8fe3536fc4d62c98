"""The C-ABI library builds for sm_100a, loads without a GPU, exports every symbol include/kgb200.h declares,
and its code generator + NVRTC produce a cubin for every IR tree of the golden set (no GPU needed)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, to_ir


def header_symbols():
    text = open(os.path.join(ROOT, "include", "kgb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kgb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built_lib):
    from klongpy_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(built_lib, s), f"libkgb200.so does not export {s}"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in klongpy_b200/_lib.py"
    for s in _lib.SIGNATURES:
        assert s in syms, f"{s} is bound by ctypes but not declared in include/kgb200.h"
    assert built_lib.kgb_abi_version() == 1


def test_compute_without_gpu_fails_loudly(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from klongpy_b200 import _lib
    assert built_lib.kgb_init(0) != 0
    assert b"CUDA" in built_lib.kgb_last_error() or b"device" in built_lib.kgb_last_error()
    from klongpy_b200 import ops
    saved = _lib._host_double
    _lib._host_double = None
    try:
        with pytest.raises(_lib.KgbError):
            ops.default_device.__wrapped__() if hasattr(ops.default_device, "__wrapped__") else _raise_default()
    finally:
        _lib._host_double = saved


def _raise_default():
    from klongpy_b200 import ops
    old = ops._default_device
    ops._default_device = None
    try:
        ops.default_device()
    finally:
        ops._default_device = old


def _check(lib, code, code2, kinds, dtypes, mode, red, red2):
    c = (ctypes.c_int32 * len(code))(*code)
    c2 = (ctypes.c_int32 * len(code2))(*code2) if code2 else None
    k = (ctypes.c_int32 * max(len(kinds), 1))(*kinds)
    d = (ctypes.c_int32 * max(len(dtypes), 1))(*dtypes)
    od = ctypes.c_int32()
    src = ctypes.c_char_p()
    rc = lib.kgb_expr_check(c, len(code), c2, len(code2) if code2 else 0, k, d, len(kinds), mode, red, red2,
                            ctypes.byref(od), ctypes.byref(src))
    assert rc == 0, lib.kgb_last_error().decode()
    return od.value, src.value.decode()


def test_codegen_compiles_every_golden_ir(built_lib, golden):
    """IR -> stage plan -> CUDA source -> NVRTC cubin for sm_100a, for every compiled golden case."""
    from klongpy_b200 import _lib, lowering
    arrays = golden["compiled_arrays"]
    kcode = {np.dtype("float64"): _lib.F64, np.dtype("int64"): _lib.I64}
    n_kernels = 0
    for case in golden["compiled"]:
        plan = lowering.Plan(to_ir(case["ir"]))
        slots = []
        for p in case["params"]:
            if p in ("k",):
                slots.append((_lib.ARG_SCALAR, _lib.I64))
            elif p in ("f",):
                slots.append((_lib.ARG_SCALAR, _lib.F64))
            else:
                slots.append((_lib.ARG_ARRAY, kcode[arrays["in_" + p].dtype]))
        for st in plan.stages:
            kinds = [slots[s][0] for s in st.slots]
            dts = [slots[s][1] for s in st.slots]
            od, src = _check(built_lib, list(st.code), list(st.code2) if st.code2 else None, kinds, dts, st.mode,
                             st.red, st.red2)
            assert "KGB_EVAL" in src
            slots.append((_lib.ARG_DEVSCALAR if st.mode in (_lib.MODE_REDUCE, _lib.MODE_SCANRED) else _lib.ARG_ARRAY, od))
            n_kernels += 1
        # result dtype of the last stage must match the reference's
        want = np.dtype(case["out_dtype"])
        assert slots[plan.result_slot][1] == kcode[want], (case["src"], slots[plan.result_slot], want)
    assert n_kernels >= len(golden["compiled"])


def test_headline_kernels_have_no_spills_and_vectorise(built_lib, tmp_path):
    """-Xptxas-style evidence without a GPU: the fused reduce / scan cubins use 128/256-bit global accesses and no
    local memory (checked with cuobjdump on the NVRTC output)."""
    import shutil
    import subprocess
    from klongpy_b200 import _lib, ops
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    os.environ["KGB200_DUMP_DIR"] = str(tmp_path)
    try:
        code = [ops.OP_ARG, 0] + ops.lit(2) + [ops.OPC["mul"]] + ops.lit(1) + [ops.OPC["add"]] + ops.lit(2) + [ops.OPC["pow"]]
        _check(built_lib, code, None, [0], [_lib.F64], _lib.MODE_REDUCE, _lib.RED_ADD, 0)
        _check(built_lib, [ops.OP_ARG, 0], None, [0], [_lib.F64], _lib.MODE_SCAN, _lib.RED_ADD, 0)
    finally:
        del os.environ["KGB200_DUMP_DIR"]
    cubins = [f for f in os.listdir(tmp_path) if f.endswith(".cubin")]
    assert len(cubins) == 2
    for f in cubins:
        path = os.path.join(tmp_path, f)
        res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
        assert "LOCAL:0" in res and "STACK:0" in res, res
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        # reduce: 256-bit LDG (sm_100); scan: 128-bit LDG in the tile kernel + 16-byte cp.async in the pipelined one
        assert re.search(r"LDG\.E[.A-Z0-9]*\.(128|256)", sass), "no wide global loads"
        # -fmad=false: `(x*2)+1` must round twice like NumPy, so neither kernel may contain a fused multiply-add
        assert "DFMA" not in sass, "FMA contraction is on: a*2+1 would round once instead of twice"


def test_groupagg_workspace_scales_with_the_tables_only(built_lib):
    """Sparse key spaces (table.group_stats passes G = max key + 1) must not reserve the per-CTA partial tables the
    shared-memory variants use: ~32 B per group beyond 12 800 groups, not ~1920 B."""
    ws = ctypes.c_size_t()
    assert built_lib.kgb_groupagg_workspace_bytes(1_000_000, 10_000, ctypes.byref(ws)) == 0
    small = ws.value
    assert small >= 160 * 10_000 * 12          # per-CTA partial sums + counts exist for 10 k stations
    assert built_lib.kgb_groupagg_workspace_bytes(1_000_000, 4_000_000, ctypes.byref(ws)) == 0
    assert ws.value <= 4_000_000 * 32 + 4096   # four 8-byte tables, nothing per CTA
    assert built_lib.kgb_groupagg_workspace_bytes(1_000_000, 12_801, ctypes.byref(ws)) == 0
    assert ws.value <= 12_801 * 32 + 4096


def test_floor_divide_and_remainder_programs_type_check(built_lib):
    """`//` and `%` of the array type have integer kernels (no float64 detour): int op int stays int64."""
    from klongpy_b200 import _lib, ops
    for name in ("floordiv", "mod"):
        code = [ops.OP_ARG, 0, ops.OP_ARG, 1, ops.OPC[name]]
        c = (ctypes.c_int32 * len(code))(*code)
        k = (ctypes.c_int32 * 2)(_lib.ARG_ARRAY, _lib.ARG_ARRAY)
        for dts, want in (((_lib.I64, _lib.I64), _lib.I64), ((_lib.I64, _lib.F64), _lib.F64), ((_lib.F64, _lib.F64), _lib.F64)):
            d = (ctypes.c_int32 * 2)(*dts)
            od = ctypes.c_int32()
            src = ctypes.c_char_p()
            rc = built_lib.kgb_expr_check(c, len(code), None, 0, k, d, 2, _lib.MODE_MAP, _lib.RED_ADD, 0, ctypes.byref(od),
                                          ctypes.byref(src))
            assert rc == 0, built_lib.kgb_last_error()
            assert od.value == want, (name, dts, od.value)
            assert (b"kgb_floordiv" if name == "floordiv" else b"kgb_mod") in src.value
