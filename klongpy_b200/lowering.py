"""Compiler-IR -> fused kernel stages  (the backend half of klongpy/compiler.py).

The reference lowers the tuple IR of `compiler.py:37-113` to a Python source string of unfused NumPy calls
(`numpy_backend.py:105-178`): every node is a pass over memory with a temporary.  Here the tree becomes a small
DAG of *stages* (SURVEY.md Appendix C scheduling rule):

  * every `reduce` / `scan` node is one kernel whose argument sub-tree is its fused prologue,
  * a `reduce` whose argument contains exactly one `scan` (max-drawdown, `|/(|\\ts)-ts`) becomes ONE
    scan+epilogue+reduce kernel (8 B/elem instead of 8+16+16),
  * the enclosing elementwise tree is one more fused MAP kernel that reads stage results as 0-d device
    scalars (never synced to the host) or arrays.

Parameter order = first reference order = `BackendProvider._collect_params` (base.py:345-365) = var_syms.
Anything the kernels do not support RAISES, so the interpreter's silent fallback
(interpreter.py:697-701) takes the eager path instead of mis-computing.
"""
import numpy as np

from . import _lib, autodiff, ops
from .array import DeviceArray

_BIN = {"+": "add", "-": "sub", "*": "mul", "%": "div", "^": "pow"}
_CMP = {"=": "eq", ">": "gt", "<": "lt"}


def collect_params(ir):
    params, seen = [], set()

    def walk(node):
        t = node[0]
        if t == "var":
            if node[1] not in seen:
                seen.add(node[1])
                params.append(node[1])
        elif t in ("binop", "cmp"):
            walk(node[2])
            walk(node[3])
        elif t == "negate":
            walk(node[1])
        elif t in ("reduce", "scan"):
            walk(node[2])
    walk(ir)
    return params


def _count(node, kinds):
    t = node[0]
    c = 1 if t in kinds else 0
    if t in ("binop", "cmp"):
        return c + _count(node[2], kinds) + _count(node[3], kinds)
    if t == "negate":
        return c + _count(node[1], kinds)
    if t in ("reduce", "scan"):
        return c + _count(node[2], kinds)
    return c


def _validate(node):
    t = node[0]
    if t == "literal":
        if type(node[1]) not in (int, float, bool):
            raise TypeError("literal must be int or float")
    elif t == "var":
        pass
    elif t == "binop":
        if node[1] not in _BIN:
            raise ValueError(f"unsupported binop {node[1]!r}")
        _validate(node[2])
        _validate(node[3])
    elif t == "cmp":
        if node[1] not in _CMP:
            raise ValueError(f"unsupported cmp {node[1]!r}")
        _validate(node[2])
        _validate(node[3])
    elif t == "negate":
        _validate(node[1])
    elif t in ("reduce", "scan"):
        if node[1] not in ops.RED:
            raise ValueError(f"unsupported fold op {node[1]!r}")
        _validate(node[2])
    else:
        raise ValueError(f"unknown IR node {t!r}")


class _Emitter:
    """Builds one stage's postfix code; maps global value slots to the stage's local argument indices."""

    def __init__(self):
        self.code = []
        self.slots = []  # global slot ids in local-arg order

    def arg(self, slot):
        if slot not in self.slots:
            self.slots.append(slot)
        self.code += [ops.OP_ARG, self.slots.index(slot)]


class Stage:
    __slots__ = ("mode", "red", "red2", "code", "code2", "slots", "out_slot", "pair", "paired_with")

    def __init__(self, mode, red, red2, code, code2, slots, out_slot):
        self.mode, self.red, self.red2 = mode, red, red2
        self.code, self.code2, self.slots, self.out_slot = tuple(code), tuple(code2) if code2 else None, slots, out_slot
        self.pair = None         # (second stage, its code re-indexed onto the merged slot list, merged slots)
        self.paired_with = None  # set on the second stage of a pair


class Plan:
    """Stage list for one IR tree.  Slots 0..n_params-1 are the call arguments; each stage adds one slot."""

    def __init__(self, ir):
        _validate(ir)
        self.params = collect_params(ir)
        self.n_slots = len(self.params)
        self.stages = []
        em = _Emitter()
        self._emit(ir, em, None)
        if len(em.code) == 2 and em.code[0] == ops.OP_ARG and em.slots[0] >= len(self.params):
            self.result_slot = em.slots[0]  # the tree's root was itself a stage
        else:
            self.result_slot = self._add_stage(_lib.MODE_MAP, 0, 0, em.code, None, em.slots)
        self._pair_reductions()

    def _pair_reductions(self):
        """Two reductions that read only call arguments (`(+/p*s)%+/s`, VWAP) become ONE two-output reduce kernel:
        the shared operands stream from HBM once.  The pairing is undone at call time if the operands turn out to
        have different lengths or a narrow accumulator (make_callable)."""
        n_params = len(self.params)
        reds = [st for st in self.stages if st.mode == _lib.MODE_REDUCE and st.pair is None and
                all(sl < n_params for sl in st.slots)]
        for a, b in zip(reds[0::2], reds[1::2]):
            slots = list(a.slots)
            remap = {}
            for j, sl in enumerate(b.slots):
                if sl not in slots:
                    slots.append(sl)
                remap[j] = slots.index(sl)
            code_b, i = [], 0
            cb = list(b.code)
            while i < len(cb):
                w = cb[i]
                if w == ops.OP_ARG:
                    code_b += [ops.OP_ARG, remap[cb[i + 1]]]
                    i += 2
                elif w in (ops.OP_LIT_I64, ops.OP_LIT_F64):
                    code_b += cb[i:i + 3]
                    i += 3
                else:
                    code_b.append(w)
                    i += 1
            a.pair = (b, tuple(code_b), slots)
            b.paired_with = a

    def _add_stage(self, mode, red, red2, code, code2, slots):
        slot = self.n_slots
        self.n_slots += 1
        self.stages.append(Stage(mode, red, red2, code, code2, list(slots), slot))
        return slot

    def _emit(self, node, em, scan_hole):
        t = node[0]
        if t == "literal":
            em.code += ops.lit(node[1])
        elif t == "var":
            em.arg(self.params.index(node[1]))
        elif t == "binop":
            self._emit(node[2], em, scan_hole)
            self._emit(node[3], em, scan_hole)
            em.code.append(ops.OPC[_BIN[node[1]]])
        elif t == "cmp":
            self._emit(node[2], em, scan_hole)
            self._emit(node[3], em, scan_hole)
            em.code.append(ops.OPC[_CMP[node[1]]])
        elif t == "negate":
            self._emit(node[1], em, scan_hole)
            em.code.append(ops.OPC["neg"])
        elif t == "scan" and scan_hole is not None and node is scan_hole:
            em.code.append(ops.OP_SCANVAL)
        elif t == "reduce":
            arg = node[2]
            if _count(arg, ("scan",)) == 1 and _count(arg, ("reduce",)) == 0:
                scan_node = _find(arg, "scan")
                if _count(scan_node[2], ("scan", "reduce")) == 0:
                    # fused scan -> epilogue -> reduce
                    sub = _Emitter()
                    self._emit(scan_node[2], sub, None)
                    epi = _Emitter()
                    epi.slots = sub.slots  # shared argument table
                    self._emit(arg, epi, scan_node)
                    slot = self._add_stage(_lib.MODE_SCANRED, ops.RED[scan_node[1]], ops.RED[node[1]], sub.code,
                                           epi.code, epi.slots)
                    em.arg(slot)
                    return
            sub = _Emitter()
            self._emit(arg, sub, None)
            slot = self._add_stage(_lib.MODE_REDUCE, ops.RED[node[1]], 0, sub.code, None, sub.slots)
            em.arg(slot)
        elif t == "scan":
            sub = _Emitter()
            self._emit(node[2], sub, None)
            slot = self._add_stage(_lib.MODE_SCAN, ops.RED[node[1]], 0, sub.code, None, sub.slots)
            em.arg(slot)
        else:
            raise ValueError(f"unknown IR node {t!r}")


def _find(node, kind):
    if node[0] == kind:
        return node
    t = node[0]
    kids = (node[2], node[3]) if t in ("binop", "cmp") else (node[1],) if t == "negate" else \
        (node[2],) if t in ("reduce", "scan") else ()
    for k in kids:
        r = _find(k, kind)
        if r is not None:
            return r
    return None


def _is_arraylike(v):
    return isinstance(v, DeviceArray) and v.ndim > 0


def make_callable(ir):
    """Return fn(*args) executing the plan on device arrays.  Raises on unsupported argument kinds."""
    plan = Plan(ir)
    n_params = len(plan.params)
    # a plan of ONE kernel whose operands are exactly the call arguments (a+b, a*2+b, +/a*b, +\a ... — five of the six
    # rows of examples/bench_compiler.kg): the pre-resolved launcher takes the call when every argument is a plain 1-d
    # device vector or a Python number; everything else (and every other plan) goes through the general loop below
    fast = None
    if len(plan.stages) == 1:
        st0 = plan.stages[0]
        if st0.pair is None and st0.code2 is None and list(st0.slots) == list(range(n_params)) and \
                st0.mode in (_lib.MODE_MAP, _lib.MODE_REDUCE, _lib.MODE_SCAN):
            fast = ops.FastStage(st0.code, st0.mode, st0.red)
    # `(+/p*s)%+/s`: two folds that share one pass + an elementwise tree over exactly those two results -> ONE launch
    # (kgb_expr_compile_epi: the last CTA of the two-output fold evaluates the epilogue)
    fast_epi = None
    if len(plan.stages) == 3:
        sa, sb, sc = plan.stages
        if sa.pair is not None and sa.pair[0] is sb and sc.mode == _lib.MODE_MAP and sc.code2 is None and \
                sc.out_slot == plan.result_slot and sc.slots and set(sc.slots) <= {sa.out_slot, sb.out_slot}:
            remap = [0 if sl == sa.out_slot else 1 for sl in sc.slots]
            code3, i, src = [], 0, list(sc.code)
            while i < len(src):
                w = src[i]
                if w == ops.OP_ARG:
                    code3 += [w, remap[src[i + 1]]]
                    i += 2
                elif w in (ops.OP_LIT_I64, ops.OP_LIT_F64):
                    code3 += src[i:i + 3]
                    i += 3
                else:
                    code3.append(w)
                    i += 1
            fast_epi = (ops.FastStage(sa.code, _lib.MODE_REDUCE, sa.red, sa.pair[1], sb.red, code3), sa.pair[2])
    # inside the general loop every plain stage (and every paired fold) tries its own pre-resolved launcher first
    stage_fast = {}
    for st in plan.stages:
        if st.pair is not None:
            stage_fast[id(st)] = ops.FastStage(st.code, _lib.MODE_REDUCE, st.red, st.pair[1], st.pair[0].red)
        elif st.code2 is None and st.mode in (_lib.MODE_MAP, _lib.MODE_REDUCE, _lib.MODE_SCAN):
            stage_fast[id(st)] = ops.FastStage(st.code, st.mode, st.red)

    def fn(*args):
        if len(args) != n_params:
            raise TypeError(f"compiled expression takes {n_params} arguments ({len(args)} given)")
        if fast is not None:
            r = fast(args)
            if r is not NotImplemented:
                return r
        if fast_epi is not None:
            try:
                r = fast_epi[0]([args[sl] for sl in fast_epi[1]])
            except _lib.KgbError:
                r = NotImplemented   # a narrow accumulator: the general path below sorts it out
            if r is not NotImplemented:
                return r
        vals = list(args) + [None] * (plan.n_slots - n_params)
        any_dev = False
        nd_shape = None
        for v in args:
            if isinstance(v, DeviceArray):
                any_dev = True
                if v.ndim > 1:
                    # N-d operands (numpy_backend.py:163-173): elementwise trees keep the shape, `np.add.reduce` & co fold
                    # axis 0, `np.cumsum` / `np.cumprod` scan the FLATTENED array (the reference's compiled semantics)
                    if nd_shape is not None and v.shape != nd_shape:
                        raise NotImplementedError("compiled kernels take N-d operands of one shape")
                    nd_shape = v.shape
            elif type(v) in (int, float) or isinstance(v, (np.integer, np.floating)):
                pass
            else:
                raise TypeError(f"unsupported compiled-expression operand {type(v).__name__}")
        if not any_dev:
            raise NotImplementedError("no device operand: scalar expressions stay in the interpreter")
        for st in plan.stages:
            if vals[st.out_slot] is not None:
                continue  # second half of a pair, already produced by the two-output kernel
            if st.pair is not None and nd_shape is None:
                other, code_b, slots = st.pair
                pargs = [vals[s] for s in slots]
                try:
                    r = stage_fast[id(st)](pargs)
                except _lib.KgbError:
                    r = NotImplemented   # a narrow accumulator: the general path below sorts it out
                if r is not NotImplemented:
                    vals[st.out_slot], vals[other.out_slot] = r
                    continue
                lens = {a.size for a in pargs if _is_arraylike(a)}
                if len(lens) == 1 and 0 not in lens:
                    try:
                        vals[st.out_slot], vals[other.out_slot] = ops.reduce2(st.code, st.red, code_b, other.red, pargs)
                        continue
                    except (_lib.KgbError, autodiff.NotDifferentiable):
                        pass  # a narrow accumulator, or a fold without a gradient rule: two separate reductions
            sargs = [vals[s] for s in st.slots]
            if nd_shape is None:
                fs = stage_fast.get(id(st))
                if fs is not None:
                    r = fs(sargs)
                    if r is not NotImplemented:
                        vals[st.out_slot] = r
                        continue
            if nd_shape is not None and any(isinstance(a, DeviceArray) and a.ndim > 1 for a in sargs):
                if any(isinstance(a, DeviceArray) and a.ndim == 1 for a in sargs):
                    raise NotImplementedError("compiled kernels do not broadcast 1-d against N-d operands")
                if st.mode == _lib.MODE_REDUCE:
                    vals[st.out_slot] = ops.reduce_axis0(st.code, sargs, st.red)
                    continue
                if st.mode == _lib.MODE_SCAN:
                    if st.red not in (_lib.RED_ADD, _lib.RED_MUL):
                        raise NotImplementedError("|\\ and &\\ on N-d operands: interpreter semantics (row by row)")
                    sargs = [a.flatten() if isinstance(a, DeviceArray) and a.ndim > 1 else a for a in sargs]
                elif st.mode == _lib.MODE_SCANRED:
                    raise NotImplementedError("fused scan+reduce takes 1-d operands")
            has_array = any(_is_arraylike(a) for a in sargs)
            if st.mode != _lib.MODE_MAP and not has_array:
                raise NotImplementedError("fold over a scalar: interpreter semantics apply")
            if st.mode in (_lib.MODE_REDUCE, _lib.MODE_SCANRED):
                n = next(a for a in sargs if _is_arraylike(a)).size
                if n == 0:
                    # np.add.reduce([]) -> 0.0 and np.multiply.reduce([]) -> 1.0 in the result dtype;
                    # np.maximum.reduce([]) raises (numpy_backend.py:163) -> interpreter fallback
                    if st.mode == _lib.MODE_REDUCE and st.red in (_lib.RED_ADD, _lib.RED_MUL):
                        vals[st.out_slot] = ops.empty_fold(st.code, sargs, st.red)
                        continue
                    raise ValueError("zero-size array to reduction operation which has no identity")
            if not any(isinstance(a, DeviceArray) for a in sargs):
                raise NotImplementedError("stage without device operands")
            # the plan's final scalar epilogue (`(+/p*s)%+/s`: a divide of two 0-d results) launches at once: deferring a
            # 0-d result that the caller reads next only adds bookkeeping
            final_scalar = st.out_slot == plan.result_slot and st.mode == _lib.MODE_MAP and not has_array
            vals[st.out_slot] = ops.run(st.code, sargs, mode=st.mode, red=st.red, code2=st.code2, red2=st.red2,
                                        _no_defer=final_scalar)
        return vals[plan.result_slot]

    fn.plan = plan
    return fn
