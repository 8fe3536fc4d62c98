"""Symbolic differentiation of postfix expression programs.

The reference differentiates by running the interpreter on `requires_grad` torch tensors: one ATen kernel per
operator forward and one per operator backward, every intermediate materialised (`torch_backend.py:692-782`,
SURVEY.md §8 a14).  Here a fused expression program f(args) is differentiated *as a program*: for every
argument i the tangent expression t_i = df/d(arg_i) is derived on the expression tree, simplified, and emitted as
another postfix program over (args..., g), so the backward pass of a fused kernel is itself ONE fused kernel per
differentiated argument that recomputes what it needs from the inputs (no saved intermediates).

Trees: ('arg', i) | ('lit', v) | ('op', name, child...) — parsed from / emitted to the opcodes of kgb200.h.
"""
import struct

from . import ops

_NAME = {v: k for k, v in ops.OPC.items()}
_ARITY = {}
for _n in ("add", "sub", "mul", "div", "pow", "eq", "gt", "lt", "max", "min", "fmod", "eq_b", "gt_b", "lt_b", "ne_b", "ge_b",
           "le_b", "isclose", "floordiv", "mod"):
    _ARITY[_n] = 2
for _n in ("neg", "abs", "floor", "floor_i64", "trunc", "recip", "not_b", "to_f64", "to_i64", "to_f32", "to_b8", "to_i32",
           "sqrt", "exp", "log", "sin", "cos", "tan", "tanh", "ceil", "sign", "isnan_b", "isinf_b", "log2", "log10", "sinh",
           "cosh", "asin", "acos", "atan", "expm1", "log1p", "square"):
    _ARITY[_n] = 1
_ARITY["where"] = 3

ZERO = ("lit", 0.0)
ONE = ("lit", 1.0)


class NotDifferentiable(NotImplementedError):
    """The program contains an operator without a derivative rule: the caller falls back to the eager verbs."""


def parse(code):
    """postfix words -> tree"""
    st = []
    i = 0
    code = list(code)
    while i < len(code):
        w = code[i]
        if w == ops.OP_ARG:
            st.append(("arg", code[i + 1]))
            i += 2
        elif w == ops.OP_LIT_I64:
            st.append(("lit", struct.unpack("<q", struct.pack("<ii", code[i + 1], code[i + 2]))[0]))
            i += 3
        elif w == ops.OP_LIT_F64:
            st.append(("lit", struct.unpack("<d", struct.pack("<ii", code[i + 1], code[i + 2]))[0]))
            i += 3
        elif w == ops.OP_SCANVAL:
            raise NotDifferentiable("scan value inside an epilogue")
        else:
            name = _NAME.get(w)
            if name is None:
                raise ValueError(f"unknown opcode {w}")
            k = _ARITY[name]
            kids = st[-k:]
            del st[-k:]
            st.append(("op", name) + tuple(kids))
            i += 1
    if len(st) != 1:
        raise ValueError("malformed program")
    return st[0]


def emit(tree, out=None):
    """tree -> postfix words"""
    out = [] if out is None else out
    t = tree[0]
    if t == "arg":
        out += [ops.OP_ARG, tree[1]]
    elif t == "lit":
        out += ops.lit(tree[1])
    else:
        for k in tree[2:]:
            emit(k, out)
        out.append(ops.OPC[tree[1]])
    return out


def _is(t, v):
    return t[0] == "lit" and float(t[1]) == v


def op(name, *kids):
    """Smart constructor with the constant folds that keep tangent programs small."""
    if name == "add":
        a, b = kids
        if _is(a, 0.0):
            return b
        if _is(b, 0.0):
            return a
    elif name == "sub":
        a, b = kids
        if _is(b, 0.0):
            return a
        if _is(a, 0.0):
            return op("neg", b)
    elif name == "mul":
        a, b = kids
        if _is(a, 0.0) or _is(b, 0.0):
            return ZERO
        if _is(a, 1.0):
            return b
        if _is(b, 1.0):
            return a
    elif name == "div":
        a, b = kids
        if _is(a, 0.0):
            return ZERO
        if _is(b, 1.0):
            return a
    elif name == "neg":
        (a,) = kids
        if _is(a, 0.0):
            return ZERO
        if a[0] == "op" and a[1] == "neg":
            return a[2]
    elif name == "where":
        c, a, b = kids
        if a == b:
            return a
    return ("op", name) + tuple(kids)


def uses(tree, i):
    if tree[0] == "arg":
        return tree[1] == i
    if tree[0] == "lit":
        return False
    return any(uses(k, i) for k in tree[2:])


def tangent(tree, i):
    """d(tree)/d(arg i) as a tree (forward-mode on the expression; the trees are small)."""
    t = tree[0]
    if t == "arg":
        return ONE if tree[1] == i else ZERO
    if t == "lit":
        return ZERO
    name = tree[1]
    kids = tree[2:]
    if not uses(tree, i):
        return ZERO
    d = [tangent(k, i) for k in kids]
    if name == "add":
        return op("add", d[0], d[1])
    if name == "sub":
        return op("sub", d[0], d[1])
    if name == "mul":
        a, b = kids
        return op("add", op("mul", d[0], b), op("mul", a, d[1]))
    if name == "div":
        a, b = kids
        return op("sub", op("div", d[0], b), op("div", op("mul", a, d[1]), op("mul", b, b)))
    if name == "pow":
        a, b = kids
        # d(a^b) = b*a^(b-1)*da + a^b*log(a)*db
        ta = ZERO
        if not _is(d[0], 0.0):
            if b[0] == "lit":
                bm1 = ("lit", b[1] - 1)
                pw = a if _is(bm1, 1.0) else (ONE if _is(bm1, 0.0) else op("pow", a, bm1))
                ta = op("mul", d[0], op("mul", ("lit", float(b[1])), pw))
            else:
                ta = op("mul", d[0], op("mul", b, op("pow", a, op("sub", b, ONE))))
        tb = ZERO
        if not _is(d[1], 0.0):
            tb = op("mul", d[1], op("mul", tree, op("log", a)))
        return op("add", ta, tb)
    if name == "neg":
        return op("neg", d[0])
    if name == "abs":
        return op("mul", d[0], op("sign", kids[0]))
    if name == "sqrt":
        return op("div", d[0], op("mul", ("lit", 2.0), tree))
    if name == "exp":
        return op("mul", d[0], tree)
    if name == "expm1":
        return op("mul", d[0], op("exp", kids[0]))
    if name == "log":
        return op("div", d[0], kids[0])
    if name == "log2":
        return op("div", d[0], op("mul", kids[0], ("lit", 0.6931471805599453)))
    if name == "log10":
        return op("div", d[0], op("mul", kids[0], ("lit", 2.302585092994046)))
    if name == "log1p":
        return op("div", d[0], op("add", kids[0], ONE))
    if name == "sin":
        return op("mul", d[0], op("cos", kids[0]))
    if name == "cos":
        return op("neg", op("mul", d[0], op("sin", kids[0])))
    if name == "tan":
        return op("mul", d[0], op("add", ONE, op("mul", tree, tree)))
    if name == "tanh":
        return op("mul", d[0], op("sub", ONE, op("mul", tree, tree)))
    if name == "sinh":
        return op("mul", d[0], op("cosh", kids[0]))
    if name == "cosh":
        return op("mul", d[0], op("sinh", kids[0]))
    if name == "asin":
        return op("div", d[0], op("sqrt", op("sub", ONE, op("mul", kids[0], kids[0]))))
    if name == "acos":
        return op("neg", op("div", d[0], op("sqrt", op("sub", ONE, op("mul", kids[0], kids[0])))))
    if name == "atan":
        return op("div", d[0], op("add", ONE, op("mul", kids[0], kids[0])))
    if name == "recip":
        return op("neg", op("div", d[0], op("mul", kids[0], kids[0])))
    if name == "square":
        return op("mul", d[0], op("mul", ("lit", 2.0), kids[0]))
    if name in ("max", "min"):
        a, b = kids
        cmp = "gt_b" if name == "max" else "lt_b"
        # ties split evenly, like torch.maximum / torch.minimum
        tie = op("mul", ("lit", 0.5), op("add", d[0], d[1]))
        return op("where", op(cmp, a, b), d[0], op("where", op("eq_b", a, b), tie, d[1]))
    if name == "where":
        return op("where", kids[0], d[1], d[2])
    if name == "fmod":
        a, b = kids
        return op("sub", d[0], op("mul", op("trunc", op("div", a, b)), d[1]))
    if name == "mod":  # a - floor(a / b) * b
        a, b = kids
        return op("sub", d[0], op("mul", op("floor", op("div", a, b)), d[1]))
    if name in ("to_f64", "to_f32"):
        return d[0]
    if name in ("eq", "gt", "lt", "eq_b", "gt_b", "lt_b", "ne_b", "ge_b", "le_b", "isclose", "floor", "floor_i64", "trunc",
                "ceil", "sign", "not_b", "to_i64", "to_i32", "to_b8", "isnan_b", "isinf_b", "floordiv"):
        return ZERO  # piecewise constant
    raise NotDifferentiable(f"no derivative rule for {name}")


def grad_program(code, i, g_index):
    """Postfix program of  g * d f/d(arg i)  over (args..., g) where g is argument `g_index`; None if identically 0."""
    tree = parse(code)
    t = tangent(tree, i)
    if _is(t, 0.0):
        return None
    return emit(op("mul", ("arg", g_index), t))
