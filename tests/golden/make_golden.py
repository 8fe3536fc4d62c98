"""Generate tests/golden/*.json + *.npz by running the REAL reference (briangu/klongpy, NumPy backend).

Run in the build container (the reference tree is mounted read-only at /root/reference and is NOT available
on the GPU box, so its outputs are committed as fixtures):

    python tests/golden/make_golden.py

For every case we record the Klong source, the compiler IR the reference's `compiler._ast_to_ir` produced
(so the GPU parity tests can feed the identical tuple tree to `B200BackendProvider.compile_expr_ir`),
the seeded inputs and the reference's result.
"""
import json
import os
import sys

import numpy as np

REF = os.environ.get("KLONGPY_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

from klongpy import KlongInterpreter  # noqa: E402
from klongpy.compiler import _ast_to_ir  # noqa: E402
from klongpy.core import KGSym  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(20261019)


def jsonable(x):
    if isinstance(x, tuple):
        return [jsonable(v) for v in x]
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.floating,)):
        return float(x)
    return x


def get_ir(klong, src):
    """Parse one expression with the reference parser and lower it with the reference compiler."""
    i, ast = klong.prog(src)
    assert len(ast) == 1, src
    var_refs = {}
    ir = _ast_to_ir(ast[0], klong, var_refs)
    return ir, [str(k) for k in var_refs.keys()]


# ------------------------------------------------------------------------------------------------ compiled expressions
N1 = 1000
inputs = {
    "a": rng.random(N1),
    "b": rng.random(N1),
    "c": rng.random(N1),
    "x": rng.integers(0, 256, N1) / 256.0,           # dyadic: every partial sum exact
    "ts": 100 + rng.standard_normal(N1) * 0.5,
    "p": 50 + rng.standard_normal(N1) * 10,
    "s": (np.floor(rng.random(N1) * 9900) + 100).astype(np.int64),
    "i": rng.integers(-50, 50, N1),
    "j": rng.integers(1, 9, N1),
    "e": np.zeros(0),
    "one": np.array([5.0]),
    "k": 3,
    "f": 2.5,
}

# every expression of examples/bench_compiler.kg plus IR corner cases (Appendix C of SURVEY.md)
EXPRS = [
    "a+b", "a*2+b", "(a+b)*(a-b)", "(a*2+b)%(a+1)", "(a*2+b*3)%(a+1)-(b*c)", "a*b+c", "(a+b)*(b+c)*(a-c)",
    "+/a", "+/a*b", "|/a-b", "&/a", "*/j", "+/a*b+c", "+/(a-b)^2", "+/a=b", "+/i=i", "+/a>b", "+/a<0.5",
    "+\\ts", "+\\x", "*\\j", "|\\ts", "&\\ts", "(|\\ts)-ts", "|/(|\\ts)-ts", "+\\i",
    "(+/p*s)%+/s", "(+/a)%+/b", "-a", "-i", "a-(+/a)%1000", "a^2", "i^2", "a^0.5", "a^b", "i*k", "a*f", "i%j",
    "i=j", "a<b", "i>0", "+/((x*2)+1)^2", "+/x", "+\\(x*2)+1", "(a*2)+1", "i+k", "i-f",
    "+/e", "*/e", "+\\e", "+/one", "(+\\a)%+/a", "+/+\\x", "|/i", "&/i", "+/i", "*/a",
]


def run_compiled(klong):
    cases = []
    arrays = {}
    for name, v in inputs.items():
        klong[name] = v
        if isinstance(v, np.ndarray):
            arrays["in_" + name] = v
    for n, src in enumerate(EXPRS):
        ir, params = get_ir(klong, src)
        assert ir is not None, f"reference compiler rejected {src}"
        res = klong._backend.compile_expr_ir(ir, params)
        env = {f"_v{k}": klong[KGSym(p)] for k, p in enumerate(params)}
        if res is not None:
            fn, _ = res
            try:
                out = fn(*[klong[KGSym(p)] for p in params])
                how = "numpy-compiled"
            except Exception as ex:  # e.g. np.maximum.reduce([]) -> interpreter path
                out = klong(src)
                how = f"interpreter ({type(ex).__name__})"
        else:
            out = klong(src)  # |\ and &\ are not compilable on NumPy: interpreter result
            how = "interpreter"
        key = f"out_{n}"
        arrays[key] = np.asarray(out)
        cases.append({"src": src, "ir": jsonable(ir), "params": params, "out": key, "how": how,
                      "out_dtype": str(np.asarray(out).dtype), "out_shape": list(np.asarray(out).shape)})
    return cases, arrays


# ------------------------------------------------------------------------------------------------ interpreter verbs
def run_verbs(klong):
    cases = []
    arrays = {}
    data = {
        "g1": rng.permutation(500),
        "g2": rng.integers(0, 20, 500),
        "g3": rng.standard_normal(300),
        "g4": np.array([3, 1, 2, 1]),
        "g5": rng.integers(-2**62, 2**62, 400),
        "g6": np.round(rng.standard_normal(300), 1),
        "w1": rng.integers(0, 4, 200),
        "ix": rng.integers(0, 500, 100),
        "f1": rng.random(500),
    }
    for name, v in data.items():
        klong[name] = v
        arrays["in_" + name] = v
    # grade on tie-heavy keys is nondeterministic in the reference (unstable argsort): record only
    # tie-free grades from the reference; tie cases are pinned by the reference's own golden vectors below.
    progs = [
        ("<g1", "grade"), (">g1", "grade"), ("<g3", "grade"), (">g3", "grade"), ("<g5", "grade"), (">g5", "grade"),
        ("=g2", "group"), ("=g4", "group"), ("=g6", "group"), ("=g1", "group"),
        ("?g2", "range"), ("?g4", "range"), ("?g6", "range"),
        ("g2?7", "find"), ("g4?1", "find"), ("g2?99", "find"), ("g6?0.1", "find"),
        ("!10", "enumerate"), ("!0", "enumerate"), ("&w1", "expand"), ("&g4", "expand"), ("&5", "expand"),
        ("g1@ix", "at"), ("f1@ix", "at"), ("f1@<f1", "at"), ("|g1", "reverse"), ("|f1", "reverse"),
        ("g2+g1", "dyad"), ("g2*2.5", "dyad"), ("g2%g4@0", "dyad"), ("g2|10", "dyad"), ("g2&10", "dyad"),
        ("g1!7", "dyad"), ("g2<10", "dyad"), ("g2>10", "dyad"), ("g2=10", "dyad"), ("g2^2", "dyad"),
        ("f1^2", "dyad"), ("f1^0.5", "dyad"), ("g2:%3", "dyad"),
        ("-g1", "monad"), ("_g3", "monad"), ("%f1", "monad"), ("#g3", "monad"), ("~w1", "monad"), ("#g1", "monad"),
        ("+/g1", "over"), ("-/g2", "over"), ("*/g4", "over"), ("%/g4", "over"), ("|/g3", "over"), ("&/g3", "over"),
        ("+\\g2", "scan"), ("-\\g2", "scan"), ("*\\g4", "scan"), ("%\\g4", "scan"), ("|\\g3", "scan"), ("&\\g3", "scan"),
        ("5#g1", "take"), ("(-5)#g1", "take"), ("5_g1", "drop"), ("(-5)_g1", "drop"), ("g4,g4", "join"), ("g4,7", "join"),
        ("2:+g4", "rotate"), ("^g1", "shape"), ("510#g1", "take"),
    ]
    for n, (src, kind) in enumerate(progs):
        out = klong(src)
        key = f"v_{n}"
        if isinstance(out, np.ndarray) and out.dtype == object and all(np.ndim(x) == 0 for x in out):
            # object array of plain numbers (e.g. ~a goes through logical_not on an object array)
            arrays[key] = np.asarray(out.tolist())
            cases.append({"src": src, "kind": kind, "out": key, "ragged": False, "object_scalars": True,
                          "out_dtype": str(arrays[key].dtype)})
        elif isinstance(out, np.ndarray) and out.dtype == object:
            # ragged group result: store as CSR
            flat = np.concatenate([np.asarray(x) for x in out]) if len(out) else np.zeros(0, dtype=np.int64)
            starts = np.concatenate([[0], np.cumsum([len(x) for x in out])]).astype(np.int64)
            arrays[key + "_flat"] = flat
            arrays[key + "_starts"] = starts
            cases.append({"src": src, "kind": kind, "out": key, "ragged": True})
        else:
            arrays[key] = np.asarray(out)
            cases.append({"src": src, "kind": kind, "out": key, "ragged": False,
                          "out_dtype": str(np.asarray(out).dtype)})
    return cases, arrays


# golden vectors copied (as data) from the reference's own suite, tests/test_suite.py (file:line in `ref`)
REFERENCE_SUITE_VECTORS = [
    {"ref": "tests/test_suite.py:74", "src": ">[1 2 3 4 5]", "expect": [4, 3, 2, 1, 0]},
    {"ref": "tests/test_suite.py:90", "src": "<[1 2 3 4 5]", "expect": [0, 1, 2, 3, 4]},
    {"ref": "tests/test_suite.py:104", "src": "=[1 2 3 4]", "expect": [[0], [1], [2], [3]]},
    {"ref": "tests/test_suite.py:185", "src": "?[1 2 3 4]", "expect": [1, 2, 3, 4]},
    {"ref": "tests/test_suite.py:186", "src": "?[1 1 1 2 2]", "expect": [1, 2]},
    {"ref": "tests/test_suite.py:452", "src": "[1 2 3 1 2 1]?1", "expect": [0, 3, 5]},
    {"ref": "tests/test_suite.py:453", "src": "[1 2 3]?4", "expect": []},
    {"ref": "SURVEY.md §8c [measured on the reference]", "src": ">[3 1 2 1]", "expect": [0, 2, 3, 1]},
    {"ref": "SURVEY.md §8c [measured on the reference]", "src": "<[3 1 2 1]", "expect": [1, 3, 2, 0]},
]


def verify_suite_vectors(klong):
    """Re-evaluate the lifted vectors on the reference so a typo here cannot silently pin a wrong answer."""
    out = []
    for v in REFERENCE_SUITE_VECTORS:
        r = klong(v["src"])
        r = [list(map(int, x)) for x in r] if (isinstance(r, np.ndarray) and r.ndim == 2) or \
            (isinstance(r, np.ndarray) and r.dtype == object) else [int(x) for x in np.asarray(r).tolist()]
        stable_ok = r == v["expect"]
        out.append({**v, "reference_result": r, "reference_agrees": stable_ok})
    return out


def main():
    import klongpy
    klong = KlongInterpreter(backend="numpy")
    ccases, carr = run_compiled(klong)
    klong2 = KlongInterpreter(backend="numpy")
    vcases, varr = run_verbs(klong2)
    suite = verify_suite_vectors(KlongInterpreter(backend="numpy"))
    np.savez_compressed(os.path.join(HERE, "golden_compiled.npz"), **carr)
    np.savez_compressed(os.path.join(HERE, "golden_verbs.npz"), **varr)
    meta = {"reference": "briangu/klongpy", "reference_version": getattr(klongpy, "__version__", "0.7.0"),
            "numpy": np.__version__, "seed": 20261019, "compiled": ccases, "verbs": vcases, "suite_vectors": suite}
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(f"{len(ccases)} compiled cases, {len(vcases)} verb cases, {len(suite)} suite vectors")
    bad = [s for s in suite if not s["reference_agrees"]]
    if bad:
        print("suite vectors where this container's reference disagrees (unstable argsort ties):", bad)


if __name__ == "__main__":
    main()
