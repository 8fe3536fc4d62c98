"""Generate tests/golden/golden_table.json by running the REFERENCE's Table class (klongpy/db/sys_fn_db.py:18-127).

    PYTHONPATH=/root/reference python tests/golden/make_table_golden.py

The reference module imports duckdb at the top (only `Database.__call__` uses it, for SQL); duckdb is not installed
here, so an empty stand-in module is put in sys.modules first.  Everything recorded below comes from the reference's
own Table / pandas code path: `rows` is `Table.get_dataframe().values`, i.e. what `db("select * from T")` is built
from (sys_fn_db.py:143-144).  The scenarios are the reference's own kg tests (tests/kgtests/db/*.kg) plus seeded
random tables; the grouped statistics are pandas' groupby on the same frame (the SQL the db examples run:
`select k, min(v), avg(v), max(v), count(*) from T group by k order by k`).
"""
import json
import os
import sys
import types

import numpy as np

sys.modules.setdefault("duckdb", types.ModuleType("duckdb"))
sys.path.insert(0, "/root/reference")
from klongpy.db.sys_fn_db import Table  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def rows(t):
    return t.get_dataframe().values.tolist()


def run(script):
    """script: list of steps; returns the list of observations."""
    out = []
    t = None
    for step in script:
        op = step[0]
        if op == "table":
            cols = step[1]
            t = Table({k: np.asarray(v) for k, v in cols}, columns=[k for k, _ in cols])
        elif op == "insert":
            t.insert(np.asarray(step[1]))
        elif op == "insertb":
            t.insertb(np.asarray(step[1]))
        elif op == "index":
            t.set_index(list(step[1]))
        elif op == "rindex":
            t.reset_index()
        elif op == "set":
            t.set(step[1], np.asarray(step[2]))
        elif op == "rows":
            out.append({"rows": rows(t), "schema": list(t.schema()), "len": len(t)})
        elif op == "get":
            out.append({"col": t.get(step[1]).tolist()})
        else:
            raise ValueError(op)
    return out


def scenarios():
    abc = [["a", [1, 2, 3]], ["b", [2, 3, 4]], ["c", [3, 4, 5]]]
    s = {}
    s["create_multi_col"] = [["table", abc], ["rows"]]
    s["one_col"] = [["table", [["a", [1, 2, 3]]]], ["rows"]]
    s["insert_no_index"] = [["table", abc], ["insert", [4, 5, 6]], ["rows"], ["get", "a"]]
    s["bulk_insert"] = [["table", abc], ["insertb", [[4, 5, 6], [7, 8, 9]]], ["rows"]]
    s["single_index"] = [["table", abc], ["insert", [4, 5, 6]], ["rows"], ["insert", [4, 5, 6]], ["rows"],
                         ["index", ["a"]], ["rows"], ["insert", [4, 5, 6]], ["rows"], ["insert", [4, 6, 7]], ["rows"],
                         ["rindex"], ["insert", [4, 6, 7]], ["rows"]]
    s["multi_index"] = [["table", abc], ["index", ["a", "b"]], ["rows"], ["insert", [4, 5, 6]], ["rows"],
                        ["insert", [4, 5, 6]], ["rows"], ["insert", [4, 6, 7]], ["rows"], ["rindex"],
                        ["insert", [4, 6, 7]], ["rows"]]
    s["set_column"] = [["table", abc[:2]], ["set", "c", [3, 4, 5]], ["rows"], ["set", "a", [9, 8, 7]], ["rows"]]
    rng = np.random.default_rng(7)
    n = 300
    cols = [["k", rng.integers(0, 40, n).tolist()], ["j", rng.integers(0, 3, n).tolist()],
            ["v", rng.integers(-50, 50, n).tolist()]]
    up = np.stack([rng.permutation(80)[:50], rng.integers(0, 3, 50), rng.integers(100, 200, 50)], axis=1).tolist()
    s["random_single_index"] = [["table", cols], ["index", ["k"]], ["rows"], ["insertb", up], ["rows"], ["rindex"], ["rows"]]
    # buffered keys (k, j) are distinct so that the upsert is well defined (duplicate keys inside ONE commit are not
    # pinned by any reference test)
    kj = rng.permutation(120)[:60]
    up2 = np.stack([kj // 3, kj % 3, rng.integers(100, 200, 60)], axis=1).tolist()
    s["random_multi_index"] = [["table", cols], ["index", ["k", "j"]], ["rows"], ["insertb", up2], ["rows"]]
    return s


def group_stats_cases():
    import pandas as pd
    rng = np.random.default_rng(11)
    out = {}
    for name, n, g, lo in (("small", 500, 13, 0), ("sparse_negative", 2000, 50, -1_000_000_007)):
        k = (rng.integers(0, g, n) * (7919 if lo else 1) + lo).astype(np.int64)
        v = np.round(rng.normal(10, 15, n), 1)
        r = pd.DataFrame({"k": k, "v": v}).groupby("k")["v"].agg(["min", "mean", "max", "count"]).sort_index()
        out[name] = {"k": k.tolist(), "v": v.tolist(), "keys": r.index.tolist(), "min": r["min"].tolist(),
                     "mean": r["mean"].tolist(), "max": r["max"].tolist(), "count": r["count"].tolist()}
    return out


if __name__ == "__main__":
    sc = scenarios()
    doc = {"scenarios": {name: {"script": script, "expect": run(script)} for name, script in sc.items()},
           "group_stats": group_stats_cases()}
    with open(os.path.join(HERE, "golden_table.json"), "w") as f:
        json.dump(doc, f)
    print("wrote golden_table.json:", {k: len(v["expect"]) for k, v in doc["scenarios"].items()})
