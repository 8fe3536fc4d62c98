"""Generate tests/golden/golden_1brc.json by running the REAL reference example on a small synthetic file.

    python tests/golden/make_golden_1brc.py

The reference's own ingest loop (`examples/1brc/par.py:_process_file_chunk`: `for line in f: location, measurement =
line.split(";")`, `np.asarray(temps, dtype=float)`) and its Klong worker (`worker.kg`: `stats`, grouping by station
NAME) are imported and executed unmodified on the NumPy backend; the text and the resulting per-station
[min max sum count] table are committed as the fixture (the reference tree is not available on the GPU box).
"""
import contextlib
import importlib.util
import io
import json
import os
import sys
import tempfile

import numpy as np

REF = os.environ.get("KLONGPY_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))

rng = np.random.default_rng(20261020)
names = ["Abha", "Zürich", "St. John's", "a", "Ürümqi", "Washington, D.C.", "Petropavlovsk-Kamchatsky", "İzmir", "Ngaoundéré",
         "Las Palmas de Gran Canaria", "x" * 100, "Xi'an", "Ab", "Abh", "Abhaa"]
lines = []
for _ in range(5000):
    s = names[rng.integers(len(names))]
    t = round(float(rng.normal(10, 25)), 1)
    t = max(-99.9, min(99.9, t))
    lines.append(f"{s};{t:.1f}")
lines += ["a;-0.0", "a;0.0", "Ab;-99.9", "Abh;99.9", "Abhaa;-0.1", "Xi'an;5.0"]
text = "\n".join(lines) + "\n"

spec = importlib.util.spec_from_file_location("ref_par", os.path.join(REF, "examples", "1brc", "par.py"))
par = importlib.util.module_from_spec(spec)
spec.loader.exec_module(par)
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "measurements.txt")
    with open(path, "w", encoding="utf-8", newline="") as f:
        f.write(text)
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "examples", "1brc"))   # _process_file_chunk does klong('.l("worker.kg")')
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            m = par._process_file_chunk(path, 0, os.path.getsize(path))
    finally:
        os.chdir(cwd)
table = {str(k): [float(v[0]), float(v[1]), float(v[2]), int(v[3])] for k, v in m.items()}
json.dump({"text": text, "table_min_max_sum_count": table}, open(os.path.join(HERE, "golden_1brc.json"), "w"),
          ensure_ascii=False, indent=0)
print(len(table), "stations;", len(lines), "rows")
