"""Build libkgb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m klongpy_b200.csrc.build [--force] [--verbose]

The JIT skeleton (jit_skeleton.cuh) is embedded into the library as a string so that the shared object is
self-contained on the GPU box.  Output: klongpy_b200/libkgb200.so (git-ignored, travels with gpurun).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
BUILD = os.path.join(HERE, "build")
OUT = os.path.join(PKG, "libkgb200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
          "-I", os.path.join(ROOT, "include"), "-I", HERE] + os.environ.get("KGB200_NVCC_DEFS", "").split()  # tuning: extra -D flags

SOURCES = ["kgb_core.cpp", "kgb_expr.cpp", "kgb_sort.cu", "kgb_select.cu", "kgb_misc.cu", "kgb_groupagg.cu", "kgb_ingest.cu"]


def _embed_skeleton():
    src = open(os.path.join(HERE, "jit_skeleton.cuh")).read()
    assert ')KGBSKEL"' not in src
    # split into chunks: string literals have implementation limits
    chunks = [src[i:i + 8000] for i in range(0, len(src), 8000)]
    body = "\n".join('R"KGBSKEL(' + c + ')KGBSKEL"' for c in chunks)
    out = os.path.join(BUILD, "jit_skeleton_embed.cpp")
    text = "// generated from jit_skeleton.cuh by build.py\nextern const char* kgb_jit_skeleton_src;\n" \
           "const char* kgb_jit_skeleton_src =\n" + body + ";\n"
    if not os.path.exists(out) or open(out).read() != text:
        with open(out, "w") as f:
            f.write(text)
    return out


def _stamp(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(ARCH + COMMON).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    embed = _embed_skeleton()
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))] + [embed]
    deps = srcs + [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".h", ".cuh"))] + \
        [os.path.join(ROOT, "include", "kgb200.h")]
    stamp_file = os.path.join(BUILD, "stamp")
    stamp = _stamp(deps)
    if not force and os.path.exists(OUT) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return OUT
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(BUILD, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = [NVCC] + ARCH + COMMON + ["-x", "cu", "-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {s}\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("libkgb200 build failed")
    link = [NVCC] + ARCH + ["-shared", "-o", OUT] + objs + ["-cudart", "static", "-ldl", "-lpthread"]
    subprocess.check_call(link)
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
