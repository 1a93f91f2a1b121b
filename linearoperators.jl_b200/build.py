"""Build libb2o.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build() and by hand:
    python linearoperators.jl_b200/build.py [--force]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb2o.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["b2o_ctx.cu", "b2o_leaf.cu", "b2o_qn.cu", "b2o_graph.cu", "b2o_graph_aot.cu", "b2o_kron.cu", "b2o_dense.cu", "b2o_sparse.cu"]
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # reference-faithful rounding: Julia broadcasts never contract a*b+c
    "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
]


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build_lib(force=False, verbose=False):
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".inc"))]
    hdrs.append(os.path.join(HERE, "..", "include", "b2o.h"))
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(o)
        if force or not _newer(o, [s] + hdrs):
            extra = ["-Xptxas", "-v"] if verbose else []
            jobs.append([NVCC] + FLAGS + extra + ["-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        logs = list(ex.map(run, jobs))
    if verbose:
        for l in logs:
            sys.stderr.write(l)
    if jobs or not os.path.exists(OUT):
        run([NVCC, "-shared", "-o", OUT] + objs + ["-lcudart", "-ldl", "-ccbin", "/usr/bin/g++"])
    return OUT


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
