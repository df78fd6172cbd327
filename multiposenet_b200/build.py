"""Build recipe of libmpn_b200.so (hand-written sm_100a CUDA + the C ABI of include/mpn_b200.h).

    python -m multiposenet_b200.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU.  The library is built IN-TREE
(multiposenet_b200/libmpn_b200.so) so that it travels with the source snapshot.
Flags that matter for parity: -fmad=false (no implicit FMA contraction on the device; every fused
operation in the kernels is an explicit fmaf), no --use_fast_math, IEEE division and square root,
and -ffp-contract=off for the host code that prepares the anchor tables.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libmpn_b200.so")
SOURCES = ["api.cu", "detect.cu", "heatmap.cu", "kpdecode.cu", "prn_simt.cu", "prn_tcgen05.cu", "prn_fused.cu", "prn_split3.cu", "prn_big.cu", "math_test.cu"]
HEADERS = ["common.cuh", "handle.cuh", "mpn_math.cuh", "tcgen05_utils.cuh", "bulk_copy.cuh", os.path.join("..", "..", "include", "mpn_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden,-Wall",
    "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


# Variant builds for measurements (never loaded unless MPN_LIB points at them): name -> extra nvcc flags
VARIANTS = {"waves1": ["-DMPN_FC1_WAVES=1"], "waves2": ["-DMPN_FC1_WAVES=2"], "rankcut2t": ["-DMPN_NMS_RANK_CUT=4096"]}


def build_variant(name, force=False):
    """libmpn_b200_<name>.so next to the product library, compiled with VARIANTS[name]."""
    global OBJ, LIB
    keep = (OBJ, LIB, list(NVCC_FLAGS))
    try:
        OBJ = os.path.join(HERE, "_obj_" + name)
        LIB = os.path.join(HERE, f"libmpn_b200_{name}.so")
        NVCC_FLAGS.extend(VARIANTS[name])
        return build_library(force=force)
    finally:
        OBJ, LIB = keep[0], keep[1]
        NVCC_FLAGS[:] = keep[2]


def build_library(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    nvcc = _nvcc()
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers + [os.path.abspath(__file__)]):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [nvcc] + NVCC_FLAGS + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, o, cmd, r

    logs = []
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, o, cmd, r in ex.map(compile_one, jobs):
            if verbose or r.returncode != 0:
                print(" ".join(cmd), file=sys.stderr)
                print(r.stdout, r.stderr, file=sys.stderr)
            logs.append((s, r.stderr))
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {s}")
    objs = [os.path.join(OBJ, src.replace(".cu", ".o")) for src in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
                                                      "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            print(" ".join(cmd), file=sys.stderr)
            print(r.stdout, r.stderr, file=sys.stderr)
        if r.returncode != 0:
            raise RuntimeError("link failed")
    with open(os.path.join(OBJ, "ptxas.log"), "a" if not force else "w") as f:
        for s, log in logs:
            f.write(f"==== {s}\n{log}\n")
    return LIB


if __name__ == "__main__":
    for v in [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]:
        print(build_variant(v, force="--force" in sys.argv))
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv))
