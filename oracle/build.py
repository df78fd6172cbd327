"""Build recipe for the CPU oracle (TEST INFRASTRUCTURE, not product code).

`python -m oracle.build` compiles oracle/mpn_oracle.c into oracle/libmpn_oracle.so.
There is no `oracle/_ref`: the reference is pure Python on top of TensorFlow 1.15
and has no C/C++ sources to compile ("Oracle").
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "mpn_oracle.c")
HDR = os.path.join(HERE, "exact_math.h")


def _cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def lib_path(generic=False):
    return os.path.join(HERE, "libmpn_oracle_generic.so" if generic else "libmpn_oracle.so")


def build_oracle(force=False, generic=False, verbose=False):
    """Compile the oracle.  `generic=True` drops -mavx2 -mfma (for a host CPU without them)."""
    out = lib_path(generic)
    if not force and os.path.exists(out):
        newest = max(os.path.getmtime(SRC), os.path.getmtime(HDR))
        if os.path.getmtime(out) >= newest:
            return out
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
           "-fvisibility=hidden", "-Wall", "-Wextra", "-o", out, SRC, "-lm"]
    if not generic:
        cmd[1:1] = ["-mavx2", "-mfma"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return out


def ensure_oracle():
    """Path of a loadable oracle library for THIS host (builds it if missing)."""
    flags = _cpu_flags()
    generic = bool(flags) and not ({"avx2", "fma"} <= flags)
    return build_oracle(generic=generic)


if __name__ == "__main__":
    print(build_oracle(force=True, verbose=True))
