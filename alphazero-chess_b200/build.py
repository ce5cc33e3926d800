"""Builds the sm_100a CUDA library in-tree (libaz_b200.so next to this file).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libaz_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(HERE, "..", "include"), "-I", CSRC]

# translation units and their extra flags; the search/rules kernels must not contract a*b+c into FMA because visit
# counts are bit-exact against the reference's f32 arithmetic (tree.rs:187-189)
UNITS = [
    ("engine.cu", []),
    ("chess_kernels.cu", []),
    ("mcts.cu", ["-fmad=false"]),
    ("nn.cu", []),
    ("nn_tc.cu", []),
    ("nn_heads.cu", []),
    ("replay.cu", ["-fmad=false"]),
    ("dbg.cu", []),
]


def _stale(obj, srcs):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(s) > t for s in srcs)


def build(verbose=False, force=False):
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers += [os.path.join(HERE, "..", "include", f) for f in os.listdir(os.path.join(HERE, "..", "include"))]
    headers.append(os.path.abspath(__file__))
    objs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.rsplit(".", 1)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = ["nvcc"] + ARCH + COMMON + extra + ["-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
    if force or _stale(OUT, objs):
        cmd = ["nvcc"] + ARCH + ["-shared", "-o", OUT] + objs + ["-cudart", "static"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(OUT)
