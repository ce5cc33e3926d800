"""Builds the sm_100a CUDA library in-tree (libaz_b200.so next to this file).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the gpurun snapshot.
Objects and the library are keyed on a CONTENT hash of their sources, headers and flags (<target>.buildhash), so stale
objects beside a fresh checkout are never trusted.
"""
import hashlib
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
    ("nn_heads_tc.cu", []),
    ("replay.cu", ["-fmad=false"]),
    ("dbg.cu", []),
]


def _digest(paths, extra=()):
    """sha256 over file CONTENTS (not mtimes) plus the compiler command: a fresh checkout next to stale objects rebuilds."""
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    for x in extra:
        h.update(str(x).encode())
    return h.hexdigest()


def _up_to_date(target, digest):
    stamp = target + ".buildhash"
    if not os.path.exists(target) or not os.path.exists(stamp):
        return False
    with open(stamp) as f:
        return f.read().strip() == digest


def _stamp(target, digest):
    with open(target + ".buildhash", "w") as f:
        f.write(digest + "\n")


def build(verbose=False, force=False):
    # AZ_NVCC_DEFINES="-DAZ_ADV_TIMING" builds the instrumented search kernel (tools/adv_timing.sh); part of the content hash
    extra_defs = os.environ.get("AZ_NVCC_DEFINES", "").split()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    inc = os.path.join(HERE, "..", "include")
    headers += [os.path.join(inc, f) for f in os.listdir(inc) if f.endswith((".h", ".hpp"))]
    headers.append(os.path.abspath(__file__))
    objs, digests = [], []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.rsplit(".", 1)[0] + ".o")
        objs.append(o)
        cmd = ["nvcc"] + ARCH + COMMON + extra + extra_defs + ["-c", s, "-o", o]
        d = _digest([s] + headers, ARCH + COMMON[:4] + extra + extra_defs)
        digests.append(d)
        if force or not _up_to_date(o, d):
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
            _stamp(o, d)
    link_digest = hashlib.sha256("".join(digests).encode()).hexdigest()
    if force or not _up_to_date(OUT, link_digest):
        cmd = ["nvcc"] + ARCH + ["-shared", "-o", OUT] + objs + ["-cudart", "static"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        _stamp(OUT, link_digest)
    return OUT


if __name__ == "__main__":
    build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(OUT)
