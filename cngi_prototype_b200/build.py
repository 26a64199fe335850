"""Builds cngi_prototype_b200/csrc/libcngi_b200.so with nvcc for sm_100a (in-tree, so it ships to the GPU box).

    python -m cngi_prototype_b200.build [--force] [--verbose]
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libcngi_b200.so")
INCLUDE = os.path.abspath(os.path.join(HERE, "..", "include"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall",
    "-Xptxas", "-v",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found (needed to build libcngi_b200.so)")
    return exe


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(f) <= t for f in _deps())


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library.  Returns the library path."""
    if not force and up_to_date():
        return LIB
    nvcc = _nvcc()
    objs = []
    logs = []
    procs = []
    for src in sources():
        obj = os.path.splitext(src)[0] + ".o"
        cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-c", src, "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, obj, pr in procs:
        out, _ = pr.communicate()
        logs.append("==== %s\n%s" % (os.path.basename(src), out))
        if pr.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % src)
        objs.append(obj)
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc)), "lib64")
    link = [nvcc, "-shared", "-o", LIB] + objs + ["-L", cuda_lib, "-lcufft", "-lcudart", "-ldl", "-lpthread",
                                                  "-Xlinker", "-rpath," + cuda_lib]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
