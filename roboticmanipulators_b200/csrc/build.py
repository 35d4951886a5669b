#!/usr/bin/env python
"""Build librk_b200.so in-tree with nvcc for sm_100a (no torch, no cmake).

The library travels to the GPU box as a built artefact (git-ignored, not
gpurun-ignored).  Usage: python roboticmanipulators_b200/csrc/build.py [--force] [-v]
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(os.path.dirname(HERE), "librk_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def sources():
    return sorted(glob.glob(os.path.join(HERE, "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(HERE, "*.cuh")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = nvcc_path()
    common = [nvcc, *ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
              "-I", os.path.join(ROOT, "include"), "-I", HERE]
    if verbose:
        common += ["-Xptxas", "-v"]
    procs = []
    objs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        procs.append((src, subprocess.Popen(common + ["-c", src, "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static", "-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
