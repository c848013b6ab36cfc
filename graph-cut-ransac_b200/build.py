"""Builds libgcr_b200.so (all CUDA kernels + the extern "C" layer) for sm_100a with nvcc, in-tree.

    python graph-cut-ransac_b200/build.py [--force] [--verbose]

No torch / pybind dependency: the library is a plain C-ABI shared object (include/gcr/gcr.h).
-fmad=false: FP64 code keeps the reference's separately rounded mul/add; FP32 hot loops spell their FMAs out.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("GCR_LIB_OUT") or os.path.join(HERE, "libgcr_b200.so")   # GCR_LIB_OUT + GCR_NVCC_EXTRA: A/B variants
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC,-O3", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))  # host/ (adapter header, batched driver) builds separately


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "gcr", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    objdir = os.path.join(HERE, "build" + ("_" + os.path.basename(OUT).replace(".so", "") if os.environ.get("GCR_LIB_OUT") else ""))
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        objs.append(obj)
        cmd = [NVCC] + FLAGS + os.environ.get("GCR_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (src, out))
        elif verbose:
            sys.stdout.write(out)
    if failed:
        raise RuntimeError("CUDA build failed")
    cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-lcudart", "-ldl"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
