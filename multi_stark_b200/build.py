"""Builds libmsgpu.so (the C-ABI CUDA library, sm_100a only) in-tree with nvcc.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot. There is no CPU
fallback: if the library is missing or cannot be loaded the package raises."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmsgpu.so")
OBJ = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unused-function", "--expt-relaxed-constexpr", "--extended-lambda"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(HERE, "host", f) for f in os.listdir(os.path.join(HERE, "host"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "msgpu.h"))
    deps.append(os.path.abspath(__file__))
    return deps


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(os.path.join(HERE, "libmshost.so")):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=False):
    """One builder at a time: under torchrun every rank imports the package at once; without the lock N ranks would run nvcc
    into the same objects and a rank could load a half-written .so. Outputs are written under a temporary name and renamed."""
    if not force and not needs_build():
        return LIB
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():  # another process built it while we waited
                return LIB
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose):
    os.makedirs(OBJ, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        cmd = [NVCC, *FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, _sources()))
    tmp = LIB + ".tmp.%d" % os.getpid()
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(tmp, LIB)
    build_host()
    return LIB


HOST_LIB = os.path.join(HERE, "libmshost.so")


def build_host():
    """libmshost.so: the host-side protocol layer (C++17, no CUDA) over the libmsgpu C ABI."""
    src = sorted(os.path.join(HERE, "host", f) for f in os.listdir(os.path.join(HERE, "host")) if f.endswith(".cpp"))
    tmp = HOST_LIB + ".tmp.%d" % os.getpid()
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-function", "-o", tmp, *src,
           "-L" + HERE, "-lmsgpu", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("host library build failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(tmp, HOST_LIB)
    return HOST_LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
