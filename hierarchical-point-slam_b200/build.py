"""Build libhps_b200.so in-tree with nvcc for sm_100a (no torch headers needed: the boundary is a plain
C ABI, see include/hps_b200.h).  `python hierarchical-point-slam_b200/build.py [--force] [--verbose]`."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhps_b200.so")
SOURCES = ["capi.cu", "grid.cu", "interp.cu", "composite.cu", "mlp.cu", "sample.cu", "loss.cu", "mlp_tc.cu", "mlp_tc_bwd.cu", "tc_selftest.cu", "optim.cu", "peer.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]


def nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "hps_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    objs = []
    odir = os.path.join(HERE, "build")
    os.makedirs(odir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(odir, src.replace(".cu", ".o"))
        cmd = [nvcc()] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode:
            sys.stderr.write(out)
        if pr.returncode:
            raise RuntimeError("nvcc failed on %s" % src)
    # same -gencode at link time: without it nvcc adds a default-arch (sm_52) stub and warns about it
    subprocess.check_call([nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", LIB] + objs + ["-lcudart"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
