"""Build libmhppo_b200.so (sm_100a only) in-tree with nvcc.

`python mh-ppo_b200/build.py` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU.
Objects are compiled in parallel (one translation unit per env class), then linked into
mh-ppo_b200/libmhppo_b200.so, which travels to the GPU box with the repo snapshot.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmhppo_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# env kernels restate double arithmetic in the reference's order: no FMA contraction (DESIGN.md "Precision")
FLAGS = {"env": ["-fmad=false"], "default": []}


def _flags_for(src):
    base = os.path.basename(src)
    # ppo_*.cu: the feature builders restate numpy fp32 mul/add (no contraction); the MLP uses explicit fmaf
    return FLAGS["env"] if base.startswith(("env_", "ppo_")) or base == "mhppo_api.cu" else FLAGS["default"]


def _deps():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "mhppo.h")]


def _compile(src, force):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    newest = max(os.path.getmtime(p) for p in [src] + _deps())
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj, ""
    cmd = [NVCC] + ARCH + COMMON + _flags_for(src) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(force=False, verbose=False, jobs=None):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    jobs = jobs or min(len(srcs), os.cpu_count() or 4)
    with cf.ThreadPoolExecutor(jobs) as ex:
        res = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [o for o, _ in res]
    log = "".join(l for _, l in res)
    if log:
        with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
            f.write(log)
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
