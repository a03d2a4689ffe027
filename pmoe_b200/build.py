"""In-tree build of the C-ABI library (nvcc, sm_100a only). The .so lands next to this file so that
it travels with the repo snapshot to the GPU box."""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libpmoe_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]
# --use_fast_math only for the bandwidth/tensor kernels; the loss / gating kernels keep IEEE log/exp/div.
FAST_MATH = {"conv_tc.cu", "conv_simt.cu", "eltwise.cu", "eltwise_bwd.cu"}


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path):
    h = hashlib.sha1()
    for f in sorted(os.listdir(CSRC)) + ["../../include/pmoe_b200.h"]:
        fp = os.path.join(CSRC, f)
        if os.path.isfile(fp) and (f.endswith((".cu", ".cuh", ".h"))):
            h.update(open(fp, "rb").read())
    h.update(" ".join(FLAGS).encode())
    h.update(path.encode())
    return h.hexdigest()


def _compile(src):
    os.makedirs(OBJ, exist_ok=True)
    obj = os.path.join(OBJ, src[:-3] + ".o")
    stamp = obj + ".sha"
    dig = _digest(src)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, ""
    cmd = [NVCC] + FLAGS + (["--use_fast_math"] if src in FAST_MATH else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    open(stamp, "w").write(dig)
    return obj, r.stderr


def build(verbose=False):
    srcs = _sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(_compile, srcs))
    objs = [o for o, _ in results]
    log = "\n".join(l for _, l in results if l)
    newest = max(os.path.getmtime(o) for o in objs)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
