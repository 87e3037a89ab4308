"""Build ``librtdf.so`` (hand-written sm_100a kernels + C-ABI) in-tree with nvcc.

The built library is git-ignored but travels with the repo snapshot to the GPU box.
``python real-time-deepfake-speech-detection_b200/build.py [--force] [--verbose]``
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "librtdf.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr",
         "-I", os.path.join(os.path.dirname(HERE), "include"), "-I", CSRC]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path):
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + [os.path.join("..", "..", "include", "rtdf.h")]:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p) and (f.endswith((".cuh", ".h")) or os.path.abspath(p) == os.path.abspath(path)):
            h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src, verbose):
    path = os.path.join(CSRC, src)
    obj = os.path.join(OUT_DIR, src[:-3] + ".o")
    stamp = obj + ".sha"
    dig = _digest(path)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False, ""
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    open(stamp, "w").write(dig)
    return obj, True, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OUT_DIR):
            os.remove(os.path.join(OUT_DIR, f))
    srcs = _sources()
    rebuilt = False
    objs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        for obj, did, log in ex.map(lambda s: _compile(s, verbose), srcs):
            objs.append(obj)
            rebuilt |= did
            if verbose and log:
                print(log)
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
