"""Build libb200d.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import hashlib
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libb200d.so")
_STAMP = os.path.join(_HERE, "csrc", ".build_stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest():
    h = hashlib.sha256()
    for p in _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(_HERE, "..", "include", "b200d.h")]:
        with open(p, "rb") as f:
            h.update(os.path.basename(p).encode())  # (not the absolute path: the stamp must stay valid when the tree is copied)
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def library_is_current() -> bool:
    """The in-tree library exists and was built from the sources next to it (digest stamp)."""
    return os.path.exists(LIB_PATH) and os.path.exists(_STAMP) and open(_STAMP).read().strip() == _digest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(_STAMP) and open(_STAMP).read().strip() == digest:
        return LIB_PATH
    import fcntl

    with open(os.path.join(CSRC, ".build_lock"), "w") as lock:  # several ranks may arrive here at once
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and os.path.exists(LIB_PATH) and os.path.exists(_STAMP) and open(_STAMP).read().strip() == digest:
            return LIB_PATH
        return _build_locked(digest, verbose)


def _build_locked(digest: str, verbose: bool) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in _sources():
        obj = src[:-3] + ".o"
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(out)
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
    subprocess.check_call(cmd)
    with open(_STAMP, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
