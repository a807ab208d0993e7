"""Build libzkb200.so (sm_100a only) in-tree with nvcc.  No JIT cache: the .so sits next to this
file so it travels to the GPU box with the repository snapshot."""

from __future__ import annotations

import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libzkb200.so")
SOURCES = ["msm_g2_acc.cu", "msm_g2_red.cu", "msm_g2_tab.cu", "msm_g1.cu", "msm_g2.cu", "crs.cu", "prove.cu", "api.cu", "ntt.cu", "pairing.cu", "shard.cu", "wire.cu", "witness.cu"]
HEADERS = ["ff.cuh", "ec.cuh", "constants.h", "common.cuh", "msm_impl.cuh", "affine_level.cuh", os.path.join("..", "..", "include", "zkb200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _digest(paths: list, extra: str = "") -> str:
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _fresh(target: str, digest: str) -> bool:
    """Content-hash stamps instead of mtimes: the snapshot that travels to the GPU box does not
    preserve modification-time order, and a needless rebuild there costs minutes of GPU time."""
    stamp = target + ".sha256"
    return os.path.exists(target) and os.path.exists(stamp) and open(stamp).read().strip() == digest


def _stamp(target: str, digest: str) -> None:
    with open(target + ".sha256", "w") as f:
        f.write(digest)


EXTRA_HEADERS = {"pairing.cu": ["pairing.cuh"]}  # headers only one translation unit includes


def _compile(src: str) -> str:
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in HEADERS + EXTRA_HEADERS.get(src, [])]
    digest = _digest(deps, " ".join(NVCC_FLAGS))
    if not _fresh(obj, digest):
        cmd = [_nvcc()] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if _digest(deps, " ".join(NVCC_FLAGS)) == digest:  # sources unchanged while nvcc ran
            _stamp(obj, digest)
    return obj


def build(force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(_compile, SOURCES))
    digest = _digest(objs)
    if not _fresh(LIB, digest):
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        subprocess.check_call(cmd)
        _stamp(LIB, digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
