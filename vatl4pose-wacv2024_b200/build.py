"""Build libvatlq.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).
Every .cu is compiled to its own object (in parallel, cached by content hash), then linked."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libvatlq.so")
STAMP = LIB + ".stamp"
SOURCES = ["api.cu", "heatmap_scan.cu", "wpu.cu", "fuse.cu", "coreset.cu", "next_rows.cu", "extras.cu", "tc_dist.cu", "kmeans.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def _headers() -> bytes:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    files.append(os.path.join(os.path.dirname(HERE), "include", "vatlq.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode() + b"\0" + fh.read())   # (names, not absolute paths)
    h.update(" ".join(FLAGS).encode())
    return h.digest()


def _src_hash(src: str, hdr: bytes) -> str:
    with open(os.path.join(CSRC, src), "rb") as fh:
        return hashlib.sha256(hdr + src.encode() + b"\0" + fh.read()).hexdigest()


def find_nvcc() -> str | None:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    return None


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile what changed (or everything with force).  Returns the path of the shared library."""
    hdr = _headers()
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hashes = {s: _src_hash(s, hdr) for s in sources}
    fp = hashlib.sha256("".join(hashes[s] for s in sources).encode()).hexdigest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == fp:
                return LIB
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libvatlq.so")
    os.makedirs(OBJ, exist_ok=True)

    def compile_one(s: str):
        obj = os.path.join(OBJ, s[:-3] + "." + hashes[s][:16] + ".o")
        if force or not os.path.exists(obj):
            for old in os.listdir(OBJ):
                if old.startswith(s[:-3] + ".") and old.endswith(".o"):
                    os.remove(os.path.join(OBJ, old))
            cmd = [nvcc, *FLAGS, "-c", "-o", obj, os.path.join(CSRC, s)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {s}:\n" + res.stdout + res.stderr)
            if verbose:
                print(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    res = subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-ldl"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    with open(STAMP, "w") as fh:
        fh.write(fp)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
