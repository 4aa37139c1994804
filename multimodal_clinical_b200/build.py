"""In-tree build of the sm_100a C-ABI library (no JIT cache: the .so travels with the repo snapshot).

Every csrc/*.cu is compiled to its own object (in parallel, only when it or a header changed) and the objects
are linked into _lf_fusion.so by nvcc; the objects live in csrc/_build/ (git- and gpurun-ignored)."""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
OUT = os.path.join(HERE, "_lf_fusion.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xptxas=-v", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))


MANIFEST = OUT + ".sources"


def _manifest() -> str:
    return "\n".join(os.path.basename(x) for x in sources()) + "\n" + " ".join(NVCC_FLAGS)


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    if not os.path.exists(MANIFEST) or open(MANIFEST).read() != _manifest():
        return True                       # a source was added / removed or the flags changed
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in sources() + _headers())


def _obj(src: str) -> str:
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def _compile(nvcc: str, src: str, extra, verbose: bool):
    res = subprocess.run([nvcc] + NVCC_FLAGS + extra + ["-c", "-o", _obj(src), src], capture_output=True, text=True)
    return src, res


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    extra = os.environ.get("LF_EXTRA_NVCC", "").split()
    os.makedirs(OBJ, exist_ok=True)
    flags_file = os.path.join(OBJ, "flags")
    flags = " ".join(NVCC_FLAGS + extra)
    if not os.path.exists(flags_file) or open(flags_file).read() != flags:
        force = True
    hdr_t = max([os.path.getmtime(h) for h in _headers()] or [0.0])
    todo = [s for s in sources()
            if force or not os.path.exists(_obj(s)) or os.path.getmtime(_obj(s)) < max(os.path.getmtime(s), hdr_t)]
    failed = False
    with cf.ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        for src, res in ex.map(lambda s: _compile(nvcc, s, extra, verbose), todo):
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            failed |= res.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building _lf_fusion.so")
    with open(flags_file, "w") as f:
        f.write(flags)
    objs = [_obj(s) for s in sources()]
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + ["-lcuda"],
                         capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed linking _lf_fusion.so")
    with open(MANIFEST, "w") as f:
        f.write(_manifest())
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
