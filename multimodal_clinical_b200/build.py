"""In-tree build of the sm_100a C-ABI library (no JIT cache: the .so travels with the repo snapshot)."""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_lf_fusion.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xptxas=-v", "-shared", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


MANIFEST = OUT + ".sources"


def _manifest() -> str:
    return "\n".join(os.path.basename(x) for x in sources()) + "\n" + " ".join(NVCC_FLAGS)


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    if not os.path.exists(MANIFEST) or open(MANIFEST).read() != _manifest():
        return True                       # a source was added / removed or the flags changed
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("LF_EXTRA_NVCC", "").split() + ["-o", OUT] + sources() + ["-lcuda"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building _lf_fusion.so")
    with open(MANIFEST, "w") as f:
        f.write(_manifest())
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
