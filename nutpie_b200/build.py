"""Build the CUDA core (libnutpie_b200.so) in-tree with nvcc for sm_100a.

Usage: python -m nutpie_b200.build [--force]
The .so is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
SO = PKG / "libnutpie_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the B200 engine cannot be built")


def sources():
    return sorted(CSRC.glob("*.cu")), sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.hpp")) +
                                             [PKG.parent / "include" / "nutpie_b200.h"])


def needs_build() -> bool:
    if not SO.exists():
        return True
    cu, hdr = sources()
    t = SO.stat().st_mtime
    return any(p.stat().st_mtime > t for p in cu + hdr)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return SO
    cu, _ = sources()
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(SO), *map(str, cu)]
    env = dict(os.environ)
    # nvcc's host compiler: the system g++ (the image's $CXX lacks some runtime specs)
    if Path("/usr/bin/g++").exists():
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    (PKG / "build.log").write_text(" ".join(cmd) + "\n" + res.stdout)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed, see nutpie_b200/build.log")
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
