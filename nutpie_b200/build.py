"""Build the CUDA core (libnutpie_b200.so) in-tree with nvcc for sm_100a.

Usage: python -m nutpie_b200.build [--force]
Every csrc/*.cu is compiled to an object in parallel (one nvcc process each,
`-gencode arch=compute_100a,code=sm_100a -lineinfo`), then linked into the shared
library.  The .so is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "build"
SO = PKG / "libnutpie_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
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
    cu, hdr = sources()
    nvcc = _nvcc()
    ccbin = ["-ccbin", "/usr/bin/g++"] if Path("/usr/bin/g++").exists() else []
    OBJ.mkdir(exist_ok=True)
    newest_hdr = max(p.stat().st_mtime for p in hdr)
    log = []

    def compile_one(src: Path):
        obj = OBJ / (src.stem + ".o")
        if not force and obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, newest_hdr):
            return obj, 0, f"[up to date] {src.name}\n"
        extra = os.environ.get("NB200_EXTRA_NVCC_FLAGS", "").split()
        cmd = [nvcc, *ccbin, *NVCC_FLAGS, *extra, "-c", "-o", str(obj), str(src)]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return obj, res.returncode, " ".join(cmd) + "\n" + res.stdout

    with ThreadPoolExecutor(max_workers=min(len(cu), os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, cu))
    objs = []
    failed = False
    for obj, rc, out in results:
        log.append(out)
        objs.append(str(obj))
        failed |= rc != 0
    if not failed:
        cmd = [nvcc, *ccbin, "-shared", "-o", str(SO), *objs]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        log.append(" ".join(cmd) + "\n" + res.stdout)
        failed = res.returncode != 0
    (PKG / "build.log").write_text("\n".join(log))
    if verbose or failed:
        print("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed, see nutpie_b200/build.log")
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
