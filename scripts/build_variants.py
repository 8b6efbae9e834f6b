"""Build-time variants of the streaming leapfrog (fused U-turn partners F, pipeline stages S) as
separate libraries nutpie_b200/variants/libnutpie_b200_F{F}S{S}.so — only the two translation
units that depend on the macros are recompiled; select one with NB200_LIB=<path>.
Usage: python scripts/build_variants.py 1,3 2,3 3,2"""
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from nutpie_b200 import build as B

B.build()
out = B.PKG / "variants"
out.mkdir(exist_ok=True)
nvcc = B._nvcc()
base_objs = [str(B.OBJ / (p.stem + ".o")) for p in B.sources()[0] if p.stem not in ("kernels_normal", "nb200_api")]


def one(spec):
    f, s = spec.split(",")
    tag = f"F{f}S{s}"
    objs = []
    for stem in ("kernels_normal", "nb200_api"):
        o = out / f"{stem}_{tag}.o"
        cmd = [nvcc, "-ccbin", "/usr/bin/g++", *B.NVCC_FLAGS, f"-DNB200_MAX_FUSED={f}", f"-DNB200_STAGES={s}",
               "-c", "-o", str(o), str(B.CSRC / f"{stem}.cu")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            print(r.stdout, r.stderr)
            raise SystemExit(1)
        if stem == "kernels_normal":
            lines = (r.stdout + r.stderr).splitlines()
            for i, ln in enumerate(lines):
                if ("nuts_kernelINS_11NormalModelELi4ELi0" in ln or "nuts_kernelINS_11NormalModelELi8ELi0" in ln) \
                        and "Compiling" in ln:
                    print(tag, "W=4" if "ELi4E" in ln else "W=8", "|", lines[i + 2].strip(), "|", lines[i + 3].strip())
        objs.append(str(o))
    so = out / f"libnutpie_b200_{tag}.so"
    subprocess.run([nvcc, "-ccbin", "/usr/bin/g++", "-shared", "-o", str(so), *objs, *base_objs, "-ldl"], check=True)
    return so


with ThreadPoolExecutor(4) as ex:
    for so in ex.map(one, sys.argv[1:]):
        print("built", so)
