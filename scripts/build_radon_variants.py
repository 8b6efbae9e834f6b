"""Code-size experiments on the radon kernel: rebuild only kernels_radon.cu with extra -D flags and
link it with the default objects -> nutpie_b200/variants/libnutpie_b200_<tag>.so (NB200_LIB selects).
Usage: python scripts/build_radon_variants.py tag=-DFLAG1,-DFLAG2 ..."""
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
KERNEL = "nuts_kernelINS_10RadonModelELi1ELi6"


def one(spec):
    tag, flags = spec.split("=")
    o = out / f"kernels_radon_{tag}.o"
    r = subprocess.run([nvcc, "-ccbin", "/usr/bin/g++", *B.NVCC_FLAGS, *[f for f in flags.split(",") if f], "-c", "-o",
                        str(o), str(B.CSRC / "kernels_radon.cu")], capture_output=True, text=True)
    if r.returncode:
        print(r.stdout, r.stderr)
        raise SystemExit(1)
    lines = (r.stdout + r.stderr).splitlines()
    info = [lines[i + 2].strip() + " | " + lines[i + 3].strip() for i, ln in enumerate(lines)
            if KERNEL in ln and "Compiling" in ln]
    objs = [str(B.OBJ / (p.stem + ".o")) for p in B.sources()[0] if p.stem != "kernels_radon"] + [str(o)]
    so = out / f"libnutpie_b200_{tag}.so"
    subprocess.run([nvcc, "-ccbin", "/usr/bin/g++", "-shared", "-o", str(so), *objs, "-ldl"], check=True,
                   capture_output=True)
    n = subprocess.run(f"cuobjdump -sass {o} | awk '/Function : .*{KERNEL}/{{f=1;next}} /Function :/{{f=0}} f' | "
                       "grep -cE '^\\s+/\\*[0-9a-f]{4,}\\*/'", shell=True, capture_output=True, text=True).stdout.strip()
    o.unlink()
    return tag, info, n


with ThreadPoolExecutor(4) as ex:
    for tag, info, n in ex.map(one, sys.argv[1:]):
        print(tag, "|", *info, "| SASS instructions:", n)
