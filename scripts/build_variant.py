"""Build a variant of the library with extra -D flags for chosen translation units:
python scripts/build_variant.py <tag> <tu1,tu2> <flags...>  ->  nutpie_b200/variants/libnutpie_b200_<tag>.so
(select it with NB200_LIB=<path>)."""
import subprocess, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from nutpie_b200 import build as B

tag, tus, flags = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
B.build()
out = B.PKG / "variants"; out.mkdir(exist_ok=True)
nvcc = B._nvcc()
objs = [str(B.OBJ / (p.stem + ".o")) for p in B.sources()[0] if p.stem not in tus]
for stem in tus:
    o = out / f"{stem}_{tag}.o"
    r = subprocess.run([nvcc, "-ccbin", "/usr/bin/g++", *B.NVCC_FLAGS, *flags, "-c", "-o", str(o), str(B.CSRC / f"{stem}.cu")],
                       capture_output=True, text=True)
    if r.returncode:
        print(r.stdout[-3000:], r.stderr[-3000:]); raise SystemExit(1)
    objs.append(str(o))
so = out / f"libnutpie_b200_{tag}.so"
subprocess.run([nvcc, "-ccbin", "/usr/bin/g++", "-shared", "-o", str(so), *objs, "-ldl"], check=True)
print("built", so)
