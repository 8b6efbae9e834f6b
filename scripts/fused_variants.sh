#!/bin/bash
# Build-time experiment: how many U-turn partners to fuse into the streaming pass (register pressure
# vs HBM traffic).  Rebuilds only kernels_normal.cu per variant, on the GPU box.
for k in 0 1 2 3; do
  touch nutpie_b200/csrc/kernels_normal.cu
  NB200_EXTRA_NVCC_FLAGS="-DNB200_MAX_FUSED=$k" python -m nutpie_b200.build > /dev/null 2>&1
  grep -E "nuts_kernelINS_11NormalModelELi8ELi0E" -A3 nutpie_b200/build.log | grep -E "spill" | head -1
  echo "NB200_MAX_FUSED=$k"; python scripts/gpu_sweep.py cfg4 2>&1 | grep "tpc=256"
done
