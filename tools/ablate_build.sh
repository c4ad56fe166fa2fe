#!/bin/bash
# Builds ablation variants of libsdab ("What bounds the level-0 launches"):
#   tools/ablate_build.sh 1 16 32 64 128     -> sda_b200/build/libsdab_abl<bits>.so
# bits: 1 no MMA issue, 16 no global epilogue operands, 32 no epilogue stores, 64 staging written but no TMA store
# issued, 128 no F output.  The variants compute WRONG results by design; run them with
#   SDAB_NO_BUILD=1 SDAB_LIB=sda_b200/build/libsdab_abl32.so python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-secondary
# (the non-finite-state assertion of bench.py fires for some of them: read the ncu launch list instead).
set -e
cd "$(dirname "$0")/.."
python -m sda_b200.build > /dev/null
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr --expt-extended-lambda"
OBJS=$(ls sda_b200/build/*.cu.o | grep -v conv_umma.cu.o)
for b in "$@"; do
  nvcc $FLAGS -DSDAB_ABLATE=$b -c sda_b200/csrc/conv_umma.cu -o /tmp/conv_umma_abl$b.o
  nvcc -shared -o sda_b200/build/libsdab_abl$b.so $OBJS /tmp/conv_umma_abl$b.o -cudart static -Xcompiler -fPIC
  echo "sda_b200/build/libsdab_abl$b.so"
done
