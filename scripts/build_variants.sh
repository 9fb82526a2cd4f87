#!/bin/bash
# Development helper: one library per combination of compile-time knobs of one kernel source, under build/variants/
# (git-ignored, shipped to the GPU box by gpurun).  usage: scripts/build_variants.sh <source.cu> name:"-DA=1 -DB=2" ...
# Run `make` first: the other objects are taken from turbosqueeze_b200/csrc/*.o.
set -e
cd "$(dirname "$0")/.."
src=$1; shift
stem=$(basename "$src" .cu)
mkdir -p build/variants
others=$(ls turbosqueeze_b200/csrc/*.o | grep -v "/$stem.o")
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $flags -c "$src" -o build/variants/$name.o &&
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_$name.so build/variants/$name.o $others -lpthread &&
    rm build/variants/$name.o && echo "built $name ($flags)" ) &
done
wait
