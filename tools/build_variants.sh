#!/usr/bin/env bash
# Kernel A/B variants: rebuilds ONE translation unit with extra -D flags and links it with the other (unchanged) objects
# into turbo-whisper-workspace_b200/variants/libtwb200_<name>.so (git-ignored; travels to the GPU box).
#   tools/build_variants.sh attention_enc max3 "-DATTN_MAX3" poly4 "-DATTN_POLY_EVERY=4" ...
# Select a variant at run time with TWB200_LIB=<path>.
set -eu
cd "$(dirname "$0")/../turbo-whisper-workspace_b200/csrc"
make > /dev/null
unit=$1; shift
mkdir -p ../variants build/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC,-Wall,-Wno-unknown-pragmas --expt-relaxed-constexpr -Xptxas -v $flags \
       -c $unit.cu -o build/variants/${unit}_${name}.o 2> build/variants/${unit}_${name}.ptxas.log
  objs=$(ls build/*.o | grep -v "build/${unit}.o")
  nvcc $ARCH -shared -o ../variants/libtwb200_${name}.so $objs build/variants/${unit}_${name}.o
  echo "built variants/libtwb200_${name}.so ($flags): $(grep -c spill build/variants/${unit}_${name}.ptxas.log) spill lines"
done
