#!/bin/bash
# Experiment helper: builds variants of the library that differ in the attention translation unit only (extra nvcc flags per
# variant), into variants/lib_<tag>.so.  Select one at run time with MAPANYTHING_B200_LIB=variants/lib_<tag>.so.
#   tools/build_variants.sh tag1 "flags1" tag2 "flags2" ...
set -e
cd "$(dirname "$0")/.."
python map-anything_b200/build.py > /dev/null
mkdir -p variants
objs=$(ls map-anything_b200/build/*.o | grep -v "/attention.o")
while [ $# -gt 0 ]; do
  tag=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
       -Imap-anything_b200/csrc -Iinclude $flags -c map-anything_b200/csrc/attention.cu -o variants/attention_$tag.o
  nvcc -shared -o variants/lib_$tag.so variants/attention_$tag.o $objs -gencode arch=compute_100a,code=sm_100a \
       -cudart shared -Xlinker -rpath=/usr/local/cuda/lib64
  echo "built variants/lib_$tag.so"
done
