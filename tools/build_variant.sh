#!/bin/bash
# A/B builds of libroargraph_b200.so: tools/build_variant.sh <name> [nvcc -D flags...] [--search-src FILE]
# writes mysteryann_b200/variants/<name>.so; select it at run time with RG_B200_LIB=<path>.
set -e
name=$1; shift
src=mysteryann_b200/csrc/rg_search.cu
defs=()
while [ $# -gt 0 ]; do
  if [ "$1" == "--search-src" ]; then src=$2; shift 2; else defs+=("$1"); shift; fi
done
out=mysteryann_b200/variants; mkdir -p $out/obj_$name
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-O2 --expt-relaxed-constexpr -ccbin /usr/bin/g++ -Imysteryann_b200/csrc"
/usr/local/cuda/bin/nvcc $FLAGS "${defs[@]}" -c $src -o $out/obj_$name/rg_search.o 2>&1 | grep -v deprecated || true
objs=""
for f in rg_index rg_knn rg_knn_sharded rg_build; do objs="$objs mysteryann_b200/csrc/$f.o"; done
/usr/local/cuda/bin/nvcc -shared -o $out/$name.so $out/obj_$name/rg_search.o $objs -ccbin /usr/bin/g++ -ldl 2>&1 | grep -v deprecated || true
ls -la $out/$name.so
