#!/bin/bash
# Builds latent2im_b200/lib/libl2i_b200_<name>.so with extra nvcc flags applied to the listed sources (others reuse build/*.o).
#   tools/build_variant.sh <name> "<flags>" file1.cu [file2.cu ...]       then:  python tools/ab_libs.py a=... b=...
set -e
name=$1; flags=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
csrc=$root/latent2im_b200/csrc
vdir=$root/build/var_$name
mkdir -p "$vdir"
make -s -C "$csrc" -j16 >/dev/null 2>&1
objs=""
for o in "$root"/build/*.o; do
  b=$(basename "$o" .o); skip=0
  for f in "$@"; do [ "$b.cu" = "$f" ] && skip=1; done
  [ $skip = 0 ] && objs="$objs $o"
done
for f in "$@"; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
    $flags -c "$csrc/$f" -o "$vdir/${f%.cu}.o" 2>/dev/null
  objs="$objs $vdir/${f%.cu}.o"
done
/usr/local/cuda/bin/nvcc -shared -o "$root/latent2im_b200/lib/libl2i_b200_$name.so" $objs -lcudart_static -lpthread -ldl -lrt 2>/dev/null
echo "$root/latent2im_b200/lib/libl2i_b200_$name.so"
