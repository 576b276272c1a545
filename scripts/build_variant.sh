#!/bin/bash
# A/B builds: scripts/build_variant.sh NAME [-DFLAG ...]  ->  mammo-clip_b200/lib/libmclip_NAME.so  (select with MCLIP_LIB=<path>)
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
OUT=mammo-clip_b200/lib/variant_$NAME
mkdir -p $OUT
for f in mammo-clip_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=default --expt-relaxed-constexpr "$@" -c $f -o $OUT/$b.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o mammo-clip_b200/lib/libmclip_$NAME.so $OUT/*.o -gencode arch=compute_100a,code=sm_100a
rm -rf $OUT
echo mammo-clip_b200/lib/libmclip_$NAME.so
