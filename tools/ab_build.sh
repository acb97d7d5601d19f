#!/bin/bash
# A/B builds of libvcb200 for kernel experiments: the same sources with one macro flipped, written next to the product
# library (git-ignored) and selected at run time with VCB_LIB_PATH.
#   tools/ab_build.sh plain -DVCB_SILU_PLAIN    ->  vehicle_counting_b200/libvcb200_plain.so
set -e
tag=$1; shift
cd "$(dirname "$0")/../vehicle_counting_b200/csrc"
tmp=$(mktemp -d)
for f in conv_umma pointwise detect_nms roi reid_stem reid_stem_direct capi; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f.cu -o $tmp/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libvcb200_$tag.so $tmp/*.o -lcudart
rm -rf $tmp
echo built libvcb200_$tag.so
