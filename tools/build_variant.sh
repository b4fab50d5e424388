#!/bin/bash
# build an alternative libcmb200 with extra -D flags:  tools/build_variant.sh TAG -DFOO=1 ...
# -> alt/libcmb200_TAG.so (use with CMB200_SO=alt/libcmb200_TAG.so)
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
mkdir -p alt
env -u CC -u CXX nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" \
    -o alt/libcmb200_$tag.so contact_map_b200/csrc/cmb_api.cu
echo alt/libcmb200_$tag.so
