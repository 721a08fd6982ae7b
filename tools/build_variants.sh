#!/bin/bash
# A/B builds of libdiffrp_b200.so with other -D switches: tools/build_variants.sh NAME "-DFOO=1 -DBAR=2" ...
# writes build/variants/NAME.so (git-ignored, travels with gpurun); select one with DIFFRP_B200_LIB=build/variants/NAME.so.
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --extended-lambda -Xcompiler -fPIC -shared -Xptxas -v $flags \
    -o build/variants/$name.so diffrp_b200/csrc/api.cu diffrp_b200/csrc/wavefront.cu diffrp_b200/csrc/flatten.cu \
    diffrp_b200/csrc/epilogue.cu diffrp_b200/csrc/conv3x3.cu > build/variants/$name.ptxas.log 2>&1 &
done
wait
ls -la build/variants
