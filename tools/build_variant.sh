#!/bin/bash
# usage: tools/build_variant.sh NAME [-DDEFINE ...]   -> variants/libsrb_NAME.so (A/B measurements; SYNCHRAD_B200_LIB selects it)
set -e
name=$1; shift
mkdir -p variants
cd synchrad_b200/csrc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -diag-suppress 177 "$@" -o ../../variants/libsrb_$name.so srb_api.cu
