#!/bin/sh
# Developer A/B: build copies of the library with different compile-time knobs into build_variants/
# (git-ignored; travels to the GPU box), then e.g.
#   JMCODEC_B200_LIB=build_variants/libjmc_minb10.so python tools/odd_sizes.py
set -e
cd "$(dirname "$0")/.."
mkdir -p build_variants
for v in "$@"; do
    name=$(echo "$v" | tr -c 'A-Za-z0-9\n' '_')
    rm -rf build_variants/obj_$name; mkdir -p build_variants/obj_$name
    for f in jmc_kernels jmc_runtime jm_nv_dec jmnv_enc; do
        /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden \
            -Iinclude -Ijmcodec_b200/csrc $v -c jmcodec_b200/csrc/$f.cu -o build_variants/obj_$name/$f.o &
    done
    /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden -Iinclude -x cu \
        -c jmcodec_b200/csrc/jmc_annexb.cpp -o build_variants/obj_$name/jmc_annexb.o &
    wait
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build_variants/libjmc$name.so build_variants/obj_$name/*.o -ldl -lpthread
    rm -rf build_variants/obj_$name
    echo built build_variants/libjmc$name.so
done
