#!/bin/bash
# usage: tools/build_variant.sh NAME "-DFOO=1 -DBAR=2"   — builds another_raytracer_b200/libartb_NAME.so with extra nvcc flags
# (experiment helper: tools/bench_variants.sh then benches every libartb*.so on the GPU box)
set -e
cd "$(dirname "$0")/../another_raytracer_b200/csrc"
name=$1; shift
mkdir -p build/variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I../../include \
    --expt-relaxed-constexpr -fmad=false -Xptxas -v $@ -c device/kernels.cu -o build/variants/kernels_$name.o 2> build/variants/kernels_$name.log
/usr/local/cuda/bin/nvcc -shared -Xlinker -Bsymbolic -o ../libartb_$name.so build/host/*.o build/abi/*.o build/variants/kernels_$name.o -lz -cudart static 2>/dev/null
grep -E "k_traverseILb1ELb0|k_confirmILb0" -A2 build/variants/kernels_$name.log | grep -E "Used|spill" | tr '\n' ' '; echo
