#!/bin/bash
# Host side of libartb.so (scene mirror, flattener, .obj / PNG readers, graph import, the C ABI and its validation) rebuilt
# with -fsanitize=address,undefined and run under the CPU test suite (ARTB_LIB selects the library). The device object is
# the normal build's. libstdc++ is preloaded next to libasan: python does not link it, and ASan's __cxa_throw interceptor
# needs it resolved before the first C++ throw inside the library.
set -e
cd "$(dirname "$0")/../another_raytracer_b200/csrc"
out=${ASAN_OUT:-/tmp/artb_asan}; mkdir -p $out
for f in host/scene.cpp host/scene_manager.cpp host/mesh.cpp host/imageio.cpp host/graph_dump.cpp host/flatten.cpp abi/artb_abi.cpp; do
  g++ -O1 -g -std=gnu++17 -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer -I../../include -I/usr/local/cuda/include -c $f -o $out/$(basename $f .cpp).o
done
g++ -shared -Wl,-Bsymbolic -fsanitize=address,undefined -o $out/libartb_asan.so $out/*.o build/device/kernels.o -L/usr/local/cuda/lib64 -lcudart_static -lz -ldl -lrt -lpthread
cd ../..
ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD="$(g++ -print-file-name=libasan.so) $(g++ -print-file-name=libstdc++.so)" ARTB_LIB=$out/libartb_asan.so \
  python -m pytest tests -q -m "not gpu" -p no:cacheprovider "$@"
