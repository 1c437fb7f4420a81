#!/bin/bash
# k_walk with a 32-entry candidate list and a step started by any candidate: the list-overflow path of walk_round under the parity suites
ARTB_LIB=$PWD/another_raytracer_b200/libartb_items32.so timeout 900 python -m pytest tests/test_gpu_walk_parity.py tests/test_gpu_bvh_build.py -x -q -m gpu 2>&1 | tail -3
ARTB_LIB=$PWD/another_raytracer_b200/libartb_items32.so ARTB_TWO_PHASE=0 timeout 900 python -m pytest tests/test_gpu_trace_parity.py -x -q -m gpu 2>&1 | tail -3
