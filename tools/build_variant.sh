#!/bin/sh
# Build a tuning variant of the library: tools/build_variant.sh NAME [-DFLAG ...]  ->  tools/_variants/lib_NAME.so
# (run it with DIRECT_DDP_LIB=tools/_variants/lib_NAME.so python tools/cycle_report.py ...)
name=$1; shift
exec nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC "$@" \
    -o tools/_variants/lib_$name.so direct_b200/csrc/direct_ddp.cu direct_b200/host/corridor_replay.cpp
