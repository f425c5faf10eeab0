#!/bin/bash
# Build tuning variants of libmbcol.so (worker warps per scan CTA) into csrc/variants/ for A/B timing on the GPU box.
set -e
cd "$(dirname "$0")/../minibase-columnar-database_b200/csrc"
mkdir -p variants
for ww in "$@"; do
  d=variants/w$ww; mkdir -p $d
  for f in mbc_api mbc_scan mbc_synth mbc_bitmap mbc_join mbc_ingest; do
    /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr -DMBC_WORKER_WARPS=$ww -c $f.cu -o $d/$f.o &
  done
  wait
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libmbcol_w$ww.so $d/*.o -lcudart
done
ls -la variants/*.so
