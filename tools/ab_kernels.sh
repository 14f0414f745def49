#!/bin/bash
# A/B several builds of libsstts on the GPU box with the kernel probes: tools/ab_kernels.sh lib1.so lib2.so ...
for lib in "$@"; do
  echo "== $lib"
  SSTTS_LIB=$PWD/$lib python tools/kernel_probe.py 2>&1 | tail -1
  SSTTS_LIB=$PWD/$lib python tools/gl1024_probe.py 2>&1 | tail -1
done
