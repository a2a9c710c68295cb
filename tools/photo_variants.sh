#!/bin/bash
# A/B of compile-time variants of photometric.cu on the GPU box: rebuilds the library with each -D set and times config 3.
for v in "$@"; do
  DVO_NVCC_EXTRA="$v" python -c "
import os
from rgbd_odometry_b200 import build as b
os.utime(os.path.join(b.CSRC, 'photometric.cu'))
b.build_cuda()" > /dev/null 2>&1
  echo "variant [$v]: $(python tools/bench_photometric.py 1024 6 2>/dev/null | tail -1 | cut -c100-200)"
done
