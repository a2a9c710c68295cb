#!/bin/bash
# A/B of compile-time variants of preprocess.cu on the GPU box: rebuilds the library with each -D set and times the stages.
# usage: tools/prepare_variants.sh "<flags of variant 1>" "<flags of variant 2>" ...
for v in "$@"; do
  DVO_NVCC_EXTRA="$v" python -c "
import os
from rgbd_odometry_b200 import build as b
os.utime(os.path.join(b.CSRC, 'preprocess.cu'))
b.build_cuda()" > /dev/null 2>&1
  echo "variant [$v]: $(python tools/prepare_quick.py 1024 2>&1 | tail -1)"
done
