#!/bin/bash
# A/B of compile-time variants of preprocess.cu on the GPU box against BASELINE config 4 (1280x720 sequences, device-resident frames).
for v in "$@"; do
  DVO_NVCC_EXTRA="$v" python -c "
import os
from rgbd_odometry_b200 import build as b
os.utime(os.path.join(b.CSRC, 'preprocess.cu'))
b.build_cuda()" > /dev/null 2>&1
  echo "variant [$v]: $(python tools/bench_sequence.py 148 8 16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['device'], d['pinned_host']['frame_pairs_per_s'])")"
done
