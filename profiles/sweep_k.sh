#!/bin/bash
# rebuild with different -D knobs on the GPU box and time kernel-only workloads: sweep_k.sh "wl1 wl2" "-DX=1" "-DX=2" ...
wls="$1"; shift
for extra in "$@"; do
  PLB_NVCC_EXTRA="$extra" python unsupervised-pseuso-lidar_b200/plb200/build.py --force > /dev/null 2>&1 || { echo "build failed: $extra"; continue; }
  echo "== $extra"
  python profiles/kbench.py $wls 2>&1 | tail -n $(echo $wls | wc -w)
done
