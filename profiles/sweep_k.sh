#!/bin/bash
# rebuild with different -D knobs on the GPU box and time the photometric launch alone (kbench.py)
#   profiles/sweep_k.sh "workloads" "-DA=1" "-DB=2 -DC=3" ...
wls="$1"; shift
for extra in "$@"; do
  PLB_NVCC_EXTRA="$extra" python unsupervised-pseuso-lidar_b200/plb200/build.py --force > /dev/null 2>&1 || { echo "build failed: $extra"; continue; }
  echo "== $extra"
  python profiles/kbench.py $wls 2>&1 | grep kernel
done
