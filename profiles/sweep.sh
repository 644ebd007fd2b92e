#!/bin/bash
# rebuild with different -D knobs on the GPU box and time the main workloads
for extra in "$@"; do
  PLB_NVCC_EXTRA="$extra" python unsupervised-pseuso-lidar_b200/plb200/build.py --force > /dev/null 2>&1 || { echo "build failed: $extra"; continue; }
  echo "== $extra"
  profiles/quick_bench.sh
done
