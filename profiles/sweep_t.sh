#!/bin/bash
# rebuild with different -D knobs and print one kernel's line of the in-graph step timeline: sweep_t.sh workload kernel-substring "-DX=1" ...
wl="$1"; pat="$2"; shift; shift
for extra in "$@"; do
  PLB_NVCC_EXTRA="$extra" python unsupervised-pseuso-lidar_b200/plb200/build.py --force > /dev/null 2>&1 || { echo "build failed: $extra"; continue; }
  echo "== $extra"
  python profiles/trace_step.py $wl 20 2>&1 | grep -E "^workload|  .*$pat" | head -2
done
