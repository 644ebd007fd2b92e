# A/B of edge.cu build variants on one box: profiles/edge_ab.sh "-DX=1" "" ...
for extra in "$@"; do
  PLB_NVCC_EXTRA="$extra" python unsupervised-pseuso-lidar_b200/plb200/build.py --force > /dev/null 2>&1 || { echo "build failed: $extra"; continue; }
  echo "== $extra"
  python profiles/trace_step.py c5e 10 2>&1 | grep -E "per replay|edge_main"  | head -2
  python profiles/edge_bench.py 2>&1 | tail -2
done
