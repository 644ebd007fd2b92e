# A/B of cloud.cu build variants on one box: profiles/cloud_ab.sh "-DX=1" "" ...
for extra in "$@"; do
  PLB_NVCC_EXTRA="$extra" python unsupervised-pseuso-lidar_b200/plb200/build.py --force > /dev/null 2>&1 || { echo "build failed: $extra"; continue; }
  echo "== $extra"
  for i in 1 2; do python profiles/cloud_kbench.py 2>&1 | grep cloud; done
done
