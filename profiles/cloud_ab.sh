# A/B of cloud.cu build variants on one box: profiles/cloud_ab.sh "-DX=1" "" ...
for extra in "$@"; do
  PLB_NVCC_EXTRA="$extra" python unsupervised-pseuso-lidar_b200/plb200/build.py --force > /dev/null 2>&1 || { echo "build failed: $extra"; continue; }
  echo "== $extra"
  for i in 1 2; do python profiles/cloud_bench.py 2>&1 | tail -1 | python -c "
import json,sys
c=json.loads(sys.stdin.read())['cloud']
print('f64 %.1f us frac %.3f | f32 %.1f us' % (c['ms']*1e3, c['frac'], c['f32_pointcloud2']['ms']*1e3))"; done
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cloud_ -s 4 -c 2 python profiles/prof_cloud.py 2>&1 | grep "gpu__time_duration"
done
