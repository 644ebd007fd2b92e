"""Config C4 alone: python profiles/cloud_bench.py  (the `cloud` / `velo` objects of bench.py's JSON line)."""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unsupervised-pseuso-lidar_b200")]
import torch  # noqa: E402
import bench  # noqa: E402
dev = torch.device("cuda:0")
peak = bench.hbm_peak()[0] if hasattr(bench, "hbm_peak") else 6534.1
c = bench.time_cloud(dev)
c["frac"] = c["achieved_gbs"] / peak
c["f32_pointcloud2"]["frac"] = c["f32_pointcloud2"]["achieved_gbs"] / peak
print(json.dumps({"cloud": c}))
if "--velo" in sys.argv:
    v = bench.time_velo(dev)
    v["frac"] = v["achieved_gbs"] / peak
    print(json.dumps({"velo": v}))
