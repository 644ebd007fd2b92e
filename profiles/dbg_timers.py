"""In-kernel stage timestamps (%globaltimer) of the fused photometric launch; build with
PLB_NVCC_EXTRA=-DPLB_DEBUG_TIMERS.  Prints stage offsets (us) relative to the main kernel's start."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unsupervised-pseuso-lidar_b200")]
import torch
import bench
from plb200 import synth, ops, _lib
wl = sys.argv[1] if len(sys.argv) > 1 else "headline"
cfg = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
sets = [synth.to_device(s, dev) for s in bench.make_sets(cfg, 4, 1234, dev)]
if os.environ.get("DBG_SAME_IMAGES"):
    # every image of the batch identical (frames, disparity, pose, intrinsics): any remaining spread between the
    # blocks' finishing times is scheduling, not content
    def same(t):
        t[:] = t[:1]
    for s_ in sets:
        same(s_["tgt"]); same(s_["poses"]); same(s_["intrinsics"])
        for r in s_["ref_imgs"]:
            same(r)
        for fr in s_["disparity"]:
            for d in fr:
                same(d)
from losses import Losses
for rep in range(3):
    ms = bench.time_photo_kernel(Losses(), sets, cfg, dev, 64)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 32)()
_lib.lib.plb_debug_timers.argtypes = [C.c_void_p]
rc = _lib.lib.plb_debug_timers(buf)
t = [int(x) for x in buf]
t0 = t[0]
names = {0: "main blk0 start", 1: "main blk0 prologue done", 2: "main blk0 units done", 3: "main blk0 end",
         4: "main blkN start", 5: "main blkN prologue done", 6: "main blkN units done", 7: "main blkN end",
         8: "fin start (t0)", 9: "fin start (prep thread)", 10: "fin prep done", 11: "fin after wait (t0)",
         12: "fin after wait (prep)", 16: "fin loads issued", 17: "fin first load back", 18: "fin pair loads back", 13: "fin loads done", 14: "fin barrier 1", 15: "fin end"}
print("kernel_ms %.4f" % ms)
for k in sorted(names):
    print("%-28s %8.2f us" % (names[k], (t[k] - t0) / 1e3))

blk = (C.c_ulonglong * 2048)()
_lib.lib.plb_debug_block_ends.argtypes = [C.c_void_p]
_lib.lib.plb_debug_block_ends(blk)
ends = [(int(x) - t0) / 1e3 for x in blk if int(x) > 0]
ends = [e for e in ends if -100 < e < 1000]
n = len(ends)
print("blocks", n, "units-done min %.2f  p10 %.2f  median %.2f  p90 %.2f  max %.2f" % (
    min(ends), sorted(ends)[n // 10], sorted(ends)[n // 2], sorted(ends)[9 * n // 10], max(ends)))
step = max(1, n // 37)
print("by block index:", " ".join("%.1f" % e for e in ends[::step]))

sm = (C.c_uint * 2048)()
_lib.lib.plb_debug_block_sms.argtypes = [C.c_void_p]
_lib.lib.plb_debug_block_sms(sm)
per_sm = {}
for i in range(n):
    per_sm.setdefault(int(sm[i]), []).append((i, ends[i]))
print("blocks per SM:", sorted(set(len(v) for v in per_sm.values())), "n_sm", len(per_sm))
rows = sorted((max(e for _, e in v), k, [i for i, _ in v]) for k, v in per_sm.items())
print("fastest SMs:", [(k, round(t, 1), ids) for t, k, ids in rows[:6]])
print("slowest SMs:", [(k, round(t, 1), ids) for t, k, ids in rows[-6:]])
lo = [t for t, k, _ in rows if k < 74]; hi = [t for t, k, _ in rows if k >= 74]
print("SM<74 mean end %.2f  SM>=74 mean end %.2f" % (sum(lo) / max(len(lo), 1), sum(hi) / max(len(hi), 1)))
import collections
bysm = sorted((k, round(max(e for _, e in v), 1)) for k, v in per_sm.items())
print("end by smid:", " ".join("%d:%.0f" % kv for kv in bysm))
