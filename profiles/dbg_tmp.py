import sys; sys.path[:0]=['/root/repo','/root/repo/unsupervised-pseuso-lidar_b200','/root/repo/tests']
import torch
from helpers import load_golden, golden_inputs, rel_err
from losses import Losses
for name in ["live_b4_s1_32x48","live_b4_s4_32x64","live_b2_s2_32x48_patched"]:
    g = load_golden(name)
    tgt, refs, disparity, poses, K = golden_inputs(g)
    dev='cuda'
    with torch.no_grad():
        loss = Losses().forward(tgt.to(dev), [r.to(dev) for r in refs], [[d.to(dev) for d in fr] for fr in disparity], poses.to(dev), K.to(dev), None)
    print(name, float(loss[0]), float(g['loss_mam']), float(loss[1]), float(g['loss_smooth']))
    disp=[[d.to(dev).requires_grad_(True) for d in fr] for fr in disparity]; p=poses.to(dev).requires_grad_(True)
    loss = Losses().forward(tgt.to(dev), [r.to(dev) for r in refs], disp, p, K.to(dev), None)
    sum(loss).backward()
    print('  grad loss', float(loss[0]), 'pose err', rel_err(p.grad.cpu(), g['g_poses']), [rel_err(t.grad.cpu(), g['g_disp_f%d_s%d'%(f,s)]) for f,fr in enumerate(disp) for s,t in enumerate(fr)])
