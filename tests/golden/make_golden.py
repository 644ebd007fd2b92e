"""Generates the committed golden vectors under tests/golden/ by running the
UNMODIFIED reference (`/root/reference`, imported through
`oracle/reference_shim.py`) on CPU, torch 2.11.0, in the build container.

    python tests/golden/make_golden.py

The GPU box has no reference tree, so these files are what pins the oracle
(`oracle/restated.py`) and, through it, the CUDA path.  Inputs are stored in
the fixtures next to the outputs, so they do not depend on the generator.
"""
import hashlib
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "unsupervised-pseuso-lidar_b200"))

from plb200 import synth  # noqa: E402
from oracle import reference_shim  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def _flatten_inputs(inp):
    d = {"tgt": _np(inp["tgt"]), "poses": _np(inp["poses"]), "K": _np(inp["intrinsics"])}
    for i, r in enumerate(inp["ref_imgs"]):
        d["ref%d" % i] = _np(r)
    for f, frame in enumerate(inp["disparity"]):
        for s, t in enumerate(frame):
            d["disp_f%d_s%d" % (f, s)] = _np(t)
    return d


def live_case(ref, name, B, H, W, n_scales, seed, regime):
    """`Losses.forward` + `sum(loss).backward()` exactly as `trainer.py:312,264`."""
    inp = synth.make_photo_inputs(B, H, W, n_src=2, n_scales=n_scales, seed=seed, regime=regime)
    disp = [[t.clone().requires_grad_(True) for t in frame] for frame in inp["disparity"]]
    poses = inp["poses"].clone().requires_grad_(True)
    tgt = inp["tgt"].clone().requires_grad_(True)
    refs = [r.clone().requires_grad_(True) for r in inp["ref_imgs"]]
    L = ref.Losses()
    with reference_shim.quiet():
        loss = L.forward(tgt, refs, disp, poses, inp["intrinsics"], None)
    sum(loss).backward()
    out = _flatten_inputs(inp)
    out["loss_mam"] = _np(loss[0])
    out["loss_smooth"] = _np(loss[1])
    out["g_poses"] = _np(poses.grad)
    out["g_tgt"] = _np(tgt.grad)
    for i, r in enumerate(refs):
        out["g_ref%d" % i] = _np(r.grad)
    for f, frame in enumerate(disp):
        for s, t in enumerate(frame):
            out["g_disp_f%d_s%d" % (f, s)] = _np(t.grad)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, float(loss[0]), float(loss[1]))


def warp_case(ref, name, B, H, W, seed):
    """`inverse_warp` forward and vjp against a fixed cotangent, both pose_inv."""
    inp = synth.make_photo_inputs(B, H, W, n_src=2, n_scales=1, seed=seed)
    gen = torch.Generator().manual_seed(seed + 1)
    cot = torch.randn(B, 3, H, W, generator=gen)
    out = {"K": _np(inp["intrinsics"]), "cot": _np(cot)}
    depth0 = 1.0 / (10 * inp["disparity"][0][0] + 0.01)
    out["img"] = _np(inp["ref_imgs"][0])
    out["depth"] = _np(depth0)
    out["pose"] = _np(inp["poses"][:, 0])
    for inv in (False, True):
        img = inp["ref_imgs"][0].clone().requires_grad_(True)
        depth = depth0.clone().requires_grad_(True)
        pose = inp["poses"][:, 0].clone().requires_grad_(True)
        proj = ref.inverse_warp(img, depth, pose, inp["intrinsics"], inv)
        (proj * cot).sum().backward()
        tag = "inv" if inv else "fwd"
        out["proj_" + tag] = _np(proj)
        out["g_img_" + tag] = _np(img.grad)
        out["g_depth_" + tag] = _np(depth.grad)
        out["g_pose_" + tag] = _np(pose.grad)
    # pose helper functions
    rot, trans = inp["poses"][:, 0, :3].unsqueeze(1), inp["poses"][:, 0, 3:].unsqueeze(1)
    M = ref.transformation_from_parameters(rot, trans)
    out["M_axisangle"] = _np(M)
    out["M_inverted"] = _np(ref.invert_pose(M))
    out["M_euler"] = _np(ref.pose_vec2mat(inp["poses"][:, 0], "euler"))
    # Transform.reconstruct / project
    t = ref.Transform()
    Xc = t.reconstruct(depth0[:, 0], inp["intrinsics"])
    out["Xc"] = _np(Xc)
    out["grid"] = _np(t.project(Xc, inp["intrinsics"], M))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok")


def dormant_case(ref, name, B, H, W, n_scales, seed):
    """SSIM, photometric mix (+clip), and the min-reprojection/automask
    composition assembled from the reference's own functions the way
    `notes/toy_problem/losses.py:107-129` composes them."""
    inp = synth.make_photo_inputs(B, H, W, n_src=2, n_scales=n_scales, seed=seed)
    L = ref.Losses()
    L.SSIM = ref.SSIM()
    tgt, refs, K = inp["tgt"], inp["ref_imgs"], inp["intrinsics"]
    out = _flatten_inputs(inp)
    out["ssim_ref0_tgt"] = _np(L.SSIM.standard_loss(refs[0], tgt))
    out["photo_clip_ref0_tgt"] = _np(L.compute_photometric_loss(refs[0], tgt))
    out["photo_clip_nossim_ref0_tgt"] = _np(L.compute_photometric_loss(refs[0], tgt, no_ssim=True))

    def photo_noclip(pred, target, no_ssim=False):
        l1 = torch.abs(target - pred)
        return l1 if no_ssim else 0.85 * L.SSIM.standard_loss(pred, target) + 0.15 * l1

    import torch.nn.functional as F
    for variant, photo in (("clip", L.compute_photometric_loss), ("noclip", photo_noclip)):
        for automask in (True, False):
            disp = [t.clone().requires_grad_(True) for t in inp["disparity"][0]]
            poses = inp["poses"].clone().requires_grad_(True)
            srcs = [r.clone().requires_grad_(True) for r in refs]
            depths = ref.disp_to_depth([disp])[0]
            total = 0
            auto = [photo(r, tgt) for r in srcs]
            for D in depths:
                if D.shape[-1] != W:
                    D = F.interpolate(D, [H, W], mode="bilinear", align_corners=False)
                D = D.squeeze(1)
                rp = [photo(ref.inverse_warp(r, D, poses[:, i, :], K, False), tgt)
                      for i, r in enumerate(srcs)]
                min_rpl = torch.minimum(rp[0], rp[1])
                if automask:
                    min_auto = torch.minimum(auto[0], auto[1])
                    mu = torch.where(min_rpl < min_auto, torch.tensor([1.]), torch.tensor([0.]))
                    min_rpl = mu * min_rpl
                m, _ = torch.max(min_rpl, dim=1)
                total = total + m.mean(1).mean(-1).mean()
            total = total / len(depths)
            total.backward()
            tag = "%s_%s" % (variant, "auto" if automask else "noauto")
            out["loss_" + tag] = _np(total)
            out["g_poses_" + tag] = _np(poses.grad)
            for s, t in enumerate(disp):
                out["g_disp_s%d_%s" % (s, tag)] = _np(t.grad)
            for i, r in enumerate(srcs):
                out["g_ref%d_%s" % (i, tag)] = _np(r.grad)
            print(name, tag, float(total))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def cloud_case(ref, name):
    """`PseudoLiDAR.project_PL` on a small depth image (full cloud stored) and on
    one full-resolution KITTI frame (count, checksums and a strided sample)."""
    with tempfile.TemporaryDirectory() as d:
        calib = synth.write_kitti_calib(d)
        out = {}
        small = synth.make_depth_images(1, 37, 124, seed=77, lo=0.5, hi=60.0)[0].numpy()
        small[5, 7] = 0.0
        small[6, 9] = -2.0
        for sp in (0, 3):
            pl = ref.PseudoLiDAR(calib, sp)
            cloud = pl.project_PL(small)
            out["small_cloud_sp%d" % sp] = cloud
        out["small_depth"] = small
        out["T"], out["P"] = pl.T, pl.P
        full = synth.make_depth_images(1, 375, 1242, seed=78)[0].numpy()
        pl = ref.PseudoLiDAR(calib, 0)
        cloud = pl.project_PL(full)
        out["full_seed"] = np.array(78)
        out["full_count"] = np.array(cloud.shape[0])
        out["full_colsum"] = cloud.sum(axis=0)
        out["full_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(cloud).tobytes()).digest(), dtype=np.uint8)
        out["full_sample"] = cloud[::997].copy()
        cloud10 = ref.PseudoLiDAR(calib, 10).project_PL(full)
        out["full_count_sp10"] = np.array(cloud10.shape[0])
        print(name, cloud.shape, cloud10.shape, cloud.dtype)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def velo_case(ref_unused, name):
    """`Transform.project_velo_to_img` (pseudo-lidar/Transform/Transform.py:69-104, the unmodified class) on a
    small sweep projected into a 124x37 image (P scaled by 0.1: many cells hit several times, so the
    later-point-wins rule is exercised; inputs and full output stored) and on one KITTI-sized sweep into
    1242x375 (regenerated from the seed; count, checksum and the non-zero cells stored)."""
    Transform = reference_shim.load_velo_transform()
    out = {}
    with tempfile.TemporaryDirectory() as d:
        calib = synth.write_kitti_calib(d)
        tr = Transform(calib, 124, 37)
        tr.P = tr.P * np.array([[0.1], [0.1], [1.0]])
        small = synth.make_velodyne_cloud(3000, seed=91)
        small[7] = [np.nan, 1.0, 1.0, 0.0]          # NaN fails every test
        small[8] = [0.0, 0.0, 0.0, 0.0]             # x > 0 fails; also uv/0
        small[9] = [150.0, 0.5, -1.0, 0.0]          # beyond the 120 m cut
        out["small_points"] = small
        out["small_P"], out["T"] = tr.P.copy(), tr.T.copy()
        out["small_depth"] = tr.project_velo_to_img(small)
        tr = Transform(calib, 1242, 375)
        out["P"] = tr.P.copy()
        full = synth.make_velodyne_cloud(123577, seed=92)
        depth = tr.project_velo_to_img(full)
        out["full_seed"], out["full_n"] = np.array(92), np.array(123577)
        nz = np.flatnonzero(depth)
        out["full_nz_index"], out["full_nz_value"] = nz.astype(np.int32), depth.reshape(-1)[nz]
        out["full_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(depth).tobytes()).digest(), dtype=np.uint8)
        print(name, out["small_depth"].shape, int((out["small_depth"] != 0).sum()), depth.shape, nz.size, depth.dtype)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def head_case(name):
    """The disparity head `alpha * sigmoid(conv) + beta` of the UNMODIFIED `models/depth/disp_net.py::DispNetS`
    (`:121-139`): the pre-activation of two scales (forward hooks on the convolution in front of the sigmoid) and
    the disparities the network returns for them."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_disp_net", os.path.join(reference_shim.REFERENCE_ROOT, "models", "depth", "disp_net.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(3)
    net = mod.DispNetS().train()
    cap = {}
    net.predict_disp1[0].register_forward_hook(lambda m, i, o: cap.__setitem__("x1", o.detach().clone()))
    net.predict_disp3[0].register_forward_hook(lambda m, i, o: cap.__setitem__("x3", o.detach().clone()))
    with torch.no_grad():
        out = net(torch.randn(2, 3, 64, 128))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x1=cap["x1"].numpy(), disp1=out[0].numpy(),
                        x3=cap["x3"].numpy(), disp3=out[2].numpy())
    print(name, [tuple(o.shape) for o in out])


def make_frame(h, w, seed):
    """Synthetic decoded camera frame, uint8 HWC: smooth colour field + texture + noise, every value of 0..255 used."""
    rng = np.random.default_rng(seed)
    v = np.arange(h, dtype=np.float64)[:, None, None]
    u = np.arange(w, dtype=np.float64)[None, :, None]
    ph = np.array([0.0, 2.1, 4.2])[None, None, :]
    x = 127.5 + 100.0 * np.sin(2 * np.pi * u / 97.0 + ph) * np.cos(2 * np.pi * v / 61.0) + 40.0 * rng.standard_normal((h, w, 3))
    x[: h // 8] = rng.integers(0, 256, (h // 8, w, 3))          # a band of pure noise: worst case for the filter
    return np.clip(np.rint(x), 0, 255).astype(np.uint8)


def prep_case(name):
    """`KittiDataset.load_img` (dataloaders.py:32-49, the UNMODIFIED method) with the transform list of
    `trainer.py:97-103` on PNG files written here, and the intrinsics scaling of `dataloaders.py:95-98`."""
    from PIL import Image
    from torchvision import transforms
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, __import__("types").ModuleType(m))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    for k in [k for k in sys.modules if k == "geometry" or k.startswith("geometry.")]:
        del sys.modules[k]                                     # the reference's own geometry package for its star-imports
    sys.path.insert(0, reference_shim.REFERENCE_ROOT)
    import dataloaders as ref_dl
    assert ref_dl.__file__.startswith(reference_shim.REFERENCE_ROOT)

    def chain(img_height, img_width):
        # trainer.py:97-103, verbatim structure
        return [transforms.ToTensor(), transforms.ToPILImage(), transforms.Resize((img_height, img_width)),
                transforms.ToTensor(), transforms.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))]

    class Self:                                                # what load_img reads from the dataset object
        pass
    out = {}
    cases = {"down": (94, 311, 48, 160), "same": (48, 160, 48, 160), "up": (30, 50, 48, 80), "odd": (53, 97, 20, 33)}
    with tempfile.TemporaryDirectory() as d:
        for tag, (h0, w0, H, W) in cases.items():
            frame = make_frame(h0, w0, seed=hash(tag) % 1000 if False else {"down": 1, "same": 2, "up": 3, "odd": 4}[tag])
            path = os.path.join(d, tag + ".png")
            Image.fromarray(frame).save(path)
            me = Self()
            me.transforms = chain(H, W)
            img, oh, ow = ref_dl.KittiDataset.load_img(me, path)
            assert (oh, ow) == (h0, w0)
            out[tag + "_frame"], out[tag + "_out"] = frame, _np(img)
            out[tag + "_size"] = np.array([H, W])
            K = synth.kitti_intrinsics(1, h0, w0)[0].numpy().copy()
            out[tag + "_K_in"] = K.copy()
            K[0] *= W / ow                                     # dataloaders.py:95-97
            K[1] *= H / oh
            out[tag + "_K_out"] = K
        # full KITTI size -> network resolution: checksum + two rows of every channel
        frame = make_frame(375, 1242, seed=5)
        path = os.path.join(d, "full.png")
        Image.fromarray(frame).save(path)
        me = Self()
        me.transforms = chain(192, 640)
        img = _np(ref_dl.KittiDataset.load_img(me, path)[0])
        out["full_seed"] = np.array(5)
        out["full_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(img).tobytes()).digest(), dtype=np.uint8)
        out["full_rows"] = img[:, [0, 95, 191], :]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if k.endswith("_out")})


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "prep":
        prep_case("prep_frames")
        return
    torch.manual_seed(0)
    torch.set_num_threads(1)  # fixed reduction order for the committed vectors
    ref = reference_shim.load(patch_batch=False)
    live_case(ref, "live_b4_s1_32x48", 4, 32, 48, 1, seed=11, regime="trained")
    live_case(ref, "live_b4_s4_32x64", 4, 32, 64, 4, seed=12, regime="trained")
    live_case(ref, "live_b4_s1_init_24x40", 4, 24, 40, 1, seed=13, regime="init")
    warp_case(ref, "warp_b4_24x40", 4, 24, 40, seed=21)
    dormant_case(ref, "dormant_b4_s2_32x48", 4, 32, 48, 2, seed=31)
    cloud_case(ref, "cloud_kitti")
    velo_case(ref, "velo_kitti")
    head_case("head_dispnet")
    prep_case("prep_frames")
    ref = reference_shim.load(patch_batch=True)
    live_case(ref, "live_b2_s2_32x48_patched", 2, 32, 48, 2, seed=14, regime="trained")
    live_case(ref, "live_b3_s1_24x40_patched", 3, 24, 40, 1, seed=15, regime="trained")


if __name__ == "__main__":
    main()
