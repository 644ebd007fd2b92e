"""GPU parity of the reference's DORMANT path:
SSIM.standard_loss, compute_photometric_loss (+clip), and the fused
min-reprojection + automask composition, against golden vectors made from the
reference's own functions (tests/golden/make_golden.py::dormant_case) and the
oracle on shapes the reference cannot run."""
import pytest
import torch

from helpers import load_golden, golden_inputs, rel_err, max_rel_err

pytestmark = pytest.mark.gpu

LOSS_TOL = 1e-5
GRAD_TOL = 1e-4
MAP_TOL = 2e-5     # max |delta| of a photometric map whose values lie in [0, 1] (+0.15*|diff|)


def _dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def test_ssim_map_golden():
    from losses import SSIM
    g = load_golden("dormant_b4_s2_32x48")
    tgt, refs, _, _, _ = golden_inputs(g)
    out = SSIM().standard_loss(refs[0].to(_dev()), tgt.to(_dev()))
    assert float((out.cpu() - torch.from_numpy(g["ssim_ref0_tgt"])).abs().max()) < MAP_TOL


@pytest.mark.parametrize("no_ssim,key", [(False, "photo_clip_ref0_tgt"), (True, "photo_clip_nossim_ref0_tgt")])
def test_photometric_map_clip_golden(no_ssim, key):
    from losses import Losses
    g = load_golden("dormant_b4_s2_32x48")
    tgt, refs, _, _, _ = golden_inputs(g)
    out = Losses().compute_photometric_loss(refs[0].to(_dev()), tgt.to(_dev()), no_ssim=no_ssim)
    ref = torch.from_numpy(g[key])
    assert float((out.cpu() - ref).abs().max()) < MAP_TOL
    # the clamp is active: a visible share of the map sits at the threshold
    assert float((out == out.max()).float().mean()) > 0.01


@pytest.mark.parametrize("B,C,H,W", [(1, 3, 2, 2), (2, 3, 17, 33), (3, 1, 40, 70)])
@pytest.mark.parametrize("variant", ["ssim", "photo", "photo_clip", "l1_clip"])
def test_photometric_map_backward_vs_oracle(B, C, H, W, variant):
    """vjp with respect to BOTH images against torch autograd through the oracle (CPU)."""
    from plb200 import ops
    from oracle import restated as O
    gen = torch.Generator().manual_seed(B * 100 + H)
    x = torch.randn(B, C, H, W, generator=gen)
    y = (x + 0.3 * torch.randn(B, C, H, W, generator=gen)).contiguous()
    go = torch.randn(B, C, H, W, generator=gen)

    def run(fn_ssim, fn_photo, xx, yy):
        if variant == "ssim":
            return fn_ssim(xx, yy)
        if variant == "photo":
            return fn_photo(xx, yy, False, None)
        if variant == "photo_clip":
            return fn_photo(xx, yy, False, 0.5)
        return fn_photo(xx, yy, True, 0.5)

    rx, ry = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    r = run(O.ssim_standard_loss, lambda a, b, n, c: O.compute_photometric_loss(a, b, n, c), rx, ry)
    (r * go).sum().backward()
    dev = _dev()
    gx, gy = x.to(dev).requires_grad_(True), y.to(dev).requires_grad_(True)
    o = run(ops.ssim_map, lambda a, b, n, c: ops.photometric_map(a, b, no_ssim=n, clip=c), gx, gy)
    (o * go.to(dev)).sum().backward()
    assert float((o.detach().cpu() - r.detach()).abs().max()) < MAP_TOL
    if "clip" in variant:
        # an element within rounding of the threshold may fall on the other side of the clamp: compare
        # where both sides agree about being clipped (all but a handful)
        thr = float(r.max())
        keep = ((r.detach() - thr).abs() > 1e-5)
        assert float(keep.float().mean()) > 0.5
    assert rel_err(gx.grad.cpu(), rx.grad) < 5e-4
    assert rel_err(gy.grad.cpu(), ry.grad) < 5e-4


def _run_min(tgt, refs, disp_scales, poses, K, automask, no_ssim=False, clip=None, image_grads=False,
             deterministic=None, upstream=1.0):
    """disparity -> depth (our kernel) -> fused min-reprojection loss; gradients back to disparity (and, with
    image_grads, to the frames: the tensors are returned as the 4th / 5th value)."""
    from losses import Losses
    from geometry.pose_geometry import disp_to_depth
    dev = _dev()
    disp = [d.to(dev).requires_grad_(True) for d in disp_scales]
    p = poses.to(dev).requires_grad_(True)
    t = tgt.to(dev).requires_grad_(image_grads)
    r = [x.to(dev).requires_grad_(image_grads) for x in refs]
    depth = disp_to_depth([disp])[0]
    loss = Losses(deterministic=deterministic).multiview_reprojection_loss(t, r, depth, p, K.to(dev), automask=automask,
                                                                           no_ssim=no_ssim, clip=clip)
    (upstream * loss).backward()
    if image_grads:
        return loss, disp, p, t, r
    return loss, disp, p


@pytest.mark.parametrize("automask", [True, False])
@pytest.mark.parametrize("variant", ["noclip", "clip"])
def test_min_reprojection_golden(automask, variant):
    """Golden composition of the reference's own functions (make_golden.py::dormant_case): `clip` = through
    compute_photometric_loss with its mean + 0.5 std clamp per map (losses.py:79-82), the way
    notes/toy_problem/losses.py:107-129 composes it; `noclip` = the same without the clamp.  Loss, pose, disparity and
    source-image gradients."""
    g = load_golden("dormant_b4_s2_32x48")
    tgt, refs, disparity, poses, K = golden_inputs(g)
    tag = "%s_%s" % (variant, "auto" if automask else "noauto")
    clip = 0.5 if variant == "clip" else None
    loss, disp, p, t, r = _run_min(tgt, refs, disparity[0], poses, K, automask, clip=clip, image_grads=True)
    errs = {"loss": abs(float(loss) - float(g["loss_" + tag])) / abs(float(g["loss_" + tag])),
            "poses": rel_err(p.grad.cpu(), g["g_poses_" + tag])}
    for s, d in enumerate(disp):
        errs["disp%d" % s] = rel_err(d.grad.cpu(), g["g_disp_s%d_%s" % (s, tag)])
    for i in range(2):
        errs["ref%d" % i] = rel_err(r[i].grad.cpu(), g["g_ref%d_%s" % (i, tag)])
    print(tag, {k: "%.2e" % v for k, v in errs.items()})
    assert errs.pop("loss") <= LOSS_TOL
    assert max(errs.values()) < GRAD_TOL, errs
    # without image gradients (the single-pass variant) the other gradients are the same
    loss2, disp2, p2 = _run_min(tgt, refs, disparity[0], poses, K, automask, clip=clip)
    assert abs(float(loss2) - float(loss)) <= 1e-6 * abs(float(loss))
    assert rel_err(p2.grad, p.grad) < 1e-6


@pytest.mark.parametrize("clip", [None, 0.5])
def test_min_reprojection_image_grads_vs_oracle(clip):
    """Gradients with respect to the TARGET and the source frames against autograd through the oracle composition
    (the golden vectors hold source gradients only), float-atomic and deterministic accumulation; the deterministic
    mode is bitwise repeatable."""
    from plb200 import synth
    from oracle import restated as O
    inp = synth.make_photo_inputs(3, 40, 70, n_src=2, n_scales=2, seed=411)
    tgt, refs, poses, K = inp["tgt"], inp["ref_imgs"], inp["poses"], inp["intrinsics"]
    rt = tgt.double().requires_grad_(True)
    rr = [x.double().requires_grad_(True) for x in refs]
    rd = [d.double().requires_grad_(True) for d in inp["disparity"][0]]
    rl = O.min_reprojection_loss(rt, rr, O.disp_to_depth([rd])[0], poses.double(), K, automask=True, clip_loss=clip)
    (0.6 * rl).backward()
    outs = []
    for det in (False, True, True):
        loss, disp, p, t, r = _run_min(tgt, refs, inp["disparity"][0], poses, K, True, clip=clip, image_grads=True,
                                       deterministic=det, upstream=0.6)
        errs = [rel_err(t.grad.cpu(), rt.grad)] + [rel_err(r[i].grad.cpu(), rr[i].grad) for i in range(2)] + \
               [rel_err(d.grad.cpu(), x.grad) for d, x in zip(disp, rd)]
        print("clip", clip, "det", det, "rel err tgt / ref0 / ref1 / disp:", ["%.2e" % e for e in errs])
        assert abs(float(loss) - float(rl)) <= 2 * LOSS_TOL * abs(float(rl))
        # fp32 kernel against the fp64 evaluation of the reference's formulas: selections that tie to fp32 rounding
        # (min / mask / channel max / clamp) move single elements
        assert max(errs[:3]) < GRAD_TOL, errs
        outs.append([t.grad.clone()] + [x.grad.clone() for x in r])
    for a, b in zip(outs[1], outs[2]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("B,H,W,S,automask,no_ssim", [(1, 9, 35, 1, True, False), (3, 50, 131, 3, True, False),
                                                      (2, 64, 128, 2, False, False), (2, 40, 70, 2, True, True),
                                                      (2, 192, 640, 4, True, False)])
def test_min_reprojection_vs_oracle(B, H, W, S, automask, no_ssim):
    """Oracle composition (oracle/restated.py::min_reprojection_loss) on shapes the reference cannot
    run.  min / automask / channel-max are selections: a pixel whose two candidates tie to fp32
    rounding may select differently, which moves single gradient elements, so gradient maps are held
    to 1e-4 of their scale on all but a small, counted number of elements (as in the live tests)."""
    from plb200 import synth
    from oracle import restated as O
    inp = synth.make_photo_inputs(B, H, W, n_src=2, n_scales=S, seed=300 + B + H)
    tgt, refs, poses, K = inp["tgt"], inp["ref_imgs"], inp["poses"], inp["intrinsics"]

    def oracle(dtype):
        c = lambda t: t.to(dtype)
        rd = [c(d).clone().requires_grad_(True) for d in inp["disparity"][0]]
        rp = c(poses).clone().requires_grad_(True)
        rdepth = O.disp_to_depth([rd])[0]
        rl = O.min_reprojection_loss(c(tgt), [c(r) for r in refs], rdepth, rp, K, automask=automask, no_ssim=no_ssim)
        rl.backward()
        return rl, rp, rd
    rl, rp, rd = oracle(torch.float32)
    _, rp64, rd64 = oracle(torch.float64)
    loss, disp, p = _run_min(tgt, refs, inp["disparity"][0], poses, K, automask, no_ssim)
    assert abs(float(loss) - float(rl)) <= 2 * LOSS_TOL * abs(float(rl))
    # pose gradients sum over every pixel, so each flipped selection moves them: the bar is a small
    # multiple of the fp32 reference's own distance from an fp64 evaluation of the same formulas
    e32 = rel_err(rp.grad, rp64.grad)
    assert rel_err(p.grad.cpu(), rp64.grad) < max(GRAD_TOL, 3 * e32), e32
    assert rel_err(p.grad.cpu(), rp.grad) < max(GRAD_TOL, 4 * e32), e32
    # per-pixel gradient maps: elements off by more than 1e-4 of the map's scale are counted against
    # the fp64 evaluation; the budget is a multiple of the fp32 reference's own count (a flipped
    # min / mask / max selection or bilinear cell moves the ~16 low-resolution elements it feeds)
    for s, (a, b32, b64) in enumerate(zip(disp, rd, rd64)):
        x = b64.grad
        scale = float(x.abs().max())
        bad = int(((a.grad.cpu().double() - x).abs() > GRAD_TOL * scale).sum())
        bad_ref = int(((b32.grad.double() - x).abs() > GRAD_TOL * scale).sum())
        assert bad <= max(16, int(5e-4 * x.numel())) + 4 * bad_ref, (s, bad, bad_ref, x.numel())


def test_min_reprojection_repeatable_and_forward_only():
    from losses import Losses
    from plb200 import synth
    inp = synth.make_photo_inputs(2, 48, 100, n_src=2, n_scales=2, seed=77)
    outs = []
    for _ in range(3):
        loss, disp, p = _run_min(inp["tgt"], inp["ref_imgs"], inp["disparity"][0], inp["poses"], inp["intrinsics"], True)
        outs.append((loss.detach().clone(), p.grad.clone(), [d.grad.clone() for d in disp]))
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1])
        for a, b in zip(o[2], outs[0][2]):
            assert torch.equal(a, b)
    dev = _dev()
    with torch.no_grad():
        depth = [1.0 / (10.0 * d.to(dev) + 0.01) for d in inp["disparity"][0]]
        l2 = Losses().multiview_reprojection_loss(inp["tgt"].to(dev), [r.to(dev) for r in inp["ref_imgs"]], depth,
                                                  inp["poses"].to(dev), inp["intrinsics"].to(dev), clip=None)
    assert abs(float(l2) - float(outs[0][0])) <= 1e-5 * abs(float(l2))
