"""torch-facing host layer over the C ABI: autograd Functions and workspaces.

PyTorch is plumbing here (device memory, streams, autograd graph); every
computation is a libplb200.so kernel.  All ops require CUDA tensors and raise
otherwise - there is no CPU path in the product.

Gradient strategy (DESIGN.md "single-pass forward+backward"): the loss is a
scalar whose gradients are linear in the upstream gradient g.  When gradients
are needed the forward launch therefore also writes the gradients for g = 1
(same memory pass: inputs are read once).  backward() relaunches the same
kernels with the real upstream values and a device-side guard that makes the
launch return at once when every upstream value is exactly 1 (the common
`sum(loss).backward()`); any other upstream recomputes the gradients exactly.
`fused_backward=False` (or image gradients being requested) selects the
classic two-pass scheme: forward computes only the loss, backward recomputes.
"""
import threading

import torch

from . import _lib
from ._lib import lib, check

_ws_lock = threading.Lock()
_ws_cache = {}


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(device=None):
    """Raw handle of torch's current stream on `device` (default: the current device) - the fast private accessor
    when this torch has it: `torch.cuda.current_stream()` costs ~18 us of Python per call, and a loss step asks
    six times."""
    idx = torch.cuda.current_device() if device is None or device.index is None else device.index
    if _raw_stream is not None:
        return _raw_stream(idx)
    return torch.cuda.current_stream(idx).cuda_stream


def _workspace(tag, nbytes, device):
    """Zero-filled once, then self-cleaning (kernels reset their tickets).
    One buffer per (op, device, stream) so concurrent streams never share."""
    key = (tag, device.index, _stream())
    with _ws_lock:
        buf = _ws_cache.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            _ws_cache[key] = buf
    return buf


_sm_limit = 0


def set_sm_limit(n):
    """Size the persistent grid of the fused loss for `n` SMs (0 = all of them).  A data-parallel step that runs an
    NCCL all-reduce beside the loss leaves the collective's CTAs their SMs this way: they need whole SMs, and a
    persistent grid that fills the GPU makes them wait for its blocks to retire."""
    global _sm_limit
    _sm_limit = int(n)


def _need_cuda(*tensors):
    """Every tensor must live on the CURRENT CUDA device: the kernels are launched on that device's stream, and a
    launch with another device's pointers would fault (or, with peer access, silently run on the wrong GPU)."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("plb200 ops need CUDA tensors (no CPU fallback); got a %s tensor" % t.device.type)
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise RuntimeError("plb200 ops launch on the current device (cuda:%d) but got a tensor on %s; wrap the "
                               "call in `with torch.cuda.device(t.device):`" % (cur, t.device))


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _kc(K):
    if K.dtype not in (torch.float64, torch.float32):
        K = K.double()
    return K.contiguous()


def _ptr(t):
    return 0 if t is None else t.data_ptr()


class LossConfig:
    """Static description of one fused-loss call (shapes, modes, weights)."""

    def __init__(self, n_src, scales_per_frame, input_is_depth=False, do_photo=True, do_smooth=True,
                 rotation_mode="axisangle", fused_backward=True, disp_a=10.0, disp_b=0.01, scale_decay=2.3,
                 mode=_lib.PHOTO_L1_MEAN, flags=0, disp_head=None, deterministic=None, clip_loss=None, edge=False):
        self.n_src = n_src
        # edge=True: the smoothness term of the step is the edge-aware first-order one (north_star's variant; not in the
        # reference) on the target frame's disparity pyramid - inside the same call, accumulated into the same gradient maps
        self.edge = bool(edge)
        self.scales_per_frame = list(scales_per_frame)  # e.g. [4, 4]: frames with a depth pyramid
        self.input_is_depth = bool(input_is_depth)
        self.do_photo, self.do_smooth = do_photo, do_smooth
        self.rotation_mode = _lib.ROT_EULER if rotation_mode == "euler" else _lib.ROT_AXISANGLE
        self.fused_backward = fused_backward
        self.disp_a, self.disp_b, self.scale_decay = disp_a, disp_b, scale_decay
        self.mode, self.flags = mode, flags
        # PHOTO_MIN_REPROJ: clamp every photometric map at mean + clip_loss * std (losses.py:79-82); None = no clamp
        self.clip_loss = None if clip_loss is None else float(clip_loss)
        if self.clip_loss is not None:
            self.flags |= _lib.PHOTO_CLIP
        # image gradients through order-independent fixed-point accumulation (everything else is always repeatable);
        # None follows torch.use_deterministic_algorithms()
        self.deterministic = torch.are_deterministic_algorithms_enabled() if deterministic is None else bool(deterministic)
        # (alpha, beta): the pyramids hold the disparity head's PRE-ACTIVATION x and disp = alpha * sigmoid(x) + beta
        # (models/depth/disp_net.py:121-139) is evaluated inside the kernels; gradients come back with respect to x
        self.disp_head = None if disp_head is None else (float(disp_head[0]), float(disp_head[1]))
        if self.disp_head is not None and self.input_is_depth:
            raise ValueError("disp_head applies to disparity inputs, not to depth")
        self.input_kind = _lib.INPUT_LOGIT if self.disp_head else (_lib.INPUT_DEPTH if self.input_is_depth else _lib.INPUT_DISP)


def _launch_loss(cfg, tgt, refs, poses, K, pyr, want_grad, g_pyr, g_poses, g_tgt, g_refs, out, up, skip):
    """One photometric launch (all directions) + one smoothness launch.
    out: float32[2] (loss_mam, loss_smooth).  up: None or (g_mam, g_smooth) 0-d float32
    CUDA tensors (either may be None = that part is inactive)."""
    up_ptr = [0, 0] if up is None else [_ptr(up[0]), _ptr(up[1])]
    dev = tgt.device
    B, _, H, W = tgt.shape
    st = _stream()
    if cfg.do_photo:
        a = _lib.PhotoArgs()
        a.B, a.H, a.W = B, H, W
        a.n_pose = poses.shape[1]
        a.rotation_mode = cfg.rotation_mode
        a.k_is_f64 = 1 if K.dtype == torch.float64 else 0
        a.input_is_depth = cfg.input_kind
        a.disp_a, a.disp_b = cfg.disp_a, cfg.disp_b
        if cfg.disp_head:
            a.head_alpha, a.head_beta = cfg.disp_head
        a.want_grad = int(want_grad)
        a.deterministic = int(cfg.deterministic)
        a.sm_limit = _sm_limit
        a.poses, a.K = poses.data_ptr(), K.data_ptr()
        a.g_poses = _ptr(g_poses) if want_grad else 0
        a.loss = out.data_ptr()
        a.upstream = up_ptr[0]
        if skip:
            a.skip_if_unit[0], a.skip_if_unit[1] = up_ptr[0], up_ptr[1]
        n_jobs = len(pyr)
        a.n_jobs = n_jobs
        entries = sum(len(p) for p in pyr)
        for j in range(n_jobs):
            job = a.jobs[j]
            if j == 0:
                job.tgt = tgt.data_ptr()
                job.n_src = len(refs)
                for i, r in enumerate(refs):
                    job.src[i] = r.data_ptr()
                    job.pose_index[i] = i
                    job.pose_inv[i] = 0
                    job.g_src[i] = _ptr(g_refs[i]) if (want_grad and g_refs) else 0
                job.g_tgt = _ptr(g_tgt) if want_grad else 0
            else:
                # losses.py:199-203: target = refs[indx], source = [tgt], pose = poses[indx-1] inverted
                job.tgt = refs[j].data_ptr()
                job.n_src = 1
                job.src[0] = tgt.data_ptr()
                job.pose_index[0] = j - 1
                job.pose_inv[0] = 1
                job.g_src[0] = _ptr(g_tgt) if want_grad else 0
                job.g_tgt = _ptr(g_refs[j]) if (want_grad and g_refs) else 0
            job.n_scales = len(pyr[j])
            for s, d in enumerate(pyr[j]):
                job.disp[s] = d.data_ptr()
                job.dh[s], job.dw[s] = d.shape[-2], d.shape[-1]
                job.g_disp[s] = _ptr(g_pyr[j][s]) if (want_grad and g_pyr is not None) else 0
            job.term_weight = 1.0 if cfg.mode == _lib.PHOTO_MIN_REPROJ else 1.0 / (entries * job.n_src)
            job.mode, job.flags = cfg.mode, cfg.flags
            job.clip_loss = cfg.clip_loss if cfg.clip_loss is not None else 0.0
        nbytes = lib.plb_photo_workspace_bytes(a)
        ws_photo = _workspace("photo", nbytes, dev)
        a.workspace, a.workspace_bytes = ws_photo.data_ptr(), ws_photo.numel()
        check(lib.plb_photo_loss(a, st), "plb_photo_loss")
    if cfg.do_smooth:
        s_ = _lib.SmoothArgs()
        s_.B = B
        s_.n_scales = len(pyr[0])
        for s, d in enumerate(pyr[0]):
            s_.disp[s] = d.data_ptr()
            s_.dh[s], s_.dw[s] = d.shape[-2], d.shape[-1]
            s_.g_disp[s] = _ptr(g_pyr[0][s]) if (want_grad and g_pyr is not None) else 0
        s_.accumulate = 1 if cfg.do_photo else 0
        s_.input_is_depth = cfg.input_kind
        s_.disp_a, s_.disp_b, s_.scale_decay = cfg.disp_a, cfg.disp_b, cfg.scale_decay
        if cfg.disp_head:
            s_.head_alpha, s_.head_beta = cfg.disp_head
        s_.want_grad = int(want_grad)
        s_.loss = out.data_ptr() + 4
        s_.upstream = up_ptr[1]
        if skip:
            s_.skip_if_unit[0], s_.skip_if_unit[1] = up_ptr[0], up_ptr[1]
        nbytes = lib.plb_smooth_workspace_bytes(s_)
        ws_smooth = _workspace("smooth", nbytes, dev)
        s_.workspace, s_.workspace_bytes = ws_smooth.data_ptr(), ws_smooth.numel()
        check(lib.plb_smooth_loss(s_, st), "plb_smooth_loss")
    # the workspace tensors travel with the argument structs: a later, larger call replaces the cached buffer, and
    # the guarded relaunch of THIS call must keep writing into memory that is still its own
    return ((a if cfg.do_photo else None), (s_ if cfg.do_smooth else None), st,
            (ws_photo if cfg.do_photo else None, ws_smooth if cfg.do_smooth else None))


_unit_upstream = False


class unit_upstream:
    """Context: the caller vouches that every upstream gradient of the fused losses evaluated inside is exactly 1 - a
    step that calls `(loss[0] + loss[1]).backward()` itself (plb200/graphed.py).  The gradients written by the forward
    launch then stand, and backward issues no guarded relaunch (five launches that would exit at once)."""

    def __enter__(self):
        global _unit_upstream
        self._old = _unit_upstream
        _unit_upstream = True
        from . import _tb
        if _tb.mod is not None and hasattr(_tb.mod, "set_unit_upstream"):
            _tb.mod.set_unit_upstream(True)
        return self

    def __exit__(self, *exc):
        global _unit_upstream
        _unit_upstream = self._old
        from . import _tb
        if _tb.mod is not None and hasattr(_tb.mod, "set_unit_upstream"):
            _tb.mod.set_unit_upstream(self._old)
        return False


def _relaunch_guarded(args, up, scratch):
    """The backward pass of a fused forward: the SAME launches again (same buffers, same stream) with the real
    upstream scalars behind the device-side "all upstream == 1" guard.  The argument structs of the forward call
    are reused; only the upstream / guard / loss pointers change - rebuilding them costs ~80 us of Python."""
    a, s_, st0, _keepalive = args
    st = _stream()
    if st != st0:
        return False                                   # another stream: other workspaces - take the general path
    up_ptr = [_ptr(up[0]), _ptr(up[1])]
    if a is not None:
        a.want_grad = 1
        a.loss, a.upstream = scratch.data_ptr(), up_ptr[0]
        a.skip_if_unit[0], a.skip_if_unit[1] = up_ptr[0], up_ptr[1]
        check(lib.plb_photo_loss(a, st), "plb_photo_loss")
    if s_ is not None:
        s_.want_grad = 1
        s_.loss, s_.upstream = scratch.data_ptr() + 4, up_ptr[1]
        s_.skip_if_unit[0], s_.skip_if_unit[1] = up_ptr[0], up_ptr[1]
        check(lib.plb_smooth_loss(s_, st), "plb_smooth_loss")
    return True


class FusedLossFn(torch.autograd.Function):
    """(cfg, tgt, poses, K, ref_0..ref_{n-1}, disp tensors frame-major) -> (loss_mam, loss_smooth)."""

    @staticmethod
    def forward(ctx, cfg, tgt, poses, K, *rest):
        refs = list(rest[:cfg.n_src])
        flat = list(rest[cfg.n_src:])
        pyr, k = [], 0
        for n in cfg.scales_per_frame:
            pyr.append(flat[k:k + n])
            k += n
        _need_cuda(tgt, poses, K, *refs, *flat)
        tgt, poses, K = _f32c(tgt), _f32c(poses), _kc(K)
        refs = [_f32c(r) for r in refs]
        pyr = [[_f32c(d) for d in p] for p in pyr]
        need = ctx.needs_input_grad
        img_grad = need[1] or any(need[4:4 + cfg.n_src])
        any_grad = img_grad or need[2] or any(need[4 + cfg.n_src:])
        fused = any_grad and cfg.fused_backward and not img_grad
        both = cfg.do_photo and cfg.do_smooth
        out = (torch.empty if both else torch.zeros)(2, dtype=torch.float32, device=tgt.device)
        g_pyr = g_poses = None
        if fused:
            g_pyr = [[torch.empty_like(d) for d in p] for p in pyr]
            g_poses = (torch.empty_like if cfg.do_photo else torch.zeros_like)(poses)
            if not cfg.do_photo:
                for j in range(1, len(g_pyr)):
                    for g in g_pyr[j]:
                        g.zero_()
        ctx.args = _launch_loss(cfg, tgt, refs, poses, K, pyr, fused, g_pyr, g_poses, None, None, out, None, False)
        ctx.cfg, ctx.fused, ctx.any_grad, ctx.img_grad = cfg, fused, any_grad, img_grad
        ctx.tensors = (tgt, refs, poses, K, pyr, g_pyr, g_poses)
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_mam, g_smooth):
        cfg = ctx.cfg
        tgt, refs, poses, K, pyr, g_pyr, g_poses = ctx.tensors
        dev = tgt.device
        up = []
        for g, active in ((g_mam, cfg.do_photo), (g_smooth, cfg.do_smooth)):
            if not active:
                up.append(None)                      # inactive part: no launch reads it
            elif g is None:
                up.append(torch.zeros((), dtype=torch.float32, device=dev))
            else:
                up.append(g.detach().to(torch.float32).reshape(()).contiguous())
        g_tgt = g_refs = None
        if ctx.fused and not getattr(ctx, "used", False):
            # first backward: the buffers written by the forward launch are handed to autograd
            skip = True
            ctx.used = True
        else:
            skip = False
            g_pyr = [[torch.empty_like(d) for d in p] for p in pyr]
            g_poses = (torch.empty_like if cfg.do_photo else torch.zeros_like)(poses)
            if not cfg.do_photo:
                for j in range(1, len(g_pyr)):
                    for g in g_pyr[j]:
                        g.zero_()
            if ctx.img_grad:
                g_tgt = torch.zeros_like(tgt)
                g_refs = [torch.zeros_like(r) for r in refs]
        scratch = torch.empty(2, dtype=torch.float32, device=dev)
        args, ctx.args = getattr(ctx, "args", None), None
        if skip and _unit_upstream:
            pass                                            # ops.unit_upstream(): the forward launch's gradients stand
        elif not (skip and args is not None and _relaunch_guarded(args, up, scratch)):
            _launch_loss(cfg, tgt, refs, poses, K, pyr, True, g_pyr, g_poses, g_tgt, g_refs, scratch, up, skip)
        # the gradient buffers leave with autograd: holding on to them would make AccumulateGrad CLONE every one of
        # them into .grad (a device copy per tensor) instead of adopting the buffer
        ctx.tensors = (tgt, refs, poses, K, pyr, None, None)
        need = ctx.needs_input_grad
        grads = [None, g_tgt if need[1] else None, g_poses if need[2] else None, None]
        for i in range(cfg.n_src):
            grads.append(g_refs[i] if (need[4 + i] and g_refs is not None) else None)
        k = 4 + cfg.n_src
        for p in g_pyr:
            while p:
                g = p.pop(0)
                grads.append(g if need[k] else None)
                k += 1
            del g
        del g_pyr, g_poses, g_tgt, g_refs
        return tuple(grads)


def _check_loss_shapes(tgt, refs, pyramids, poses, K):
    """The kernels see pointers, not extents: every shape the reference's torch ops would have rejected (or broadcast)
    is rejected here, before a launch can read or write past a buffer.  (~6 us per call: plain size comparisons.)"""
    ts = tgt.shape
    if len(ts) != 4 or ts[1] != 3:
        raise ValueError("tgt must be [B,3,H,W], got %s" % (tuple(ts),))
    B = ts[0]
    n_ref = len(refs)
    if not 1 <= n_ref <= _lib.MAX_SRC:
        raise ValueError("1..%d reference images, got %d" % (_lib.MAX_SRC, n_ref))
    for r in refs:
        if r.shape != ts:
            raise ValueError("every reference image must have the target's shape %s, got %s" % (tuple(ts), tuple(r.shape)))
    n_pyr = len(pyramids)
    if not 1 <= n_pyr <= 1 + n_ref:
        raise ValueError("1..%d disparity pyramids (target frame first), got %d" % (1 + n_ref, n_pyr))
    for p in pyramids:
        if not 1 <= len(p) <= _lib.MAX_SCALES:
            raise ValueError("1..%d scales per pyramid, got %d" % (_lib.MAX_SCALES, len(p)))
        for d in p:
            ds = d.shape
            nd = len(ds)
            if nd < 3 or ds[0] != B or ds[-1] < 1 or ds[-2] < 1 or (nd == 4 and ds[1] != 1) or nd > 4:
                raise ValueError("a disparity map must be [%d,1,h,w], got %s" % (B, tuple(ds)))
    ps = poses.shape
    n_pose = n_ref if n_ref > n_pyr - 1 else n_pyr - 1
    if len(ps) != 3 or ps[0] != B or ps[1] < n_pose or ps[2] != 6:
        raise ValueError("poses must be [%d,>=%d,6], got %s" % (B, n_pose, tuple(ps)))
    ks = K.shape
    if len(ks) != 3 or ks[0] != B or ks[1] != 3 or ks[2] != 3:
        raise ValueError("intrinsics must be [%d,3,3], got %s" % (B, tuple(ks)))


def fused_losses(tgt, refs, pyramids, poses, K, binding=None, **cfg_kw):
    """pyramids: list[frame] of list[scale] of [B,1,h,w].  Returns (loss_mam, loss_smooth).
    binding: "torch" = the C++ torch binding (csrc/torch_binding.cpp: autograd node and argument structs in C++),
    "ctypes" = FusedLossFn above; None = the C++ one when it is built.  Same C ABI, same kernels, same results."""
    cfg = LossConfig(len(refs), [len(p) for p in pyramids], **cfg_kw)
    flat = [d for p in pyramids for d in p]
    _check_loss_shapes(tgt, refs, pyramids, poses, K)
    from . import _tb
    if binding == "torch" and _tb.mod is None:
        raise RuntimeError("the torch C++ binding is not built (plb200/build.py --torch)")
    if cfg.edge and (binding == "ctypes" or _tb.mod is None):
        raise RuntimeError("edge=True (the edge-aware term inside the fused step) needs the torch C++ binding")
    if binding != "ctypes" and _tb.mod is not None:
        head = cfg.disp_head or (0.0, 0.0)
        out = _tb.mod.fused_losses([tgt, poses, K, *refs, *flat], cfg.n_src, cfg.scales_per_frame, cfg.input_kind,
                                   cfg.do_photo, cfg.do_smooth, cfg.rotation_mode, cfg.fused_backward, cfg.disp_a,
                                   cfg.disp_b, cfg.scale_decay, cfg.mode, cfg.flags, head[0], head[1],
                                   cfg.deterministic, cfg.clip_loss or 0.0, _sm_limit, cfg.edge)
        return out[0], out[1]
    return FusedLossFn.apply(cfg, tgt, poses, K, *refs, *flat)


# ---------------------------------------------------------------------------
# inverse_warp and the small geometry ops
# ---------------------------------------------------------------------------
class InverseWarpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, depth, pose, K, pose_inv, rotation_mode):
        _need_cuda(img, depth, pose, K)
        img, depth, pose, K = _f32c(img), _f32c(depth), _f32c(pose), _kc(K)
        if img.dim() != 4 or img.shape[1] != 3:
            raise ValueError("img must be [B,3,H,W], got %s" % (tuple(img.shape),))
        B, _, H, W = img.shape
        if depth.numel() != B * H * W or depth.shape[0] != B or tuple(depth.shape[-2:]) != (H, W):
            raise ValueError("depth must be [%d,%d,%d], got %s" % (B, H, W, tuple(depth.shape)))
        if pose.numel() != B * 6 or tuple(K.shape) != (B, 3, 3):
            raise ValueError("pose must be [%d,6] and intrinsics [%d,3,3], got %s / %s" % (B, B, tuple(pose.shape), tuple(K.shape)))
        out = torch.empty_like(img)
        a = _lib.WarpArgs()
        a.B, a.H, a.W = B, H, W
        a.rotation_mode, a.pose_inv = rotation_mode, int(bool(pose_inv))
        a.k_is_f64 = 1 if K.dtype == torch.float64 else 0
        a.img, a.depth, a.pose, a.K, a.out = img.data_ptr(), depth.data_ptr(), pose.data_ptr(), K.data_ptr(), out.data_ptr()
        a.pose_stride = 6
        check(lib.plb_warp_forward(a, _stream()), "plb_warp_forward")
        ctx.save_for_backward(img, depth, pose, K)
        ctx.meta = (rotation_mode, int(bool(pose_inv)))
        return out

    @staticmethod
    def backward(ctx, g_out):
        img, depth, pose, K = ctx.saved_tensors
        B, _, H, W = img.shape
        g_out = _f32c(g_out)
        need = ctx.needs_input_grad
        g_img = torch.zeros_like(img) if need[0] else None
        g_depth = torch.empty_like(depth) if need[1] else None
        g_pose = torch.empty_like(pose) if need[2] else None
        a = _lib.WarpArgs()
        a.B, a.H, a.W = B, H, W
        a.rotation_mode, a.pose_inv = ctx.meta
        a.k_is_f64 = 1 if K.dtype == torch.float64 else 0
        a.img, a.depth, a.pose, a.K = img.data_ptr(), depth.data_ptr(), pose.data_ptr(), K.data_ptr()
        a.pose_stride = 6
        a.g_out, a.g_img, a.g_depth, a.g_pose = g_out.data_ptr(), _ptr(g_img), _ptr(g_depth), _ptr(g_pose)
        ws = _workspace("warp", lib.plb_warp_workspace_bytes(a), img.device)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        check(lib.plb_warp_backward(a, _stream()), "plb_warp_backward")
        return g_img, g_depth, g_pose, None, None, None


class PoseMatrixFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pose, rotation_mode, invert):
        _need_cuda(pose)
        pose = _f32c(pose)
        B = pose.shape[0]
        M = torch.empty(B, 4, 4, dtype=torch.float32, device=pose.device)
        check(lib.plb_pose_matrix(pose.data_ptr(), 6, B, rotation_mode, int(invert), M.data_ptr(), _stream()),
              "plb_pose_matrix")
        ctx.save_for_backward(pose)
        ctx.meta = (rotation_mode, int(invert))
        return M

    @staticmethod
    def backward(ctx, gM):
        (pose,) = ctx.saved_tensors
        gM = _f32c(gM)
        g = torch.empty_like(pose)
        check(lib.plb_pose_matrix_backward(pose.data_ptr(), 6, pose.shape[0], ctx.meta[0], ctx.meta[1],
                                           gM.data_ptr(), g.data_ptr(), _stream()), "plb_pose_matrix_backward")
        return g, None, None


class DispToDepthFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, a, b):
        _need_cuda(disp)
        d = _f32c(disp)
        out = torch.empty_like(d)
        check(lib.plb_disp_to_depth(d.data_ptr(), d.numel(), a, b, out.data_ptr(), _stream()), "plb_disp_to_depth")
        ctx.save_for_backward(d)
        ctx.ab = (a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        g = _f32c(g)
        out = torch.empty_like(d)
        check(lib.plb_disp_to_depth_backward(d.data_ptr(), g.data_ptr(), d.numel(), ctx.ab[0], ctx.ab[1],
                                             out.data_ptr(), _stream()), "plb_disp_to_depth_backward")
        return out, None, None


def reconstruct(depth, K):
    """[B,H,W] depth, [B,3,3] K -> Xc [B,3,H,W] (forward only)."""
    _need_cuda(depth, K)
    depth, K = _f32c(depth), _kc(K)
    B, H, W = depth.shape
    out = torch.empty(B, 3, H, W, dtype=torch.float32, device=depth.device)
    check(lib.plb_reconstruct(depth.data_ptr(), K.data_ptr(), int(K.dtype == torch.float64), B, H, W,
                              out.data_ptr(), _stream()), "plb_reconstruct")
    return out


def project(X, K, Tcw):
    """[B,3,H,W] points, [B,3,3] K, [B,4,4] Tcw -> grid [B,H,W,2] (forward only)."""
    _need_cuda(X, K, Tcw)
    X, K, Tcw = _f32c(X), _kc(K), _f32c(Tcw)
    B, _, H, W = X.shape
    out = torch.empty(B, H, W, 2, dtype=torch.float32, device=X.device)
    check(lib.plb_project(X.data_ptr(), K.data_ptr(), int(K.dtype == torch.float64), Tcw.data_ptr(), B, H, W,
                          out.data_ptr(), _stream()), "plb_project")
    return out


# ---------------------------------------------------------------------------
# stand-alone SSIM / photometric maps (the reference's dormant functions)
# ---------------------------------------------------------------------------
class PhotometricMapFn(torch.autograd.Function):
    """out = w_ssim * clamp((1 - SSIM(x, y)) / 2, 0, 1) + w_l1 * |y - x|, optional clip at mean + clip * std."""

    @staticmethod
    def _args(x, y, C1, C2, w_ssim, w_l1, clip, thr):
        a = _lib.PhotomapArgs()
        a.B, a.C, a.H, a.W = x.shape
        a.x, a.y = x.data_ptr(), y.data_ptr()
        a.C1, a.C2, a.w_ssim, a.w_l1 = C1, C2, w_ssim, w_l1
        a.clip = -1.0 if clip is None else float(clip)
        a.threshold = _ptr(thr)
        return a

    @staticmethod
    def forward(ctx, x, y, C1, C2, w_ssim, w_l1, clip):
        _need_cuda(x, y)
        x, y = _f32c(x), _f32c(y)
        if x.dim() != 4 or x.shape != y.shape:
            raise ValueError("expected two [B,C,H,W] images of the same shape")
        out = torch.empty_like(x)
        thr = torch.empty((), dtype=torch.float32, device=x.device) if clip is not None else None
        a = PhotometricMapFn._args(x, y, C1, C2, w_ssim, w_l1, clip, thr)
        a.out = out.data_ptr()
        if clip is not None:
            ws = _workspace("photomap", lib.plb_photometric_map_workspace_bytes(a), x.device)
            a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        check(lib.plb_photometric_map(a, _stream()), "plb_photometric_map")
        ctx.save_for_backward(x, y)
        ctx.meta = (C1, C2, w_ssim, w_l1, clip, thr)
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, y = ctx.saved_tensors
        C1, C2, w_ssim, w_l1, clip, thr = ctx.meta
        g_out = _f32c(g_out)
        need = ctx.needs_input_grad
        g_x = torch.empty_like(x) if need[0] else None
        g_y = torch.empty_like(y) if need[1] else None
        a = PhotometricMapFn._args(x, y, C1, C2, w_ssim, w_l1, clip, thr)
        a.g_out, a.g_x, a.g_y = g_out.data_ptr(), _ptr(g_x), _ptr(g_y)
        check(lib.plb_photometric_map_backward(a, _stream()), "plb_photometric_map_backward")
        return g_x, g_y, None, None, None, None, None


def ssim_map(x, y, C1=1e-4, C2=9e-4):
    """`SSIM.standard_loss` (losses.py:12-54)."""
    return PhotometricMapFn.apply(x, y, float(C1), float(C2), 1.0, 0.0, None)


def photometric_map(pred, target, no_ssim=False, clip=0.5, C1=1e-4, C2=9e-4):
    """`Losses.compute_photometric_loss` (losses.py:66-84); clip=None skips the clamp."""
    w = (0.0, 1.0) if no_ssim else (0.85, 0.15)
    return PhotometricMapFn.apply(pred, target, float(C1), float(C2), w[0], w[1], clip)


# ---------------------------------------------------------------------------
# edge-aware smoothness (not in the reference; north_star's kernel list)
# ---------------------------------------------------------------------------
class EdgeSmoothFn(torch.autograd.Function):
    """(tgt, normalize, disp_0..disp_{S-1}) -> scalar loss; gradients for the disparities only."""

    @staticmethod
    def _launch(tgt, disps, normalize, want_grad, g_disp, g_scratch, loss, upstream):
        a = _lib.EdgeArgs()
        a.B, _, a.H, a.W = tgt.shape
        a.n_scales = len(disps)
        a.tgt = tgt.data_ptr()
        for s, d in enumerate(disps):
            a.disp[s] = d.data_ptr()
            a.dh[s], a.dw[s] = d.shape[-2], d.shape[-1]
            a.g_disp[s] = _ptr(g_disp[s]) if want_grad else 0
            a.g_scratch[s] = _ptr(g_scratch[s]) if (want_grad and normalize) else 0
        a.accumulate, a.normalize, a.want_grad = 0, int(bool(normalize)), int(bool(want_grad))
        a.loss, a.upstream = loss.data_ptr(), _ptr(upstream)
        ws = _workspace("edge", lib.plb_edge_smooth_workspace_bytes(a), tgt.device)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        check(lib.plb_edge_smooth_loss(a, _stream()), "plb_edge_smooth_loss")

    @staticmethod
    def forward(ctx, tgt, normalize, *disps):
        _need_cuda(tgt, *disps)
        if len(disps) < 1 or len(disps) > _lib.MAX_SCALES:
            raise ValueError("1..%d scales" % _lib.MAX_SCALES)
        tgt = _f32c(tgt)
        disps = [_f32c(d) for d in disps]
        if tgt.dim() != 4 or tgt.shape[1] != 3:
            raise ValueError("tgt must be [B,3,H,W], got %s" % (tuple(tgt.shape),))
        for d in disps:
            if d.dim() < 3 or d.shape[0] != tgt.shape[0] or d.numel() != tgt.shape[0] * d.shape[-2] * d.shape[-1]:
                raise ValueError("a disparity map must be [%d,1,h,w], got %s" % (tgt.shape[0], tuple(d.shape)))
        loss = torch.empty((), dtype=torch.float32, device=tgt.device)
        # single pass: when a disparity needs a gradient the forward launch writes d loss / d disp for a unit upstream
        # (the loss is a scalar, its gradient is linear in the upstream value); backward only scales it
        want = any(ctx.needs_input_grad[2:])
        g_disp = [torch.empty_like(d) for d in disps] if want else None
        g_scratch = [torch.empty_like(d) for d in disps] if (want and normalize) else None
        EdgeSmoothFn._launch(tgt, disps, normalize, want, g_disp, g_scratch, loss, None)
        ctx.g_disp = g_disp
        return loss

    @staticmethod
    def backward(ctx, g):
        g_disp, ctx.g_disp = ctx.g_disp, None
        g = g.detach().to(torch.float32)
        need = ctx.needs_input_grad[2:]
        scaled = torch._foreach_mul(list(g_disp), g)           # one multi-tensor launch for the whole pyramid
        return (None, None) + tuple((gd if n else None) for gd, n in zip(scaled, need))


def edge_aware_smooth(disps, tgt, normalize=True):
    """Edge-aware smoothness over a disparity pyramid ([B,1,H/2^s,W/2^s]) against the target image."""
    return EdgeSmoothFn.apply(tgt, bool(normalize), *disps)


# ---------------------------------------------------------------------------
# pseudo-LiDAR
# ---------------------------------------------------------------------------
def cloud_project(depth, P, Tinv, sparsity=0, want_f64=True, want_f32=False, want_index=False, want_valid=False):
    """depth [B,H,W] f32 CUDA -> dict(count[B] int32, cloud_f64 [B,HW,4], ...).  No sync."""
    _need_cuda(depth)
    depth = _f32c(depth)
    B, H, W = depth.shape
    dev = depth.device
    a = _lib.CloudArgs()
    a.B, a.H, a.W, a.sparsity = B, H, W, int(sparsity or 0)
    a.depth = depth.data_ptr()
    for i, v in enumerate([float(x) for x in P.reshape(-1)]):
        a.P[i] = v
    for i, v in enumerate([float(x) for x in Tinv.reshape(-1)]):
        a.Tinv[i] = v
    res = {"count": torch.empty(B, dtype=torch.int32, device=dev)}
    a.count = res["count"].data_ptr()
    if want_f64:
        res["cloud_f64"] = torch.empty(B, H * W, 4, dtype=torch.float64, device=dev)
        a.cloud_f64 = res["cloud_f64"].data_ptr()
    if want_f32:
        res["cloud_f32"] = torch.empty(B, H * W, 4, dtype=torch.float32, device=dev)
        a.cloud_f32 = res["cloud_f32"].data_ptr()
    if want_index:
        res["index"] = torch.empty(B, H * W, dtype=torch.int32, device=dev)
        a.index = res["index"].data_ptr()
    if want_valid:
        res["valid"] = torch.empty(B, H * W, dtype=torch.uint8, device=dev)
        a.valid = res["valid"].data_ptr()
    ws = _workspace("cloud", lib.plb_cloud_workspace_bytes(a), dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    check(lib.plb_cloud_project(a, _stream()), "plb_cloud_project")
    return res


def velo_project(points, T, P, height, width, counts=None, want_f64=True, want_f32=False, want_winner=False):
    """points [B,N,C>=3] f32 CUDA (KITTI .bin rows: x,y,z,reflectance) -> dict(depth_f64 [B,H,W], ...): the
    sparse depth image of `Transform.project_velo_to_img` (pseudo-lidar/Transform/Transform.py:69-104).  No sync."""
    _need_cuda(points)
    points = _f32c(points)
    B, N, Cs = points.shape
    dev = points.device
    a = _lib.VeloArgs()
    a.B, a.N, a.H, a.W, a.point_stride = B, N, int(height), int(width), Cs
    a.points = points.data_ptr() if N > 0 else 0
    if counts is not None:
        counts = counts.to(dev, torch.int32).contiguous()
        a.counts = counts.data_ptr()
    for i, v in enumerate([float(x) for x in T.reshape(-1)]):
        a.T[i] = v
    for i, v in enumerate([float(x) for x in P.reshape(-1)]):
        a.P[i] = v
    res = {}
    if want_f64:
        res["depth_f64"] = torch.empty(B, a.H, a.W, dtype=torch.float64, device=dev)
        a.depth_f64 = res["depth_f64"].data_ptr()
    if want_f32:
        res["depth_f32"] = torch.empty(B, a.H, a.W, dtype=torch.float32, device=dev)
        a.depth_f32 = res["depth_f32"].data_ptr()
    if want_winner:
        res["winner"] = torch.empty(B, a.H, a.W, dtype=torch.int32, device=dev)
        a.winner = res["winner"].data_ptr()
    ws = _workspace("velo", lib.plb_velo_workspace_bytes(a), dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    check(lib.plb_velo_project(a, _stream()), "plb_velo_project")
    return res


def smooth_only(depth_maps, input_is_depth=True, fused_backward=True):
    """Losses.smooth_loss on a list of [B,1,h,w] maps (no photometric launch)."""
    d0 = depth_maps[0]
    cfg = LossConfig(0, [len(depth_maps)], input_is_depth=input_is_depth, do_photo=False, do_smooth=True,
                     fused_backward=fused_backward)
    poses = torch.zeros(d0.shape[0], 1, 6, dtype=torch.float32, device=d0.device)
    K = torch.zeros(d0.shape[0], 3, 3, dtype=torch.float32, device=d0.device)
    # with do_photo=False the first tensor only supplies the batch size and device
    _, smooth = FusedLossFn.apply(cfg, d0.detach(), poses, K, *depth_maps)
    return smooth


# ---------------------------------------------------------------------------
# loader-side frame preparation
# ---------------------------------------------------------------------------
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def prep_frames(frames, height, width, K=None, mean=IMAGENET_MEAN, std=IMAGENET_STD, want_planar=True, want_nhwc4=False,
                out=None):
    """frames [B,h,w,3] uint8 CUDA (decoded RGB, HWC), K [n,3,3] (n <= B: one matrix per frame or per training sample)
    -> dict(planar [B,3,H,W] f32, nhwc4 [B,H,W,4] f32, K [n,3,3] f64):
    the reference loader's transform chain and intrinsics scaling (trainer.py:97-103, dataloaders.py:32-49,95-98),
    bit-exact.  `out`: an optional preallocated [B,3,H,W] float32 tensor for the planar result.  No sync."""
    _need_cuda(frames, K)
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise ValueError("frames must be uint8 [B, h, w, 3]")
    frames = frames.contiguous()
    B, h, w, _ = frames.shape
    dev = frames.device
    a = _lib.PrepArgs()
    a.B, a.in_h, a.in_w, a.H, a.W = B, h, w, int(height), int(width)
    a.frames = frames.data_ptr()
    for c in range(3):
        a.mean[c], a.stdev[c] = float(mean[c]), float(std[c])
    res = {}
    if want_planar:
        if out is not None:
            if out.shape != (B, 3, a.H, a.W) or out.dtype != torch.float32 or not out.is_contiguous():
                raise ValueError("out must be a contiguous float32 [B,3,H,W] tensor")
            res["planar"] = out
        else:
            res["planar"] = torch.empty(B, 3, a.H, a.W, dtype=torch.float32, device=dev)
        a.out_planar = res["planar"].data_ptr()
    if want_nhwc4:
        res["nhwc4"] = torch.empty(B, a.H, a.W, 4, dtype=torch.float32, device=dev)
        a.out_nhwc4 = res["nhwc4"].data_ptr()
    if K is not None:
        # one matrix per frame, or one per training sample (its 1 + n_src frames share it): [n,3,3] with n <= B
        if K.dim() != 3 or tuple(K.shape[1:]) != (3, 3) or K.shape[0] < 1 or K.shape[0] > B:
            raise ValueError("K must be [n,3,3] with 1 <= n <= %d frames, got %s" % (B, tuple(K.shape)))
        K = K.to(torch.float64).contiguous()
        res["K"] = torch.empty_like(K)
        a.K_in, a.K_out, a.n_K = K.data_ptr(), res["K"].data_ptr(), K.shape[0]
    ws = _workspace("prep", lib.plb_prep_workspace_bytes(a), dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    check(lib.plb_prep_frames(a, _stream()), "plb_prep_frames")
    return res
