"""GPU parity of the stand-alone geometry drop-ins against golden vectors."""
import pytest
import torch

from helpers import load_golden, rel_err, max_rel_err

pytestmark = pytest.mark.gpu


def test_inverse_warp_forward_backward_golden():
    from geometry.pose_geometry import inverse_warp
    g = load_golden("warp_b4_24x40")
    dev = torch.device("cuda:0")
    K = torch.from_numpy(g["K"]).to(dev)
    cot = torch.from_numpy(g["cot"]).to(dev)
    for inv, tag in ((False, "fwd"), (True, "inv")):
        img = torch.from_numpy(g["img"]).to(dev).requires_grad_(True)
        depth = torch.from_numpy(g["depth"]).to(dev).requires_grad_(True)
        pose = torch.from_numpy(g["pose"]).to(dev).requires_grad_(True)
        proj = inverse_warp(img, depth, pose, K, inv)
        (proj * cot).sum().backward()
        assert max_rel_err(proj.cpu(), g["proj_" + tag]) < 1e-5
        assert rel_err(img.grad.cpu(), g["g_img_" + tag]) < 1e-5
        assert rel_err(depth.grad.cpu(), g["g_depth_" + tag]) < 1e-4
        assert rel_err(pose.grad.cpu(), g["g_pose_" + tag]) < 1e-4


def test_pose_functions_and_transform_golden():
    from geometry.pose_geometry import (transformation_from_parameters, invert_pose, pose_vec2mat, euler2mat,
                                        disp_to_depth)
    from geometry.transform import Transform
    g = load_golden("warp_b4_24x40")
    dev = torch.device("cuda:0")
    pose = torch.from_numpy(g["pose"]).to(dev)
    M = transformation_from_parameters(pose[:, :3].unsqueeze(1), pose[:, 3:].unsqueeze(1))
    assert torch.allclose(M.cpu(), torch.from_numpy(g["M_axisangle"]), atol=2e-7)
    assert torch.allclose(invert_pose(M).cpu(), torch.from_numpy(g["M_inverted"]), atol=2e-7)
    assert torch.allclose(pose_vec2mat(pose, "euler").cpu(), torch.from_numpy(g["M_euler"]), atol=2e-7)
    assert euler2mat(pose[:, :3]).shape == (4, 3, 3)
    K = torch.from_numpy(g["K"]).to(dev)
    depth = torch.from_numpy(g["depth"]).to(dev)
    t = Transform()
    Xc = t.reconstruct(depth[:, 0], K)
    assert max_rel_err(Xc.cpu(), g["Xc"]) < 1e-6
    assert max_rel_err(t.project(Xc, K, M).cpu(), g["grid"]) < 1e-5
    d = torch.rand(2, 1, 8, 8, device=dev)
    assert torch.allclose(disp_to_depth([[d]])[0][0], 1 / (10 * d + 0.01), rtol=1e-6)


@pytest.mark.parametrize("mode", ["axisangle", "euler"])
@pytest.mark.parametrize("invert", [False, True])
def test_pose_matrix_vjp_against_oracle(mode, invert):
    from plb200 import ops, _lib
    from oracle import restated as O
    torch.manual_seed(0)
    pose = torch.cat([0.3 * torch.randn(6, 3), torch.randn(6, 3)], 1)
    pose[0, :3] = 0  # zero rotation: axis = 0/(0+1e-7)
    cot = torch.randn(6, 4, 4)
    p = pose.clone().requires_grad_(True)
    M = O.pose_matrix(p, invert, mode)
    (M * cot).sum().backward()
    q = pose.cuda().requires_grad_(True)
    Mg = ops.PoseMatrixFn.apply(q, _lib.ROT_EULER if mode == "euler" else _lib.ROT_AXISANGLE, invert)
    (Mg * cot.cuda()).sum().backward()
    assert torch.allclose(Mg.cpu(), M.detach(), atol=1e-6)
    assert rel_err(q.grad.cpu(), p.grad) < 1e-5
