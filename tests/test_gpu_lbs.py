"""SMPL LBS: CUDA path (through the C-ABI) vs the CPU oracle.  Tolerance: vertex / joint L2 <= 1e-4 m
(BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import smpl as osmpl
from oracle import so3
from util import smpl_data

pytestmark = pytest.mark.gpu
TOL_M = 1e-4


def _smpl(batch_size=1, create_transl=True):
    import humaniflow_b200 as hb
    return hb.SMPL.from_arrays(smpl_data(), batch_size=batch_size, create_transl=create_transl).cuda()


def _l2(a, b):
    return (a.cpu() - b).norm(dim=-1).max().item()


def _inputs(M, seed=0, pose_std=0.3):
    g = torch.Generator().manual_seed(seed)
    betas = torch.randn(M, 10, generator=g)
    theta = torch.randn(M, 24, 3, generator=g) * pose_std
    return betas, theta


def test_config1_axis_angle_and_rotmats():
    """BASELINE.json configs[0]: batch=4, random betas(10) / theta(24x3) -> 6890 verts, 90 joints."""
    smpl = _smpl()
    data = smpl_data()
    betas, theta = _inputs(4)
    v_ref, j_ref = osmpl.smpl_forward(data, betas, theta[:, 1:].reshape(4, 69), theta[:, 0], pose2rot=True,
                                      transl=torch.zeros(1, 3))
    out = smpl(betas=betas.cuda(), body_pose=theta[:, 1:].reshape(4, 69).cuda(), global_orient=theta[:, 0].cuda(), pose2rot=True)
    assert out.vertices.shape == (4, 6890, 3) and out.joints.shape == (4, 90, 3) and out.vertices.dtype == torch.float32
    assert _l2(out.vertices, v_ref) <= TOL_M and _l2(out.joints, j_ref) <= TOL_M
    R = so3.batch_rodrigues(theta.reshape(-1, 3)).view(4, 24, 3, 3)
    v2, j2 = osmpl.smpl_forward(data, betas, R[:, 1:], R[:, :1], pose2rot=False)
    out2 = smpl(betas=betas.cuda(), body_pose=R[:, 1:].cuda(), global_orient=R[:, :1].cuda(), pose2rot=False)
    assert _l2(out2.vertices, v2) <= TOL_M and _l2(out2.joints, j2) <= TOL_M


@pytest.mark.parametrize('M', [1, 63, 65, 130])
def test_ragged_batches_and_transl(M):
    smpl = _smpl()
    data = smpl_data()
    betas, theta = _inputs(M, seed=M, pose_std=0.8)
    R = so3.batch_rodrigues(theta.reshape(-1, 3)).view(M, 24, 3, 3)
    transl = torch.randn(M, 3, generator=torch.Generator().manual_seed(3))
    v_ref, j_ref = osmpl.smpl_forward(data, betas, R[:, 1:], R[:, :1], pose2rot=False, transl=transl)
    out = smpl(betas=betas.cuda(), body_pose=R[:, 1:].cuda(), global_orient=R[:, :1].cuda(), transl=transl.cuda(), pose2rot=False)
    assert _l2(out.vertices, v_ref) <= TOL_M and _l2(out.joints, j_ref) <= TOL_M


def test_module_defaults_tpose_and_beta_broadcast():
    """None inputs fall back to the module's zero parameters (predict_humaniflow.py:147); betas with batch 1
    broadcast over the pose batch ([upstream] SMPL.forward)."""
    smpl = _smpl(batch_size=3)
    data = smpl_data()
    out = smpl()
    v_ref, j_ref = osmpl.smpl_forward(data, torch.zeros(3, 10), torch.zeros(3, 69), torch.zeros(3, 3), pose2rot=True,
                                      transl=torch.zeros(3, 3))
    assert _l2(out.vertices, v_ref) <= TOL_M and _l2(out.joints, j_ref) <= TOL_M
    smpl = _smpl(batch_size=1)
    betas, theta = _inputs(5, seed=9)
    out = smpl(betas=betas[:1].cuda(), body_pose=theta[:, 1:].reshape(5, 69).cuda(), global_orient=theta[:, 0].cuda())
    v_ref, _ = osmpl.smpl_forward(data, betas[:1], theta[:, 1:].reshape(5, 69), theta[:, 0], transl=torch.zeros(1, 3))
    assert _l2(out.vertices, v_ref) <= TOL_M


def test_full_size_properties():
    """B*N = 3200 (BASELINE configs[1..3]): oracle-free properties + oracle spot rows.
    identity pose => vertices == v_template + shapedirs.beta; first 24 joints == chain translations;
    a global rotation about the root joint rotates the whole mesh rigidly."""
    M = 3200
    smpl = _smpl(create_transl=False)
    data = smpl_data()
    betas, theta = _inputs(M, seed=11, pose_std=0.5)
    eye = torch.eye(3).expand(M, 24, 3, 3).contiguous()
    out = smpl(betas=betas.cuda(), body_pose=eye[:, 1:].cuda(), global_orient=eye[:, :1].cuda(), pose2rot=False)
    v_shaped = data['v_template'][None] + torch.einsum('bl,mkl->bmk', betas, data['shapedirs'])
    assert _l2(out.vertices, v_shaped) <= 1e-5      # shape terms: three-product fp16 split, ~1e-7 m
    J = torch.einsum('bik,ji->bjk', v_shaped, data['J_regressor'])
    assert _l2(out.joints[:, :24], J) <= 2e-6
    R = so3.batch_rodrigues(theta.reshape(-1, 3)).view(M, 24, 3, 3)
    outp = smpl(betas=betas.cuda(), body_pose=R[:, 1:].cuda(), global_orient=R[:, :1].cuda(), pose2rot=False)
    rows = [0, 1, 777, 1599, 3135, 3199]
    v_ref, j_ref = osmpl.smpl_forward(data, betas[rows], R[rows][:, 1:], R[rows][:, :1], pose2rot=False)
    assert _l2(outp.vertices[rows], v_ref) <= TOL_M and _l2(outp.joints[rows], j_ref) <= TOL_M
    # rigidity: same body pose, identity global orientation, then rotate about the root joint
    out0 = smpl(betas=betas.cuda(), body_pose=R[:, 1:].cuda(), global_orient=eye[:, :1].cuda(), pose2rot=False)
    root = out0.joints[:, :1].cpu()
    rotated = torch.einsum('bij,bvj->bvi', R[:, 0], out0.vertices.cpu() - root) + root
    assert (outp.vertices.cpu() - rotated).norm(dim=-1).max().item() <= 1e-5
    assert torch.isfinite(outp.vertices).all()


def test_joint_layout_90():
    """joints = 24 chain + 21 vertex picks + 9 + 19 + 17 regressed (models/smpl.py:30-34, label_conversions.py:17-21)."""
    smpl = _smpl()
    betas, theta = _inputs(2, seed=4)
    out = smpl(betas=betas.cuda(), body_pose=theta[:, 1:].reshape(2, 69).cuda(), global_orient=theta[:, 0].cuda())
    v = out.vertices.cpu()
    from humaniflow_b200.smpl import VERTEX_JOINT_IDS
    assert torch.allclose(out.joints[:, 24:45].cpu(), v[:, VERTEX_JOINT_IDS], atol=0, rtol=0)
    data = smpl_data()
    h36m = torch.einsum('bik,ji->bjk', v, data['J_regressor_h36m'])
    assert (out.joints[:, 73:90].cpu() - h36m).norm(dim=-1).max().item() <= 1e-5


def test_no_cpu_fallback():
    import humaniflow_b200 as hb
    smpl = hb.SMPL.from_arrays(smpl_data())
    with pytest.raises(RuntimeError):
        smpl(betas=torch.zeros(1, 10), body_pose=torch.zeros(1, 69), global_orient=torch.zeros(1, 3))


def test_blend_implementations_cross_check():
    """The product path (persistent fp16 tcgen05 blend: one fp16 product per pose term, three-product split for the
    shape terms), the FP32 CUDA-core blend (impl 1) and the split-bf16 three-pass tcgen05 blend (impl 2) agree with
    each other and with the oracle."""
    smpl = _smpl(create_transl=False)
    data = smpl_data()
    M = 200
    betas, theta = _inputs(M, seed=21, pose_std=0.7)
    R = so3.batch_rodrigues(theta.reshape(-1, 3)).view(M, 24, 3, 3)
    v_ref, j_ref = osmpl.smpl_forward(data, betas, R[:, 1:], R[:, :1], pose2rot=False)
    args = dict(betas=betas.cuda(), body_pose=R[:, 1:].cuda(), global_orient=R[:, :1].cuda(), pose2rot=False)
    tc = smpl(**args)
    smpl.set_impl(1)
    cc = smpl(**args)
    smpl.set_impl(2)
    t3 = smpl(**args)
    smpl.set_impl(0)
    print('fp16-tc err %.2e  fp32-cc err %.2e  bf16x3-tc err %.2e' % (_l2(tc.vertices, v_ref), _l2(cc.vertices, v_ref), _l2(t3.vertices, v_ref)))
    assert _l2(tc.vertices, v_ref) <= TOL_M and _l2(cc.vertices, v_ref) <= TOL_M and _l2(t3.vertices, v_ref) <= TOL_M
    assert _l2(tc.joints, j_ref) <= TOL_M
    assert (t3.vertices - cc.vertices).norm(dim=-1).max().item() <= 2e-5
    assert (tc.vertices - cc.vertices).norm(dim=-1).max().item() <= TOL_M


@pytest.mark.parametrize('M', [64, 192, 1000])
def test_sample_tile_switches(M):
    """The persistent kernel replaces its resident sample tile (coefficients + transforms) mid-range: cover batch sizes
    where CTAs own whole tiles, fractions of a tile and ranges spanning two tiles."""
    smpl = _smpl(create_transl=False)
    data = smpl_data()
    betas, theta = _inputs(M, seed=100 + M, pose_std=0.4)
    R = so3.batch_rodrigues(theta.reshape(-1, 3)).view(M, 24, 3, 3)
    out = smpl(betas=betas.cuda(), body_pose=R[:, 1:].cuda(), global_orient=R[:, :1].cuda(), pose2rot=False)
    rows = sorted(set([0, 1, 63, M // 2, M - 65 if M > 65 else 0, M - 2, M - 1]))
    v_ref, j_ref = osmpl.smpl_forward(data, betas[rows], R[rows][:, 1:], R[rows][:, :1], pose2rot=False)
    assert _l2(out.vertices[rows], v_ref) <= TOL_M and _l2(out.joints[rows], j_ref) <= TOL_M
    smpl.set_impl(1)
    cc = smpl(betas=betas.cuda(), body_pose=R[:, 1:].cuda(), global_orient=R[:, :1].cuda(), pose2rot=False)
    smpl.set_impl(0)
    assert (out.vertices - cc.vertices).norm(dim=-1).max().item() <= TOL_M      # every row, against the FP32 CUDA-core path


@pytest.mark.parametrize('M,with_transl', [(37, False), (200, True), (1000, False)])
def test_lanes_as_samples_kernel_cross_check(M, with_transl):
    """impl 3 (round-2 experiment: TMEM lanes = samples, joint transforms in registers along per-tile sorted vertices, flagged
    vertices through a sample-contiguous side buffer) against the oracle and, row by row, against the FP32 CUDA-core path;
    covers ragged sample tiles, sample-tile switches and the translation."""
    smpl = _smpl(create_transl=False)
    data = smpl_data()
    betas, theta = _inputs(M, seed=300 + M, pose_std=0.6)
    R = so3.batch_rodrigues(theta.reshape(-1, 3)).view(M, 24, 3, 3)
    transl = torch.randn(M, 3, generator=torch.Generator().manual_seed(M)) * 0.3 if with_transl else None
    args = dict(betas=betas.cuda(), body_pose=R[:, 1:].cuda(), global_orient=R[:, :1].cuda(), pose2rot=False,
                transl=None if transl is None else transl.cuda())
    smpl.set_impl(3)
    t3 = smpl(**args)
    smpl.set_impl(1)
    cc = smpl(**args)
    smpl.set_impl(0)
    rows = sorted(set([0, 1, M // 2, M - 2, M - 1]))
    v_ref, j_ref = osmpl.smpl_forward(data, betas[rows], R[rows][:, 1:], R[rows][:, :1], pose2rot=False,
                                      transl=None if transl is None else transl[rows])
    assert _l2(t3.vertices[rows], v_ref) <= TOL_M and _l2(t3.joints[rows], j_ref) <= TOL_M
    assert (t3.vertices - cc.vertices).norm(dim=-1).max().item() <= TOL_M
    assert (t3.joints - cc.joints).norm(dim=-1).max().item() <= TOL_M


@pytest.mark.parametrize('impl', [0, 1, 2, 3])
def test_forward_samples_split_rotations(impl):
    """hf_lbs_forward_split (body rotations + one global rotation per image shared by its N samples) gives bit-identical
    results to hf_lbs_forward on the expanded / concatenated (M,24,3,3) tensor, for every blend implementation."""
    smpl = _smpl(create_transl=False)
    B, N = 5, 7
    M = B * N
    betas, theta = _inputs(M, seed=400, pose_std=0.5)
    R = so3.batch_rodrigues(theta.reshape(-1, 3)).view(M, 24, 3, 3)
    glob = R[::N, 0].contiguous()                                         # (B,3,3)
    body = R[:, 1:].contiguous()
    full = torch.cat([glob[:, None].expand(-1, N, -1, -1).reshape(M, 1, 3, 3), body], 1).contiguous()
    smpl.set_impl(impl)
    a = smpl.lbs(betas.cuda(), full.cuda())
    b = smpl.forward_samples(betas.cuda(), body.cuda(), glob.cuda(), N)
    smpl.set_impl(0)
    assert torch.equal(a[0], b.vertices) and torch.equal(a[1], b.joints)


def _loss_weights(M, V, JO, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(M, V, 3, generator=g, dtype=torch.float64), torch.randn(M, JO, 3, generator=g, dtype=torch.float64)


@pytest.mark.parametrize('M,use_verts,use_joints', [(5, True, True), (3, False, True), (20, True, False), (130, True, True)])
def test_backward_against_autograd_through_the_oracle(M, use_verts, use_joints):
    """hf_lbs_backward (three launches) against torch.autograd through the float64 oracle LBS, for a random
    linear functional of the vertices and / or the 90 joints.  Tolerance: 2e-4 of the largest gradient entry per tensor (fp32
    forward recompute + split-tf32 GEMM + atomics vs float64)."""
    smpl = _smpl()
    data = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in smpl_data().items()}
    betas, theta = _inputs(M, seed=40 + M, pose_std=0.6)
    R = so3.batch_rodrigues(theta.reshape(-1, 3)).view(M, 24, 3, 3)
    transl = torch.randn(M, 3, generator=torch.Generator().manual_seed(5))
    wv, wj = _loss_weights(M, 6890, 90, seed=M)

    b64, R64, t64 = betas.double().requires_grad_(), R.double().requires_grad_(), transl.double().requires_grad_()
    v_ref, j_ref = osmpl.smpl_forward(data, b64, R64[:, 1:], R64[:, :1], pose2rot=False, transl=t64)
    loss_ref = (v_ref * wv).sum() * float(use_verts) + (j_ref * wj).sum() * float(use_joints)
    loss_ref.backward()

    bc, Rc, tc = betas.cuda().requires_grad_(), R.cuda().requires_grad_(), transl.cuda().requires_grad_()
    out = smpl(betas=bc, body_pose=Rc[:, 1:], global_orient=Rc[:, :1], transl=tc, pose2rot=False)
    loss = 0.
    if use_verts:
        loss = loss + (out.vertices * wv.float().cuda()).sum()
    if use_joints:
        loss = loss + (out.joints * wj.float().cuda()).sum()
    loss.backward()
    for name, got, ref in (('betas', bc.grad, b64.grad), ('rotmats', Rc.grad, R64.grad), ('transl', tc.grad, t64.grad)):
        err = (got.double().cpu() - ref).abs().max().item()
        scale = ref.abs().max().item()
        assert err <= 2e-4 * scale, (name, err, scale)


def test_fitting_step_through_axis_angle_pose():
    """The fitting-loop pattern: axis-angle pose and betas as leaf tensors, a joints loss, gradient descent through the CUDA
    forward / backward; the gradients match autograd through the oracle and the loss goes down."""
    smpl = _smpl(create_transl=False)
    data = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in smpl_data().items()}
    M = 6
    betas, theta = _inputs(M, seed=77, pose_std=0.4)
    tb, tt = _inputs(M, seed=78, pose_std=0.4)
    with torch.no_grad():
        target = smpl(betas=tb.cuda(), body_pose=tt[:, 1:].reshape(M, 69).cuda(), global_orient=tt[:, 0].cuda()).joints
    b64, th64 = betas.double().requires_grad_(), theta.double().requires_grad_()
    _, j_ref = osmpl.smpl_forward(data, b64, th64[:, 1:].reshape(M, 69), th64[:, 0], pose2rot=True)
    ((j_ref - target.double().cpu()) ** 2).sum().backward()

    bc, thc = betas.cuda().requires_grad_(), theta.cuda().requires_grad_()
    losses = []
    for it in range(5):
        out = smpl(betas=bc, body_pose=thc[:, 1:].reshape(M, 69), global_orient=thc[:, 0], pose2rot=True)
        loss = ((out.joints - target) ** 2).sum()
        bc.grad = thc.grad = None
        loss.backward()
        if it == 0:
            for got, ref in ((bc.grad, b64.grad), (thc.grad, th64.grad)):
                assert (got.double().cpu() - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()
        losses.append(loss.item())
        with torch.no_grad():
            bc -= 0.02 * bc.grad
            thc -= 0.02 * thc.grad
    assert losses[-1] < 0.7 * losses[0], losses


def test_argument_errors_are_reported_not_crashed():
    """Bad calls through the C-ABI come back as error codes with a message (hf_last_error), never as a crashed launch: both
    incoming gradients NULL, workspaces that are too small."""
    from humaniflow_b200 import _lib
    lib = _lib.load()
    smpl = _smpl()
    dev = torch.device('cuda')
    h = smpl._handle(dev)
    M = 3
    betas = torch.zeros(M, 10, device=dev)
    R = torch.eye(3, device=dev).expand(M, 24, 3, 3).contiguous()
    gb, gr = torch.empty_like(betas), torch.empty_like(R)
    ws = torch.empty(lib.hf_lbs_backward_workspace_bytes(h, M), device=dev, dtype=torch.uint8)
    rc = lib.hf_lbs_backward(h, _lib.ptr(betas), _lib.ptr(R), None, None, _lib.ptr(gb), _lib.ptr(gr), _lib.ptr(ws), ws.numel(), M, _lib.stream())
    assert rc != 0 and b'NULL' in lib.hf_last_error()
    gj = torch.zeros(M, 90, 3, device=dev)
    rc = lib.hf_lbs_backward(h, _lib.ptr(betas), _lib.ptr(R), None, _lib.ptr(gj), _lib.ptr(gb), _lib.ptr(gr), _lib.ptr(ws), 16, M, _lib.stream())
    assert rc != 0 and b'workspace' in lib.hf_last_error()
    x = torch.zeros(32, 2048, device=dev)
    W = torch.zeros(1024, 2048, device=dev)
    y = torch.zeros(32, 1024, device=dev)
    small = torch.empty(64, device=dev, dtype=torch.uint8)
    assert lib.hf_linear_workspace_bytes(32, 2048, 1024) > 64
    rc = lib.hf_linear_ws(_lib.ptr(x), 2048, _lib.ptr(W), 2048, None, _lib.ptr(y), 1024, 32, 2048, 1024, 0, 0, _lib.ptr(small), 64, _lib.stream())
    assert rc != 0 and b'workspace' in lib.hf_last_error()
    pred = torch.zeros(8, 100, 20, 3, device=dev)
    tgt = torch.zeros(8, 20, 3, device=dev)
    out = torch.empty(8, 100, 3, device=dev)
    rc = lib.hf_pointset_errors_ws(_lib.ptr(pred), _lib.ptr(tgt), 8, 100, 20, _lib.ptr(out), _lib.ptr(small), 64, _lib.stream())
    assert rc != 0 and b'workspace' in lib.hf_last_error()
    # zero gradients in: zero gradients out (and a finite result for identity poses)
    rc = lib.hf_lbs_backward(h, _lib.ptr(betas), _lib.ptr(R), None, _lib.ptr(gj), _lib.ptr(gb), _lib.ptr(gr), _lib.ptr(ws), ws.numel(), M, _lib.stream())
    torch.cuda.synchronize()
    assert rc == 0 and gb.abs().max().item() == 0.0 and gr.abs().max().item() == 0.0
