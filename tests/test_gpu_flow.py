"""Ancestor-conditioned SO(3) flow: CUDA path (C-ABI via HumaniflowModel) vs the CPU oracle on identical
weights, features and injected noise.  Tolerances (written here, from BASELINE.json north_star):
rotation matrices <= 2e-5 abs (fp32 noise floor of the 23-joint chain is ~4e-6), log_prob rel err <= 1e-3."""
import math

import pytest
import torch

from oracle import model as om
from util import make_model, special_rotations
from humaniflow_b200.synthetic import SMPL_PARENTS

pytestmark = pytest.mark.gpu
ROT_TOL = 2e-5
LP_RTOL = 1e-3


def _noise(B, N, seed=1, positive_feats=True, feat_dim=512):
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, feat_dim, generator=g)
    if positive_feats:
        feats = feats.abs()            # post-ReLU avg-pooled features are non-negative
    z = torch.randn(B, N, 23, 3, generator=g) * 0.6
    se = torch.randn(B, N, 10, generator=g)
    return feats, z, se


def _run(m, feats, z, se, **kw):
    N = z.shape[1] if z is not None else 0
    return m(None, input_feats=feats.cuda(), num_samples=N, base_noise=None if z is None else z.cuda(),
             shape_eps=None if se is None else se.cuda(), **kw)


@pytest.mark.parametrize('B,N,scale,layers', [(3, 7, 1.0, 18), (4, 25, 1.5, 18), (32, 100, 1.0, 18), (5, 9, 1.5, 50), (32, 100, 1.0, 50)])
def test_sampling_and_point_estimate_parity(B, N, scale, layers):
    """(32,100) is BASELINE configs[1]: B=32, N=100, 23 joints, random-init flow + synthetic encoder features; layers=50 is
    the benchmarked width (feats (32,2048), 2048->1024 fc1, 1024-wide heads, 2070-wide image-level Linear)."""
    m, sd, cfg = make_model(layers, seed=0, flow_scale=scale)
    m = m.cuda()
    feats, z, se = _noise(B, N, feat_dim=m.input_feats_dim)
    ref = om.forward(sd, cfg, SMPL_PARENTS, input_feats=feats, num_samples=N, shape_eps=se, base_noise=z)
    out = _run(m, feats, z, se)
    assert out['pose_rotmats_samples'].shape == (B, N, 23, 3, 3) and out['pose_rotmats_samples'].dtype == torch.float32
    assert out['shape_samples'].shape == (B, N, 10)
    assert out['pose_rotmats_point_est'].shape == (B, 23, 3, 3) and out['pose_axisangle_point_est'].shape == (B, 23, 3)
    for k in ('cam_wp', 'glob_rotmat', 'shape_mode', 'shape_log_std', 'shape_samples'):
        assert torch.allclose(out[k].cpu(), ref[k], atol=2e-5, rtol=1e-5), k
    tol = ROT_TOL * (1 if scale == 1.0 else 4)
    assert (out['pose_rotmats_samples'].cpu() - ref['pose_rotmats_samples']).abs().max().item() <= tol
    assert (out['pose_rotmats_point_est'].cpu() - ref['pose_rotmats_point_est']).abs().max().item() <= tol
    assert (out['pose_axisangle_point_est'].cpu() - ref['pose_axisangle_point_est']).abs().max().item() <= tol
    # size-independent properties: every sample is a rotation
    R = out['pose_rotmats_samples'].double()
    assert (R @ R.transpose(-1, -2) - torch.eye(3, device=R.device, dtype=R.dtype)).abs().max().item() <= 1e-6
    assert (torch.linalg.det(R) - 1).abs().max().item() <= 1e-6


def test_tensor_core_prologue_against_the_cuda_core_prologue(monkeypatch):
    """The image-feature part of every context Linear runs as split-tf32 on tcgen05 (flow_ctx_gemm_kernel); the fp32 CUDA-core
    prologue inside the sampling kernel is kept as a cross-check (HF_FLOW_SIMT_PROLOGUE=1).  Both must agree far inside ROT_TOL,
    and both must meet ROT_TOL against the oracle, at a row count that is not a multiple of the 128-row tile."""
    m, sd, cfg = make_model(50, seed=4)
    m = m.cuda()
    B, N = 7, 37
    feats, z, se = _noise(B, N, seed=9, feat_dim=m.input_feats_dim)
    ref = om.forward(sd, cfg, SMPL_PARENTS, input_feats=feats, num_samples=N, shape_eps=se, base_noise=z)['pose_rotmats_samples']
    tc = _run(m, feats, z, se)['pose_rotmats_samples'].cpu().clone()
    monkeypatch.setenv('HF_FLOW_SIMT_PROLOGUE', '1')
    simt = _run(m, feats, z, se)['pose_rotmats_samples'].cpu().clone()
    monkeypatch.delenv('HF_FLOW_SIMT_PROLOGUE')
    e_tc, e_simt, d = (tc - ref).abs().max().item(), (simt - ref).abs().max().item(), (tc - simt).abs().max().item()
    print('prologue: tensor-core vs oracle %.2e, CUDA-core vs oracle %.2e, between them %.2e' % (e_tc, e_simt, d))
    assert e_tc <= ROT_TOL and e_simt <= ROT_TOL and d <= ROT_TOL


@pytest.mark.parametrize('B,N', [(32, 100), (3, 5), (9, 40)])
def test_level_parallel_kernel_equals_the_joint_by_joint_kernel(B, N, monkeypatch):
    """The product sampler walks the kinematic tree level by level (three thread groups, one joint each per round, weights
    streamed layer by layer); the joint-by-joint kernel (HF_FLOW_CHAIN=1) does the same arithmetic in the same order, so the two
    agree bit for bit -- at all three rows-per-CTA variants (8 / 16 / 24)."""
    m, sd, cfg = make_model(50, seed=5)
    m = m.cuda()
    feats, z, se = _noise(B, N, seed=B + N, feat_dim=m.input_feats_dim)
    lev = _run(m, feats, z, se)
    lev = {k: lev[k].clone() for k in ('pose_rotmats_samples', 'pose_rotmats_point_est', 'pose_axisangle_point_est')}
    monkeypatch.setenv('HF_FLOW_CHAIN', '1')
    chain = _run(m, feats, z, se)
    monkeypatch.delenv('HF_FLOW_CHAIN')
    for k in lev:
        assert torch.equal(lev[k], chain[k]), k


def test_modes_of_forward():
    m, sd, cfg = make_model(18, seed=2)
    m = m.cuda()
    feats, z, se = _noise(5, 4, seed=3)
    # samples only
    o = _run(m, feats, z, se, compute_point_est=False)
    assert 'pose_rotmats_point_est' not in o and o['pose_rotmats_samples'].shape == (5, 4, 23, 3, 3)
    ref = om.forward(sd, cfg, SMPL_PARENTS, input_feats=feats, compute_point_est=False, num_samples=4, shape_eps=se, base_noise=z)
    assert (o['pose_rotmats_samples'].cpu() - ref['pose_rotmats_samples']).abs().max().item() <= ROT_TOL
    # point estimate only
    o = m(None, input_feats=feats.cuda())
    assert 'pose_rotmats_samples' not in o
    ref = om.forward(sd, cfg, SMPL_PARENTS, input_feats=feats)
    assert (o['pose_rotmats_point_est'].cpu() - ref['pose_rotmats_point_est']).abs().max().item() <= ROT_TOL
    # shape mode for samples
    o = _run(m, feats, z, None, use_shape_mode_for_samples=True)
    ref = om.forward(sd, cfg, SMPL_PARENTS, input_feats=feats, num_samples=4, use_shape_mode_for_samples=True, base_noise=z)
    assert torch.allclose(o['shape_samples'].cpu(), ref['shape_samples'], atol=1e-6)
    assert (o['pose_rotmats_samples'].cpu() - ref['pose_rotmats_samples']).abs().max().item() <= ROT_TOL
    # own RNG path: finite rotations, different draws differ
    a = m(None, input_feats=feats.cuda(), num_samples=3)['pose_rotmats_samples']
    b = m(None, input_feats=feats.cuda(), num_samples=3)['pose_rotmats_samples']
    assert torch.isfinite(a).all() and (a - b).abs().max().item() > 1e-3
    # return_input_feats(_only)
    assert set(m(None, input_feats=feats.cuda(), return_input_feats_only=True).keys()) == {'input_feats'}
    assert 'input_feats' in m(None, input_feats=feats.cuda(), return_input_feats=True)


@pytest.mark.parametrize('scale,layers', [(1.0, 18), (1.5, 18), (1.0, 50)])
def test_log_prob_parity(scale, layers):
    """Teacher-forced log-likelihood (humaniflow_model.py:314-320; losses/humaniflow_loss.py:25-35):
    targets include theta<1e-6, theta~pi/2 and |pi-theta|<1e-2, plus the model's own samples."""
    m, sd, cfg = make_model(layers, seed=4, flow_scale=scale)
    m = m.cuda()
    Rt = special_rotations()                       # (40,3,3) f64
    B = Rt.shape[0]
    feats, z, se = _noise(B, 1, seed=6, feat_dim=m.input_feats_dim)
    g = torch.Generator().manual_seed(7)
    perm = torch.stack([torch.randperm(B, generator=g) for _ in range(23)], 1)       # (B,23)
    pose_R = Rt[perm].float()                                                          # (B,23,3,3), every joint sees every case
    own = _run(m, feats, z, se)['pose_rotmats_samples'][:, 0].cpu()
    pose_R[: B // 2] = own[: B // 2]
    shape = se[:, 0]
    glob = _run(m, feats, None, None)['glob_rotmat'].cpu()
    ref = om.forward(sd, cfg, SMPL_PARENTS, input_feats=feats, compute_point_est=False, shape_for_loglik=shape,
                     pose_R_for_loglik=pose_R, glob_R_for_loglik=glob)
    out = m(None, input_feats=feats.cuda(), compute_point_est=False, compute_for_loglik=True, shape_for_loglik=shape.cuda(),
            pose_R_for_loglik=pose_R.cuda(), glob_R_for_loglik=glob.cuda())
    dists = out['conditioned_pose_SO3flow_dists_for_loglik']
    assert len(dists) == 23 and len(out['conditioned_pose_so3flow_dists_for_loglik']) == 23
    ctx_ref = torch.stack(ref['loglik_contexts'], 1)
    assert torch.allclose(out['flow_contexts_for_loglik'].cpu(), ctx_ref, atol=2e-5, rtol=1e-5)
    lp = torch.stack([dists[j].log_prob(pose_R[:, j].double().cuda()) for j in range(23)], 1).cpu()
    assert lp.shape == (B, 23) and lp.dtype == torch.float32
    # Reference quirk kept on purpose: for 1e-10 < theta <~ 1e-8 the fp64 expression log((2-2cos t)/t^2) of
    # utils/rigid_transform_utils.py:298-314 cancels to log(0), so the reference's log_prob is +inf there
    # (special_rotations() contains theta = 1e-9).  Non-finite entries must agree exactly, finite ones to 1e-3.
    fin = torch.isfinite(ref['pose_loglik'])
    assert torch.equal(torch.isfinite(lp), fin) and torch.equal(lp[~fin], ref['pose_loglik'][~fin])
    assert fin.float().mean() > 0.95
    err = (lp - ref['pose_loglik']).abs()[fin] / ref['pose_loglik'].abs()[fin].clamp_min(1.0)
    assert err.max().item() <= LP_RTOL, err.max()
    # the batched entry point gives the same numbers
    lp_all = m.pose_log_prob(out['flow_contexts_for_loglik'], pose_R.cuda()).cpu()
    assert torch.equal(lp_all, lp)
    # density on the algebra (the `..._so3flow_...` objects)
    from oracle import flow as oflow
    v = torch.randn(B, 3, generator=g) * 0.8
    j = 11
    ref_alg = oflow.algebra_log_prob(om.joint_couplings(sd, j, 2), v, ref['loglik_contexts'][j], cfg.NORM_FLOW.COMPACT_SUPPORT_RADIUS, 0.6)
    got = out['conditioned_pose_so3flow_dists_for_loglik'][j].log_prob(v.cuda()).cpu()
    assert torch.isfinite(got).all()
    assert ((got - ref_alg).abs() / ref_alg.abs().clamp_min(1.0)).max().item() <= LP_RTOL


def test_density_consistency_property():
    """Oracle-free: log_prob(exp(v)) of a sample equals the single-pre-image change of variables whenever the
    other pre-images fall outside the support (|v| < pi/2): log p(v) - log|det J_exp(v)|."""
    m, sd, cfg = make_model(18, seed=8)
    m = m.cuda()
    B = 64
    feats, z, se = _noise(B, 1, seed=9)
    o = _run(m, feats, z * 0.3, se, compute_point_est=False)
    R = o['pose_rotmats_samples'][:, 0]
    glob = o['glob_rotmat']
    out = m(None, input_feats=feats.cuda(), compute_point_est=False, compute_for_loglik=True, shape_for_loglik=se[:, 0].cuda(),
            pose_R_for_loglik=R, glob_R_for_loglik=glob)
    lp = m.pose_log_prob(out['flow_contexts_for_loglik'], R)
    from oracle import so3
    v = so3.so3_log(R.double().cpu())
    small = v.norm(dim=-1) < math.pi / 2 - 0.05
    lp_alg = torch.stack([out['conditioned_pose_so3flow_dists_for_loglik'][j].log_prob(v[:, j].float().cuda()) for j in range(23)], 1).cpu()
    expect = lp_alg - so3.so3_log_abs_det_jacobian(v).float()
    assert small.float().mean() > 0.5
    assert ((lp.cpu() - expect).abs()[small]).max().item() <= 1e-3


def test_state_dict_roundtrip_and_repack():
    """Reference checkpoints load with strict=True (same key names); changing weights re-packs the kernels' copies."""
    m, sd, cfg = make_model(18, seed=10)
    m2, sd2, _ = make_model(18, seed=11)
    m2 = m2.cuda()
    feats, z, se = _noise(2, 3, seed=12)
    before = _run(m2, feats, z, se)['pose_rotmats_samples'].clone()
    missing = m2.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    after = _run(m2, feats, z, se)['pose_rotmats_samples']
    ref = om.forward(sd, cfg, SMPL_PARENTS, input_feats=feats, num_samples=3, shape_eps=se, base_noise=z)
    assert (after.cpu() - ref['pose_rotmats_samples']).abs().max().item() <= ROT_TOL
    assert (after - before).abs().max().item() > 1e-3


@pytest.mark.parametrize('M,K,O,act,acc', [(32, 2048, 512, 1, 0), (32, 2048, 256, 0, 0), (5, 512, 29, 0, 0), (33, 128, 7, 2, 1),
                                            (1, 64, 4, 0, 0), (32, 9, 256, 0, 1), (32, 2048, 1024, 1, 0), (32, 1024, 29, 0, 1),
                                            (70, 1024, 40, 2, 0)])
def test_linear_layers(M, K, O, act, acc):
    """hf_linear / hf_linear_ws (humaniflow_model.py:232-258 fc1 / heads / image-level features) against torch fp32 on the CPU:
    ragged row tiles, neuron counts that are not a multiple of the CTA tile, ELU / ReLU, accumulate-into-output; with a
    workspace the layer is K-sliced across CTAs (same tolerance, and the same result on every call)."""
    import torch.nn.functional as F
    from humaniflow_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(M * 7 + K)
    x = torch.randn(M, K, generator=g)
    W = torch.randn(O, K, generator=g) / K ** 0.5
    b = torch.randn(O, generator=g)
    y0 = torch.randn(M, O, generator=g)
    ref = x.double() @ W.double().T + b.double() + (y0.double() if acc else 0)
    ref = F.elu(ref) if act == 1 else (F.relu(ref) if act == 2 else ref)
    xd, Wd, bd, yd = x.cuda(), W.cuda(), b.cuda(), y0.clone().cuda()
    _lib.check(lib.hf_linear(_lib.ptr(xd), K, _lib.ptr(Wd), K, _lib.ptr(bd), _lib.ptr(yd), O, M, K, O, act, acc, _lib.stream()))
    torch.cuda.synchronize()
    assert (yd.cpu().double() - ref).abs().max().item() <= 2e-5
    nbytes = lib.hf_linear_workspace_bytes(M, K, O)
    ws = torch.empty(max(nbytes, 16), device='cuda', dtype=torch.uint8)
    outs = []
    for _ in range(2):
        y2 = y0.clone().cuda()
        _lib.check(lib.hf_linear_ws(_lib.ptr(xd), K, _lib.ptr(Wd), K, _lib.ptr(bd), _lib.ptr(y2), O, M, K, O, act, acc,
                                    _lib.ptr(ws), nbytes, _lib.stream()))
        torch.cuda.synchronize()
        outs.append(y2.cpu())
    assert (outs[0].double() - ref).abs().max().item() <= 2e-5
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize('scale', [1.0, 1.5])
def test_log_prob_backward_against_autograd_through_the_oracle(scale):
    """d log_prob / d (context, target) of both densities (hf_flow_log_prob_backward,
    hf_flow_algebra_log_prob_backward) against torch.autograd through the oracle, weights fixed: rotations incl. theta ~ 0 and
    theta ~ pi/2, random contexts, random upstream weights.  Tolerance: 2e-3 of the largest gradient entry per tensor (fp32
    chain of two spline couplings; autograd differentiates the same fp32 arithmetic in a different order)."""
    from humaniflow_b200 import _lib
    from oracle import flow as oflow
    lib = _lib.load()
    m, sd, cfg = make_model(18, seed=8, flow_scale=scale)
    m = m.cuda()
    m._ensure_packed(torch.device('cuda'))
    nf = cfg.NORM_FLOW
    J, T = 23, nf.NUM_TRANSFORMS
    g = torch.Generator().manual_seed(3)
    from oracle import so3
    import numpy as np
    rs = np.random.RandomState(11)
    axes = rs.standard_normal((16, 3))
    axes /= np.linalg.norm(axes, axis=1, keepdims=True)
    # (theta = pi / 2 exactly puts the second pre-image ON the edge of the support, where the density diverges and the mask is a
    #  rounding decision: not a point to compare gradients at)
    ang = np.concatenate([rs.uniform(0.05, 3.0, 9), [1e-3, math.pi / 2 - 1e-3, math.pi / 2 + 1e-3, 1.9, 2.0, 2.5, 3.0]])
    Rt = so3.so3_exp(torch.tensor(axes * ang[:, None], dtype=torch.float64))
    Rn = Rt.shape[0]
    perm = torch.stack([torch.randperm(Rn, generator=g) for _ in range(J)], 1)
    pose = Rt[perm].double().contiguous()                                   # (Rn,J,3,3)
    ctx = (torch.randn(Rn, J, 64, generator=g) * 0.5).float()
    w = torch.randn(Rn, J, generator=g).float()
    valg = (torch.randn(Rn, J, 3, generator=g) * 0.8).float()

    ctx_r = ctx.clone().requires_grad_()
    pose_r = pose.clone().requires_grad_()
    valg_r = valg.clone().requires_grad_()
    ctx_a = ctx.clone().requires_grad_()
    tot = 0.
    tot_a = 0.
    for j in range(J):
        cpl = om.joint_couplings(sd, j, T)
        tot = tot + (w[:, j] * oflow.so3_log_prob(cpl, pose_r[:, j], ctx_r[:, j], nf.COMPACT_SUPPORT_RADIUS, nf.BASE_DIST_STD)).sum()
        tot_a = tot_a + (w[:, j] * oflow.algebra_log_prob(cpl, valg_r[:, j], ctx_a[:, j], nf.COMPACT_SUPPORT_RADIUS, nf.BASE_DIST_STD)).sum()
    tot.backward()
    tot_a.backward()

    cd, pd, wd, vd = ctx.cuda().contiguous(), pose.cuda().contiguous(), w.cuda().contiguous(), valg.cuda().contiguous()
    g_ctx = torch.empty(Rn, J, 64, device='cuda')
    g_rot = torch.empty(Rn, J, 3, 3, device='cuda', dtype=torch.float64)
    _lib.check(lib.hf_flow_log_prob_backward(m._flow, _lib.ptr(cd), J * 64, 0, J, _lib.ptr(pd), _lib.ptr(wd), Rn, _lib.ptr(g_ctx), _lib.ptr(g_rot),
                                             _lib.stream()))
    g_ctx_a = torch.empty(Rn, J, 64, device='cuda')
    g_v = torch.empty(Rn, J, 3, device='cuda')
    _lib.check(lib.hf_flow_algebra_log_prob_backward(m._flow, _lib.ptr(cd), J * 64, 0, J, _lib.ptr(vd), _lib.ptr(wd), Rn, _lib.ptr(g_ctx_a),
                                                     _lib.ptr(g_v), _lib.stream()))
    torch.cuda.synchronize()
    for name, got, ref in (('algebra ctx', g_ctx_a, ctx_a.grad), ('algebra v', g_v, valg_r.grad), ('SO3 ctx', g_ctx, ctx_r.grad),
                           ('SO3 rot', g_rot, pose_r.grad)):
        err = (got.double().cpu() - ref.double()).abs().max().item()
        sc = ref.abs().max().item()
        print('%-12s max err %.2e of %.2e' % (name, err, sc))
        assert err <= 2e-3 * sc, (name, err, sc)


def test_pose_prior_gradient_of_a_fitting_loop():
    """The pose-prior term of optimise/optimise_humaniflow.py:96-114 end to end: axis-angle pose, shape and global rotation are
    leaf tensors; contexts come from forward(compute_for_loglik=True), the loss is -sum_j log p_j(R_j | ctx_j).  Gradients through
    hf_flow_log_prob_backward + the context backward against torch.autograd through the oracle; then a few descent steps raise
    the log-likelihood."""
    from oracle import so3
    m, sd, cfg = make_model(18, seed=12)
    m = m.cuda()
    B = 6
    g = torch.Generator().manual_seed(21)
    feats = torch.randn(B, 512, generator=g).abs()
    pose_aa = (torch.randn(B, 23, 3, generator=g) * 0.3)
    glob_aa = (torch.randn(B, 3, generator=g) * 0.3)
    shape = torch.randn(B, 10, generator=g)

    def rodrigues(aa):                 # differentiable torch Rodrigues (the fitting script uses smplx's batch_rodrigues)
        if aa.is_cuda:
            from humaniflow_b200.smpl import _rodrigues_torch
            return _rodrigues_torch(aa.reshape(-1, 3)).view(*aa.shape[:-1], 3, 3)
        return so3.batch_rodrigues(aa.reshape(-1, 3)).view(*aa.shape[:-1], 3, 3)

    pa, ga, sh = pose_aa.clone().requires_grad_(), glob_aa.clone().requires_grad_(), shape.clone().requires_grad_()
    ref = om.forward(sd, cfg, SMPL_PARENTS, input_feats=feats, compute_point_est=False, shape_for_loglik=sh,
                     pose_R_for_loglik=rodrigues(pa), glob_R_for_loglik=rodrigues(ga))
    (-ref['pose_loglik'].sum() / B).backward()

    pc, gc, sc = pose_aa.clone().cuda().requires_grad_(), glob_aa.clone().cuda().requires_grad_(), shape.clone().cuda().requires_grad_()
    fc = feats.cuda()
    vals = []
    for it in range(4):
        pose_R = rodrigues(pc)
        out = m(None, input_feats=fc, compute_point_est=False, compute_for_loglik=True, shape_for_loglik=sc, pose_R_for_loglik=pose_R,
                glob_R_for_loglik=rodrigues(gc))
        dists = out['conditioned_pose_SO3flow_dists_for_loglik']
        lp = sum(dists[j].log_prob(pose_R[:, j].double()).sum() for j in range(23))
        loss = -lp / B
        pc.grad = gc.grad = sc.grad = None
        loss.backward()
        if it == 0:
            assert abs(loss.item() - (-ref['pose_loglik'].sum().item() / B)) <= 1e-3 * abs(loss.item())
            for name, got, want in (('pose', pc.grad, pa.grad), ('glob', gc.grad, ga.grad), ('shape', sc.grad, sh.grad)):
                err = (got.cpu() - want).abs().max().item()
                print('%-6s max err %.2e of %.2e' % (name, err, want.abs().max().item()))
                assert err <= 3e-3 * want.abs().max().item(), (name, err)
        vals.append(loss.item())
        with torch.no_grad():
            pc -= 0.01 * pc.grad
    assert vals[-1] < vals[0], vals


def test_log_prob_backward_at_the_log_map_edge_cases():
    """The same gradients where the log map switches formulas: theta within 1e-2 of pi (the reference's sqrt-of-diagonal branch),
    theta -> 0 (series branch).  Per-row tolerance 1e-3 of that row's largest entry (near pi the rotation gradient is ~1e4-1e5)."""
    from humaniflow_b200 import _lib
    from oracle import flow as oflow, so3
    lib = _lib.load()
    m, sd, cfg = make_model(18, seed=8)
    m = m.cuda()
    m._ensure_packed(torch.device('cuda'))
    nf = cfg.NORM_FLOW
    ang = torch.tensor([3.1, math.pi - 9e-3, math.pi - 5e-3, math.pi - 1e-3, math.pi - 1e-4, 1e-4, 1e-6, 1e-3], dtype=torch.float64)
    axes = torch.nn.functional.normalize(torch.randn(8, 3, generator=torch.Generator().manual_seed(1)), dim=1).double()
    Rt = so3.so3_exp(axes * ang[:, None])
    Rn, J, j = 8, 23, 5
    pose = Rt[:, None].expand(Rn, J, 3, 3).contiguous()
    ctx = torch.randn(Rn, J, 64, generator=torch.Generator().manual_seed(3)) * 0.5
    ctx_r, pose_r = ctx.clone().requires_grad_(), pose.clone().requires_grad_()
    oflow.so3_log_prob(om.joint_couplings(sd, j, 2), pose_r[:, j], ctx_r[:, j], nf.COMPACT_SUPPORT_RADIUS, nf.BASE_DIST_STD).sum().backward()
    cd, pd, wd = ctx.cuda(), pose.cuda(), torch.ones(Rn, J, device='cuda')
    g_ctx = torch.zeros(Rn, J, 64, device='cuda')
    g_rot = torch.zeros(Rn, J, 3, 3, device='cuda', dtype=torch.float64)
    _lib.check(lib.hf_flow_log_prob_backward(m._flow, _lib.ptr(cd), J * 64, 0, J, _lib.ptr(pd), _lib.ptr(wd), Rn, _lib.ptr(g_ctx), _lib.ptr(g_rot),
                                             _lib.stream()))
    torch.cuda.synchronize()
    for r in range(Rn):
        for got, ref in ((g_ctx[r, j].cpu(), ctx_r.grad[r, j]), (g_rot[r, j].cpu(), pose_r.grad[r, j])):
            assert (got.double() - ref.double()).abs().max().item() <= 1e-3 * ref.abs().max().item(), (r, ang[r].item())
