"""Oracle-free invariants of the CPU oracle.  The upstream pieces (pyro spline coupling, smplx LBS)
are unpinned — the reference has no tests for them — so they are held to the mathematical properties any
correct implementation must satisfy."""
import math

import torch

from oracle import flow as oflow
from oracle import model as om
from oracle import smpl as osmpl
from oracle import so3
from oracle import spline as osp
from util import RADIUS, make_model, smpl_data
from humaniflow_b200.synthetic import SMPL_PARENTS


def _couplings(dtype=torch.float64, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    dims = [(65, 64), (64, 32), (32, 32), (32, 62)]
    mk = lambda o, i: ((torch.rand(o, i, generator=g) * 2 - 1) / math.sqrt(i) * scale).to(dtype)
    return [[(mk(o, i), mk(o, 1)[:, 0]) for (i, o) in dims] for _ in range(2)]


def test_spline_inverse_and_logdet_cancel():
    cps = _couplings(scale=2.0)
    g = torch.Generator().manual_seed(1)
    ctx = torch.randn(512, 64, generator=g, dtype=torch.float64)
    x = torch.randn(512, 3, generator=g, dtype=torch.float64) * 2.5
    x[0] = torch.tensor([0.3, RADIUS + 1.0, -RADIUS - 2.0])      # identity outside the box
    y, ld = osp.coupling_forward(cps[0], x, ctx, RADIUS)
    xb, ldb = osp.coupling_inverse(cps[0], y, ctx, RADIUS)
    assert torch.allclose(xb, x, atol=1e-9) and torch.allclose(ld, ldb, atol=1e-8)
    assert torch.equal(y[0], x[0]) and ld[0] == 0
    assert torch.equal(y[:, 0], x[:, 0])                           # identity=True on the conditioning coordinate
    # log-det equals the autograd Jacobian
    xs = x[1:9].clone().requires_grad_(True)
    J = torch.autograd.functional.jacobian(lambda t: osp.coupling_forward(cps[0], t, ctx[1:9], RADIUS)[0].sum(0), xs)
    Jd = torch.stack([J[:, i, :] for i in range(8)])              # (8,3,3) per-sample Jacobians
    assert torch.allclose(torch.linalg.slogdet(Jd)[1], ld[1:9], atol=1e-7)


def test_spline_is_monotone_and_continuous_at_knots():
    cps = _couplings(scale=1.0, seed=3)
    steps = []
    for n in (4001, 40001):
        ctx = torch.zeros(n, 64, dtype=torch.float64)
        t = torch.linspace(-RADIUS, RADIUS, n, dtype=torch.float64)
        x = torch.stack([torch.zeros_like(t), t, t], -1)
        y, _ = osp.coupling_forward(cps[0], x, ctx, RADIUS)
        assert (y[1:, 1] > y[:-1, 1]).all() and (y[1:, 2] > y[:-1, 2]).all()
        assert abs(y[0, 1] + RADIUS) < 1e-9 and abs(y[-1, 1] - RADIUS) < 1e-9
        steps.append(max((y[1:, 1] - y[:-1, 1]).max().item(), (y[1:, 2] - y[:-1, 2]).max().item()))
    # no jumps across bin edges: the largest increment shrinks with the grid spacing (a discontinuity would not)
    assert steps[1] < 0.2 * steps[0] and steps[1] < 0.05, steps


def test_flow_density_integrates_to_one():
    """The so(3)-algebra density of one conditioned flow integrates to ~1 over the support ball (quadrature)."""
    cps = _couplings(scale=1.0, seed=5)
    n = 121                                                        # 61 under-resolves steep spline bins (mass 1.17)
    ax = torch.linspace(-RADIUS, RADIUS, n, dtype=torch.float64)
    grid = torch.stack(torch.meshgrid(ax, ax, ax, indexing='ij'), -1).reshape(-1, 3)
    inside = grid.norm(dim=-1) < RADIUS - 1e-6
    v = grid[inside]
    ctx = torch.zeros(1, 64, dtype=torch.float64).expand(v.shape[0], -1)
    lp = oflow.algebra_log_prob(cps, v, ctx, RADIUS, 0.6)
    mass = lp.exp().sum() * (ax[1] - ax[0]) ** 3
    assert abs(mass.item() - 1.0) < 0.03, mass


def test_log_prob_of_sample_equals_sampling_path_density():
    """log_prob(rsample) == base log-prob - forward log-dets - log|det J_exp| when the other pre-images are
    outside the support (|v| < pi/2)."""
    cps = _couplings(dtype=torch.float32, seed=7)
    g = torch.Generator().manual_seed(8)
    ctx = torch.randn(256, 64, generator=g)
    z = torch.randn(256, 3, generator=g) * 0.3
    v, ld = oflow.flow_forward(cps, z, ctx, RADIUS, with_logdet=True)
    R = so3.so3_exp(v.double())
    lp = oflow.so3_log_prob(cps, R, ctx, RADIUS, 0.6)
    expect = oflow.normal_log_prob(z, 0.6) - ld - so3.so3_log_abs_det_jacobian(v.double()).float()
    small = v.norm(dim=-1) < math.pi / 2 - 0.05
    assert small.float().mean() > 0.2 and small.sum() >= 64
    assert (lp - expect).abs()[small].max() < 2e-3


def test_exp_log_roundtrip_and_rot6d():
    g = torch.Generator().manual_seed(2)
    v = torch.randn(1000, 3, generator=g, dtype=torch.float64)
    v = v / v.norm(dim=-1, keepdim=True) * torch.rand(1000, 1, generator=g, dtype=torch.float64) * (math.pi - 0.02)
    R = so3.so3_exp(v)
    assert torch.allclose(R @ R.transpose(-1, -2), torch.eye(3, dtype=torch.float64).expand_as(R), atol=1e-12)
    assert torch.allclose(torch.linalg.det(R), torch.ones(1000, dtype=torch.float64), atol=1e-12)
    assert torch.allclose(so3.so3_log(R), v, atol=1e-6)
    assert torch.allclose(so3.rot6d_to_rotmat(so3.rotmat_to_rot6d(R.float())), R.float(), atol=1e-6)
    # all three pre-images map to the same rotation
    xs = so3.so3_xset(v)
    assert torch.allclose(so3.so3_exp(xs[0]), R, atol=1e-9) and torch.allclose(so3.so3_exp(xs[1]), R, atol=1e-9)


def test_lbs_identity_pose_and_chain():
    data = smpl_data()
    g = torch.Generator().manual_seed(3)
    betas = torch.randn(3, 10, generator=g)
    eye = torch.eye(3).expand(3, 24, 3, 3)
    v, j = osmpl.smpl_forward(data, betas, eye[:, 1:], eye[:, :1], pose2rot=False)
    v_shaped = data['v_template'][None] + torch.einsum('bl,mkl->bmk', betas, data['shapedirs'])
    assert torch.allclose(v, v_shaped, atol=1e-6)
    assert torch.allclose(j[:, :24], torch.einsum('bik,ji->bjk', v_shaped, data['J_regressor']), atol=1e-6)
    assert j.shape == (3, 90, 3)
    # rotating only the root rotates every vertex rigidly about the root joint
    Rg = so3.batch_rodrigues(torch.tensor([[0.3, -0.2, 0.9]])).expand(3, 1, 3, 3)
    v2, j2 = osmpl.smpl_forward(data, betas, eye[:, 1:], Rg, pose2rot=False)
    root = j[:, :1]
    assert torch.allclose(v2, torch.einsum('bij,bvj->bvi', Rg[:, 0], v - root) + root, atol=1e-5)


def test_model_forward_structure():
    m, sd, cfg = make_model(18, seed=0)
    g = torch.Generator().manual_seed(1)
    B, N = 3, 5
    feats = torch.randn(B, 512, generator=g).abs()
    z = torch.randn(B, N, 23, 3, generator=g) * 0.6
    se = torch.randn(B, N, 10, generator=g)
    out = om.forward(sd, cfg, SMPL_PARENTS, input_feats=feats, num_samples=N, shape_eps=se, base_noise=z)
    R = out['pose_rotmats_samples']
    assert R.shape == (B, N, 23, 3, 3) and R.dtype == torch.float32
    assert torch.allclose(R @ R.transpose(-1, -2), torch.eye(3).expand_as(R), atol=1e-5)
    G = out['glob_rotmat']
    assert torch.allclose(G @ G.transpose(-1, -2), torch.eye(3).expand_as(G), atol=1e-5)
    # ancestor conditioning: changing the noise of joint 0 (SMPL joint 1) leaves siblings 1,2 untouched
    # but changes its descendants (3, 6, 9, ...)
    z2 = z.clone()
    z2[:, :, 0] += 0.5
    out2 = om.forward(sd, cfg, SMPL_PARENTS, input_feats=feats, num_samples=N, shape_eps=se, base_noise=z2, compute_point_est=False)
    d = (out2['pose_rotmats_samples'] - R).abs().amax(dim=(0, 1, 3, 4))
    anc = om.ancestors_of(SMPL_PARENTS)
    for j in range(23):
        if j == 0 or 0 in anc[j]:
            assert d[j] > 1e-6, j
        else:
            assert d[j] == 0, j
