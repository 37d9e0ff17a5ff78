"""Model-level parity of the CUDA path at the benchmarked configuration and from the image.

1. CUDA ``HumaniflowModel`` against golden vectors of the REAL reference class (tests/golden/model_golden.npz, made by
   tests/golden/make_golden_model.py from models/humaniflow_model.py with injected noise), ResNet-18 and ResNet-50 widths.
2. Config 3 end to end: image -> ResNet-50 (bf16 tensor-core encoder) -> heads -> flow -> SMPL vertices, against the fp32
   CPU oracle fed the SAME image.  The bf16 encoder's feature error (rel L2 ~4e-3, tests/test_gpu_encoder.py) is propagated
   through heads, the 23-joint chain and the LBS here; the tolerances below are what that allows and are far looser than the
   1e-4 m bar that holds for identical (beta, theta) (tests/test_gpu_lbs.py): see DESIGN.md 5.
3. Sharding: running the image axis in 2 or 3 shards reproduces the unsharded result bit for bit.
"""
import os

import numpy as np
import pytest
import torch

import humaniflow_b200 as hb
from detweights import det_input, fill_state_dict
from humaniflow_b200.sharding import shard_range
from humaniflow_b200.synthetic import SMPL_PARENTS, synthetic_proxy_input
from oracle import model as om
from oracle import smpl as osmpl
from oracle import so3
from util import GOLDEN, make_model, smpl_data

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('layers', [18, 50])
def test_cuda_model_against_real_reference_golden(layers):
    """Same weights / features / noise as make_golden_model.py; tolerances: heads 2e-5, rotations 5e-5, log_prob rel 1e-3."""
    g = np.load(os.path.join(GOLDEN, 'model_golden.npz'))
    t = lambda k: torch.tensor(g['r%d_%s' % (layers, k)])
    cfg = hb.get_model_cfg_defaults()
    cfg.NUM_RESNET_LAYERS = layers
    m = hb.HumaniflowModel('cpu', cfg, SMPL_PARENTS).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.startswith('image_encoder.')}
    sd = fill_state_dict(shapes, seed=700 + layers)
    sd = {k: (v * 0.5 if v.dim() >= 2 else v) for k, v in sd.items()}          # as in make_golden_model.py
    sd['init_glob'], sd['init_cam'] = m.state_dict()['init_glob'].clone(), m.state_dict()['init_cam'].clone()
    missing = m.load_state_dict(sd, strict=False)
    assert all(k.startswith('image_encoder.') for k in missing.missing_keys) and not missing.unexpected_keys
    m = m.cuda()
    B, N, F = 5, 4, m.input_feats_dim
    feats = det_input((B, F), 710 + layers).abs()
    shape_eps = det_input((N, B, 10), 711 + layers).transpose(0, 1).contiguous()
    base_noise = torch.stack([det_input((B, N, 3), 720 + layers + j) for j in range(23)], 2) * 0.6
    out = m(None, input_feats=feats.cuda(), num_samples=N, base_noise=base_noise.cuda(), shape_eps=shape_eps.cuda())
    for k in ('cam_wp', 'glob_rotmat', 'shape_mode', 'shape_log_std', 'shape_samples'):
        assert torch.allclose(out[k].cpu(), t(k), atol=2e-5, rtol=1e-5), k
    assert torch.allclose(out['shape_dist_for_loglik'].scale.cpu(), t('scale_of_shape_dist'), atol=2e-5, rtol=1e-5)
    for k in ('pose_axisangle_point_est', 'pose_rotmats_point_est', 'pose_rotmats_samples'):
        err = (out[k].cpu() - t(k)).abs().max().item()
        assert err <= 5e-5, (k, err)
    v_t = det_input((B, 23, 3), 760 + layers) * 0.7
    v_t[0, 3] = 0.0
    v_t[1, 5] = v_t[1, 5] / v_t[1, 5].norm() * 2.6
    R_t = so3.so3_exp(v_t.double()).float()
    shape_t = det_input((B, 10), 761 + layers)
    glob_t = so3.so3_exp(det_input((B, 3), 762 + layers).double() * 0.5).float()
    o = m(None, input_feats=feats.cuda(), compute_point_est=False, compute_for_loglik=True, shape_for_loglik=shape_t.cuda(),
          pose_R_for_loglik=R_t.cuda(), glob_R_for_loglik=glob_t.cuda())
    assert torch.allclose(o['flow_contexts_for_loglik'].cpu(), t('ctx_ll'), atol=2e-5, rtol=1e-5)
    lp = torch.stack([o['conditioned_pose_SO3flow_dists_for_loglik'][j].log_prob(R_t[:, j].double().cuda()) for j in range(23)], 1).cpu()
    ref = t('lp_SO3')
    fin = torch.isfinite(ref)
    assert torch.equal(torch.isfinite(lp), fin)
    assert ((lp - ref).abs()[fin] / ref.abs()[fin].clamp_min(1.0)).max().item() <= 1e-3
    v_alg = det_input((B, 23, 3), 763 + layers) * 0.8
    lpa = torch.stack([o['conditioned_pose_so3flow_dists_for_loglik'][j].log_prob(v_alg[:, j].cuda()) for j in range(23)], 1).cpu()
    assert ((lpa - t('lp_so3')).abs() / t('lp_so3').abs().clamp_min(1.0)).max().item() <= 1e-3


def _path(m, smpl, x, z, se):
    """image -> meshes through the public classes (what predict_humaniflow.py:112-160 does)."""
    B, N = z.shape[:2]
    out = m(x, num_samples=N, base_noise=z, shape_eps=se, return_input_feats=True)
    R = out['pose_rotmats_samples'].reshape(B * N, 23, 3, 3)
    glob = out['glob_rotmat'][:, None].expand(-1, N, -1, -1).reshape(B * N, 1, 3, 3)
    so = smpl(betas=out['shape_samples'].reshape(B * N, 10), body_pose=R, global_orient=glob, pose2rot=False)
    return out, so


# Tolerances of the image -> vertex path with the bf16 encoder, ~4x what was measured on B200 at B=32, N=100:
# features rel-L2 3.0e-3, heads 1.2e-3, rotation entries max 2.8e-4 / mean 2.2e-5, vertex L2 max 1.05e-3 m / mean 1.4e-4 m.
# I.e. from the IMAGE the 1e-4 m bar is met on average only; it is met strictly for identical (beta, theta) (test_gpu_lbs.py).
E2E_FEAT_REL = 1.2e-2
E2E_ROT_MAX = 1.2e-3        # max |dR| entry over all samples / joints
E2E_ROT_MEAN = 1e-4
E2E_VERT_MAX = 4e-3         # metres, max over all vertices of all samples
E2E_VERT_MEAN = 6e-4        # metres, mean vertex L2


@pytest.mark.parametrize('B,N', [(2, 10), (32, 100)])
def test_config3_image_to_vertices(B, N):
    """BASELINE configs[2]: (B,18,256,256) -> ResNet-50 -> flow -> N SMPL meshes per image vs the fp32 oracle on the same image."""
    m, sd, cfg = make_model(50, seed=40, res_gain=0.3)
    m = m.cuda()
    data = smpl_data()
    smpl = hb.SMPL.from_arrays(data, create_transl=False).cuda()
    x = synthetic_proxy_input(B, 18, 256, seed=41)
    g = torch.Generator().manual_seed(42)
    z = torch.randn(B, N, 23, 3, generator=g) * 0.6
    se = torch.randn(B, N, 10, generator=g)
    out, so = _path(m, smpl, x.cuda(), z.cuda(), se.cuda())
    with torch.no_grad():
        ref = om.forward(sd, cfg, SMPL_PARENTS, input=x, num_samples=N, shape_eps=se, base_noise=z)
        Rr = ref['pose_rotmats_samples'].reshape(B * N, 23, 3, 3)
        gr = ref['glob_rotmat'][:, None].expand(-1, N, -1, -1).reshape(B * N, 1, 3, 3)
        v_ref, j_ref = osmpl.smpl_forward(data, ref['shape_samples'].reshape(B * N, 10), Rr, gr, pose2rot=False)
    f_rel = ((out['input_feats'].cpu() - ref['input_feats']).norm() / ref['input_feats'].norm()).item()
    dR = (out['pose_rotmats_samples'].cpu() - ref['pose_rotmats_samples']).abs()
    dv = (so.vertices.cpu() - v_ref).norm(dim=-1)
    dj = (so.joints.cpu() - j_ref).norm(dim=-1)
    heads = max((out[k].cpu() - ref[k]).abs().max().item() for k in ('cam_wp', 'glob_rotmat', 'shape_mode', 'shape_log_std'))
    print('config3 B=%d N=%d: feats rel %.2e, heads max %.2e, rot max %.2e mean %.2e, vertex L2 max %.3e mean %.3e m, joint L2 max %.3e m'
          % (B, N, f_rel, heads, dR.max().item(), dR.mean().item(), dv.max().item(), dv.mean().item(), dj.max().item()))
    assert f_rel <= E2E_FEAT_REL
    assert dR.max().item() <= E2E_ROT_MAX and dR.mean().item() <= E2E_ROT_MEAN
    assert dv.max().item() <= E2E_VERT_MAX and dv.mean().item() <= E2E_VERT_MEAN
    # the same CUDA flow + LBS fed the ORACLE's fp32 features meets the tight bars (rotations 2e-5, vertices 1e-4 m):
    # the encoder precision is the only source of the looser numbers above
    out2 = m(None, input_feats=ref['input_feats'].cuda(), num_samples=N, base_noise=z.cuda(), shape_eps=se.cuda())
    assert (out2['pose_rotmats_samples'].cpu() - ref['pose_rotmats_samples']).abs().max().item() <= 2e-5
    R2 = out2['pose_rotmats_samples'].reshape(B * N, 23, 3, 3)
    g2 = out2['glob_rotmat'][:, None].expand(-1, N, -1, -1).reshape(B * N, 1, 3, 3)
    so2 = smpl(betas=out2['shape_samples'].reshape(B * N, 10), body_pose=R2, global_orient=g2, pose2rot=False)
    assert (so2.vertices.cpu() - v_ref).norm(dim=-1).max().item() <= 1e-4


@pytest.mark.parametrize('layers,size,B,world', [(18, 64, 4, 2), (50, 64, 5, 3), (50, 256, 4, 2)])
def test_image_shards_reproduce_the_unsharded_rows_exactly(layers, size, B, world):
    """every rank runs the whole path on its image range with its slice of the global noise; the concatenated
    shards must equal the single-process result BIT FOR BIT (encoder features, heads, rotations, vertices, joints)."""
    m, sd, cfg = make_model(layers, seed=50, res_gain=0.3)
    m = m.cuda()
    smpl = hb.SMPL.from_arrays(smpl_data(), create_transl=False).cuda()
    N = 6
    x = synthetic_proxy_input(B, 18, size, seed=51).cuda()
    g = torch.Generator().manual_seed(52)
    z = (torch.randn(B, N, 23, 3, generator=g) * 0.6).cuda()
    se = torch.randn(B, N, 10, generator=g).cuda()
    full, so_full = _path(m, smpl, x, z, se)
    full = {k: v.clone() for k, v in full.items() if torch.is_tensor(v)}
    v_full, j_full = so_full.vertices.clone(), so_full.joints.clone()
    parts, vparts, jparts = [], [], []
    for r in range(world):
        a, b = shard_range(B, world, r)
        o, so = _path(m, smpl, x[a:b], z[a:b], se[a:b])
        parts.append({k: v.clone() for k, v in o.items() if torch.is_tensor(v)})
        vparts.append(so.vertices.clone())
        jparts.append(so.joints.clone())
    for k in ('input_feats', 'cam_wp', 'glob_rotmat', 'shape_mode', 'shape_log_std', 'shape_samples', 'pose_rotmats_samples',
              'pose_rotmats_point_est', 'pose_axisangle_point_est'):
        assert torch.equal(torch.cat([p[k] for p in parts], 0), full[k]), k
    assert torch.isfinite(v_full).all() and torch.isfinite(j_full).all()
    assert torch.equal(torch.cat(vparts, 0), v_full) and torch.equal(torch.cat(jparts, 0), j_full)


def test_optimise_pattern_end_to_end():
    """SURVEY 8f N3, the fitting loop of optimise/optimise_humaniflow.py:71-135 on the CUDA path: axis-angle pose, global rotation,
    shape and camera are leaf tensors; loss = w1 * 2-D joint error through SMPL (hf_lbs_forward / hf_lbs_backward) and the
    orthographic projection - w2 * pose prior (contexts + hf_flow_log_prob, CUDA backward) - w3 * shape prior.  The gradients of the
    first iteration match torch.autograd through the oracle (LBS + flow); a few descent steps lower the loss."""
    import humaniflow_b200 as hb
    from humaniflow_b200.smpl import _rodrigues_torch
    from oracle import so3, smpl as osmpl
    from util import smpl_data
    m, sd, cfg = make_model(18, seed=31)
    m = m.cuda()
    data = smpl_data()
    smpl = hb.SMPL.from_arrays(data, create_transl=False).cuda()
    B = 4
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(B, 512, generator=g).abs()
    pose0 = torch.randn(B, 69, generator=g) * 0.25
    glob0 = torch.randn(B, 3, generator=g) * 0.2
    shape0 = torch.randn(B, 10, generator=g) * 0.5
    cam0 = torch.tensor([[0.9, 0.0, 0.0]]).repeat(B, 1)
    target = torch.rand(B, 17, 2, generator=g) * 256
    joint_map = list(range(17, 34))            # any fixed 17 of the 90 output joints (picked + regressed ones)
    W1, W2, W3, IMG = 1e-4, 0.05, 0.1, 256.0

    def project(j3d, cam):                     # x-flip about pi + orthographic projection + undo normalisation (:80-88)
        j = torch.stack([j3d[..., 0], -j3d[..., 1]], -1)
        return (cam[:, None, :1] * (j + cam[:, None, 1:]) + 1.0) * 0.5 * IMG

    # ---- oracle
    po, go, so_, co = (t.clone().requires_grad_() for t in (pose0, glob0, shape0, cam0))
    _, j_ref = osmpl.smpl_forward(data, so_, po, go, pose2rot=True)
    j2d = ((project(j_ref[:, joint_map], co) - target) ** 2).mean()
    pose_R = so3.batch_rodrigues(po.reshape(-1, 3)).view(B, 23, 3, 3)
    ref = om.forward(sd, cfg, SMPL_PARENTS, input_feats=feats, compute_point_est=False, shape_for_loglik=so_, pose_R_for_loglik=pose_R,
                     glob_R_for_loglik=so3.batch_rodrigues(go))
    heads = om.forward(sd, cfg, SMPL_PARENTS, input_feats=feats)
    shape_lp = torch.distributions.Normal(heads['shape_mode'], torch.exp(heads['shape_log_std'])).log_prob(so_).sum() / B
    loss_ref = W1 * j2d - W2 * ref['pose_loglik'].sum() / B - W3 * shape_lp
    loss_ref.backward()

    # ---- CUDA
    pc, gc, sc, cc = (t.clone().cuda().requires_grad_() for t in (pose0, glob0, shape0, cam0))
    fc, tc = feats.cuda(), target.cuda()
    losses = []
    for it in range(5):
        out_s = smpl(body_pose=pc, global_orient=gc, betas=sc, pose2rot=True)
        j2d = ((project(out_s.joints[:, joint_map], cc) - tc) ** 2).mean()
        pose_R = _rodrigues_torch(pc.reshape(-1, 3)).view(B, 23, 3, 3)
        out = m(None, input_feats=fc, compute_point_est=False, num_samples=0, compute_for_loglik=True, shape_for_loglik=sc,
                pose_R_for_loglik=pose_R, glob_R_for_loglik=_rodrigues_torch(gc))
        dists = out['conditioned_pose_SO3flow_dists_for_loglik']
        pose_lp = sum(dists[j].log_prob(pose_R[:, j].double()).sum() for j in range(23)) / B
        shape_lp = out['shape_dist_for_loglik'].log_prob(sc).sum() / B
        loss = W1 * j2d - W2 * pose_lp - W3 * shape_lp
        for t in (pc, gc, sc, cc):
            t.grad = None
        loss.backward()
        if it == 0:
            assert abs(loss.item() - loss_ref.item()) <= 2e-3 * abs(loss_ref.item())
            for name, got, want in (('pose', pc.grad, po.grad), ('glob', gc.grad, go.grad), ('shape', sc.grad, so_.grad), ('cam', cc.grad, co.grad)):
                err = (got.cpu() - want).abs().max().item()
                assert err <= 5e-3 * want.abs().max().item(), (name, err, want.abs().max().item())
        losses.append(loss.item())
        with torch.no_grad():
            for t, lr in ((pc, 0.02), (gc, 0.02), (sc, 0.02), (cc, 1e-3)):
                t -= lr * t.grad / (t.grad.abs().max() + 1e-12)
    assert losses[-1] < losses[0], losses
