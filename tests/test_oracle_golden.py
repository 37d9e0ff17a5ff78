"""The oracle against golden vectors produced by the REAL reference files (tests/golden/make_golden.py).

This is what pins the oracle: rotation utilities, the radial tanh, ResNet, and the SO(3) log_prob plumbing
were run from /root/reference in the build container; the committed .npz hold their outputs.
"""
import math
import os

import numpy as np
import torch

from detweights import det_input, fill_state_dict
from oracle import flow as oflow
from oracle import so3
from oracle.resnet import resnet_forward

RADIUS = 1.5 * math.pi


def _load(golden_dir, name):
    return {k: torch.tensor(v) for k, v in np.load(os.path.join(golden_dir, name)).items()}


def test_so3_exp_log_xset_logdet(golden_dir):
    g = _load(golden_dir, 'so3_golden.npz')
    R = so3.so3_exp(g['v'])
    assert torch.allclose(R, g['R'], atol=1e-14, rtol=0)
    logv = so3.so3_log(g['R'].clone())
    assert torch.allclose(logv, g['logv'], atol=1e-12, rtol=0)
    xs = so3.so3_xset(g['logv'])
    both_nan = torch.isnan(xs) & torch.isnan(g['xset'])
    assert torch.allclose(torch.where(both_nan, torch.zeros_like(xs), xs),
                          torch.where(both_nan, torch.zeros_like(xs), g['xset']), atol=1e-12, rtol=0)
    assert torch.allclose(so3.so3_log_abs_det_jacobian(g['v']), g['lad'], atol=1e-13, rtol=0)
    assert torch.allclose(so3.so3_log_abs_det_jacobian(g['v'].float()), g['lad32'], atol=1e-6, rtol=0)


def test_rot6d(golden_dir):
    g = _load(golden_dir, 'so3_golden.npz')
    R = so3.rot6d_to_rotmat(g['r6'])
    assert torch.allclose(R, g['R6'], atol=1e-6, rtol=0)
    assert torch.equal(so3.rotmat_to_rot6d(g['R6']), g['back6'])


def test_radial_tanh(golden_dir):
    g = _load(golden_dir, 'rtanh_golden.npz')
    y = oflow.radial_tanh_forward(g['x'], RADIUS)
    assert torch.allclose(y, g['y'], atol=1e-6, rtol=1e-6)
    assert torch.allclose(oflow.radial_tanh_inverse(g['y'], RADIUS), g['xinv'], atol=1e-5, rtol=1e-5)
    assert torch.allclose(oflow.radial_tanh_log_abs_det(g['x'], g['y'], RADIUS), g['ld'], atol=1e-5, rtol=1e-5)


def _resnet_sd(layers):
    import humaniflow_b200.resnet as hr
    net = (hr.resnet18 if layers == 18 else hr.resnet50)(in_channels=18)
    shapes = {k: v.shape for k, v in net.state_dict().items()}
    sd = fill_state_dict(shapes, seed=100 + layers)
    return {'image_encoder.' + k: v for k, v in sd.items()}


def test_resnet_matches_reference_class(golden_dir):
    g = _load(golden_dir, 'resnet_golden.npz')
    for layers in (18, 50):
        sd = _resnet_sd(layers)
        x = det_input((2, 18, 64, 64), 200 + layers, kind='uniform')
        with torch.no_grad():
            f = resnet_forward(sd, x, layers)
        ref = g['feats%d' % layers]
        assert f.shape == ref.shape
        assert torch.allclose(f, ref, atol=1e-4 * ref.abs().max().item(), rtol=1e-4)


def _golden_flow():
    dims = [(65, 64), (64, 32), (32, 32), (32, 62)]
    shapes = {}
    for t in range(2):
        for l, (i, o) in enumerate(dims):
            shapes['c%d.%d.weight' % (t, l)] = (o, i)
            shapes['c%d.%d.bias' % (t, l)] = (o,)
    sd = fill_state_dict(shapes, seed=321)
    couplings = [[(sd['c%d.%d.weight' % (t, l)] * 0.5, sd['c%d.%d.bias' % (t, l)] * 0.5) for l in range(4)] for t in range(2)]
    return couplings, det_input((48, 64), 322)


def test_so3_log_prob_plumbing_matches_reference_class(golden_dir):
    """Real LocalDiffeoTransformedDistribution / SO3ExpCompactTransform / ToTransform / ScaledRadialTanhTransform
    over oracle-spline couplings vs oracle.flow.so3_log_prob (pre-images, masks, dtypes, logsumexp)."""
    g = _load(golden_dir, 'logprob_golden.npz')
    couplings, ctx = _golden_flow()
    lp = oflow.so3_log_prob(couplings, g['Rt'], ctx, RADIUS, 0.6)
    assert torch.allclose(lp, g['lp'], atol=2e-4, rtol=1e-4), (lp - g['lp']).abs().max()
    lpa = oflow.algebra_log_prob(couplings, g['vt'].float() * 0.9, ctx, RADIUS, 0.6)
    assert torch.allclose(lpa, g['lp_alg'], atol=2e-4, rtol=1e-4)
    v = oflow.flow_forward(couplings, g['z'], ctx, RADIUS)
    assert torch.allclose(v, g['v_alg'], atol=1e-6, rtol=1e-6)
    R = oflow.so3_sample(couplings, g['z'], ctx, RADIUS)
    assert torch.allclose(R, g['R_s'], atol=1e-12, rtol=0)
    lps = oflow.so3_log_prob(couplings, g['R_s'], ctx, RADIUS, 0.6)
    assert torch.allclose(lps, g['lp_s'], atol=2e-4, rtol=1e-4)


def test_real_regressors_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, 'regressors_sparse.npz'))
    assert tuple(g['extra_shape']) == (9, 6890) and tuple(g['cocoplus_shape']) == (19, 6890) and tuple(g['h36m_shape']) == (17, 6890)
    assert len(g['extra_val']) == 62 and len(g['cocoplus_val']) == 86 and len(g['h36m_val']) == 107


def test_sampling_helpers(golden_dir):
    """oracle/sampling.py against the real utils/sampling_utils.py / cam_utils.py / joints2d_utils.py outputs."""
    from oracle import sampling as osamp
    g = _load(golden_dir, 'sampling_golden.npz')
    avg, std = osamp.compute_vertex_variance_from_samples(g['verts'])
    assert torch.equal(avg, g['avg']) and torch.equal(std, g['std'])
    assert list(g['coco'].tolist()) == osamp.ALL_JOINTS_TO_COCO_MAP
    proj = osamp.project_joints2d(g['joints'], g['cam'], flip_x=False)
    assert torch.equal(proj, g['proj'])
    assert torch.equal(osamp.project_joints2d(g['joints'], g['cam'], flip_x=False, img_wh=256), g['pix'])
    flipped = osamp.project_joints2d(g['joints'], g['cam'], flip_x=True)
    assert torch.equal(flipped[..., 0], g['proj'][..., 0])      # rotation by pi about x leaves x alone, negates y


def test_metrics(golden_dir):
    """oracle/metrics.py against the real utils/eval_utils.py outputs (incl. a mirrored and a near-planar prediction)."""
    from oracle import metrics as omet
    g = np.load(os.path.join(golden_dir, 'metrics_golden.npz'))
    e = omet.pointset_errors(g['pred'], g['target'])
    for k in ('plain', 'sc', 'pa'):          # the oracle works in float64, the reference's numpy in float32
        assert np.allclose(e[k], g[k], rtol=2e-6, atol=1e-7), (k, np.abs(e[k] - g[k]).max())


def test_proxy_representation(golden_dir):
    """oracle/proxy_rep.py against the real models/canny_edge_detector.py and utils/label_conversions.py outputs."""
    from oracle import proxy_rep as opr
    g = np.load(os.path.join(golden_dir, 'proxy_golden.npz'))
    img = torch.tensor(g['img'])
    for thr, nms in ((0.0, True), (0.2, True), (0.1, False)):
        r = opr.canny(img, threshold=thr, nms=nms)
        key = 'thresholded_thin_edges' if nms else 'thresholded_grad_magnitude'
        assert torch.equal(r[key], torch.tensor(g['edge_%g_%d' % (thr, int(nms))])), (thr, nms)
    assert torch.equal(r['grad_magnitude'], torch.tensor(g['mag'])) and torch.equal(r['grad_orientation'], torch.tensor(g['ori']))
    assert torch.equal(opr.heatmaps(torch.tensor(g['j2d']), img.shape[-1]), torch.tensor(g['heat']))


def _model_problem(layers):
    """Weights / inputs / noise of tests/golden/make_golden_model.py, regenerated from the same seeds."""
    import humaniflow_b200 as hb
    from humaniflow_b200.synthetic import SMPL_PARENTS
    cfg = hb.get_model_cfg_defaults()
    cfg.NUM_RESNET_LAYERS = layers
    ours = hb.HumaniflowModel('cpu', cfg, SMPL_PARENTS)
    shapes = {k: tuple(v.shape) for k, v in ours.state_dict().items() if not k.startswith('image_encoder.')}
    sd = fill_state_dict(shapes, seed=700 + layers)
    sd = {k: (v * 0.5 if v.dim() >= 2 else v) for k, v in sd.items()}          # as in make_golden_model.py
    sd['init_glob'] = ours.state_dict()['init_glob'].clone()
    sd['init_cam'] = ours.state_dict()['init_cam'].clone()
    B, N, F = 5, 4, (512 if layers == 18 else 2048)
    feats = det_input((B, F), 710 + layers).abs()
    shape_eps = det_input((N, B, 10), 711 + layers).transpose(0, 1)               # rsample([N]).transpose(0, 1)
    base_noise = torch.stack([det_input((B, N, 3), 720 + layers + j) for j in range(23)], 2) * 0.6
    v_t = det_input((B, 23, 3), 760 + layers) * 0.7
    v_t[0, 3] = 0.0
    v_t[1, 5] = v_t[1, 5] / v_t[1, 5].norm() * 2.6
    tgt = {'R': so3.so3_exp(v_t.double()).float(), 'shape': det_input((B, 10), 761 + layers),
           'glob': so3.so3_exp(det_input((B, 3), 762 + layers).double() * 0.5).float(),
           'v_alg': det_input((B, 23, 3), 763 + layers) * 0.8}
    return cfg, SMPL_PARENTS, ours, shapes, sd, feats, shape_eps, base_noise, tgt


def test_model_glue_against_the_real_reference_class(golden_dir):
    """oracle/model.py::forward vs the REAL models/humaniflow_model.py::HumaniflowModel (heads, image-level features,
    ancestor-ordered contexts, 2j+t module index, point estimate, injected-noise samples, teacher-forced log-likelihood),
    at both encoder widths (ResNet-18: 512-d, ResNet-50: 2048-d features / 1024-d fc1)."""
    from oracle import model as om
    g = np.load(os.path.join(golden_dir, 'model_golden.npz'))
    for layers in (18, 50):
        t = lambda k: torch.tensor(g['r%d_%s' % (layers, k)])
        cfg, parents, ours, shapes, sd, feats, shape_eps, base_noise, tgt = _model_problem(layers)
        # state-dict key names (and with them the key -> joint mapping) are the reference's
        assert sorted(shapes.keys()) == [str(k) for k in g['r%d_keys' % layers]]
        anc = om.ancestors_of(parents)
        assert [len(a) for a in anc] == t('anc_len').tolist() and [x for a in anc for x in a] == t('anc_flat').tolist()
        with torch.no_grad():
            ref = om.forward(sd, cfg, parents, input_feats=feats, num_samples=4, shape_eps=shape_eps, base_noise=base_noise)
        for k in ('cam_wp', 'glob_rotmat', 'shape_mode', 'shape_log_std', 'shape_samples'):
            assert torch.allclose(ref[k], t(k), atol=1e-5, rtol=1e-5), (layers, k)
        assert torch.allclose(torch.exp(ref['shape_log_std']), t('scale_of_shape_dist'), atol=1e-6, rtol=1e-6)
        # fp32 on both sides, but the 23-joint chain amplifies the last-bit differences of differently blocked / threaded
        # matmuls (observed up to 4e-5 with these N(0, 2/fan_in) weights); a glue error (order, index, concat) is O(1)
        for k in ('pose_axisangle_point_est', 'pose_rotmats_point_est', 'pose_rotmats_samples'):
            assert (ref[k] - t(k)).abs().max().item() <= 2e-5, (layers, k, (ref[k] - t(k)).abs().max().item())
        # contexts the reference computed on the way (recorded around compute_flow_context)
        with torch.no_grad():
            fe_pe = om.image_level_feats(sd, feats, ref['shape_mode'], ref['glob_rotmat'], ref['cam_wp'])
            fe_s = om.image_level_feats(sd, feats, ref['shape_samples'], ref['glob_rotmat'], ref['cam_wp'])
            for j in range(23):
                c = om.flow_context(sd, j, anc[j], fe_pe, ref['pose_rotmats_point_est'])
                assert torch.allclose(c, t('ctx_pe')[:, j], atol=1e-4, rtol=1e-4), (layers, j)
                c = om.flow_context(sd, j, anc[j], fe_s, ref['pose_rotmats_samples'])
                assert torch.allclose(c, t('ctx_s')[:, :, j], atol=1e-4, rtol=1e-4), (layers, j)
            ll = om.forward(sd, cfg, parents, input_feats=feats, compute_point_est=False, shape_for_loglik=tgt['shape'],
                            pose_R_for_loglik=tgt['R'], glob_R_for_loglik=tgt['glob'])
        assert torch.allclose(torch.stack(ll['loglik_contexts'], 1), t('ctx_ll'), atol=2e-5, rtol=1e-5)
        lp_ref = t('lp_SO3')
        assert torch.equal(torch.isfinite(ll['pose_loglik']), torch.isfinite(lp_ref))
        fin = torch.isfinite(lp_ref)
        assert ((ll['pose_loglik'] - lp_ref).abs()[fin] / lp_ref.abs()[fin].clamp_min(1.0)).max().item() <= 1e-4
        with torch.no_grad():
            lp_alg = torch.stack([oflow.algebra_log_prob(om.joint_couplings(sd, j, 2), tgt['v_alg'][:, j], ll['loglik_contexts'][j],
                                                         RADIUS, 0.6) for j in range(23)], 1)
        assert ((lp_alg - t('lp_so3')).abs() / t('lp_so3').abs().clamp_min(1.0)).max().item() <= 1e-4


def test_sample_stats_against_the_real_tracker(golden_dir):
    """oracle/metrics.py::sample_stats vs the per-frame values of the REAL metrics/eval_metrics_tracker.py::EvalMetricsTracker
    (sample diversity of vertices / joints incl. the (in)visible-joint forms, samples-L2E against labels and against the input joints)."""
    from oracle import metrics as omet
    g = np.load(os.path.join(golden_dir, 'tracker_golden.npz'))
    close = lambda a, b: np.allclose(a, b, rtol=2e-6, atol=1e-7)
    assert close(omet.sample_stats(g['verts'])['diversity'], g['verts3D_sample_diversity'])
    assert close(omet.sample_stats(g['j3d'])['diversity'], g['joints3D_sample_diversity'])
    vis = g['in_vis'].astype(np.float64)
    assert close(omet.sample_stats(g['j3d'], weights=1.0 - vis)['diversity'], g['joints3D_invis_sample_diversity'])
    assert close(omet.sample_stats(g['j3d'], weights=vis)['diversity'], g['joints3D_vis_sample_diversity'])
    assert close(omet.sample_stats(g['j2d_samples'], g['tgt_j2d'], g['tgt_vis'])['l2e'], g['joints2Dsamples_L2E'])
    assert close(omet.sample_stats(g['j2d_samples'], g['in_j2d'], vis)['l2e'], g['input_joints2Dsamples_L2E'])
