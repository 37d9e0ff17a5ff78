"""Oracle (test infrastructure): HumaniflowModel.forward restated with INJECTED noise.

Follows /root/reference/models/humaniflow_model.py line by line (cited below).  The reference draws
its noise from the global torch generator and pyro consumes extra RNG per ``.condition`` call
(SURVEY F8), so parity is defined on explicit noise:
  shape_eps (B,N,10) ~ N(0,1)      -> shape_samples = mode + exp(log_std) * shape_eps   (:253-258)
  base_noise (B,N,23,3) ~ N(0,std^2) = the base-distribution draw of each joint's flow   (:309)
Weights come from a state dict with the reference's key names.
"""
from collections import defaultdict

import torch
import torch.nn.functional as F

from . import flow, so3
from .resnet import resnet_forward


def ancestors_of(parents):
    """humaniflow_model.py:16-30 (root excluded, nearest ancestor first)."""
    anc = defaultdict(list)
    for i in range(1, len(parents)):
        p = parents[i] - 1
        if p >= 0:
            anc[i - 1] += [p] + anc[p]
    return [anc[j] for j in range(len(parents) - 1)]


def joint_couplings(sd, j, num_transforms):
    """Layer lists of joint j's couplings: pose_so3flow_transform_modules.{num_transforms*j+t}.nn.layers.{l}
    (registration order humaniflow_model.py:111; layer naming [upstream] ConditionalDenseNN)."""
    out = []
    for t in range(num_transforms):
        p = 'pose_so3flow_transform_modules.%d.nn.layers.' % (num_transforms * j + t)
        n = len([k for k in sd if k.startswith(p) and k.endswith('.weight')])
        out.append([(sd[p + '%d.weight' % l], sd[p + '%d.bias' % l]) for l in range(n)])
    return out


def _lin(sd, name, x):
    return F.linear(x, sd[name + '.weight'], sd[name + '.bias'])


def image_level_feats(sd, input_feats, shape, glob_R, cam):
    """humaniflow_model.py:116-150."""
    if shape.dim() == 3:
        B, N = shape.shape[:2]
        cat = torch.cat([input_feats[:, None].expand(-1, N, -1), shape,
                         glob_R.reshape(B, 1, -1).expand(-1, N, -1), cam[:, None].expand(-1, N, -1)], dim=-1)
    else:
        cat = torch.cat([input_feats, shape, glob_R.reshape(shape.shape[0], -1), cam], dim=-1)
    return F.elu(_lin(sd, 'fc_input_shape_glob_cam_feats', cat))


def flow_context(sd, j, anc, feats, pose_SO3):
    """humaniflow_model.py:152-186: ancestors gathered in ancestor-list order, flattened row-major."""
    if len(anc) > 0:
        if pose_SO3.dim() == 5:
            B, N = pose_SO3.shape[:2]
            feats = torch.cat([feats, pose_SO3[:, :, anc].reshape(B, N, -1)], dim=-1)
        else:
            feats = torch.cat([feats, pose_SO3[:, anc].reshape(pose_SO3.shape[0], -1)], dim=-1)
    return F.elu(_lin(sd, 'fc_flow_context.%d' % j, feats))


def forward(sd, cfg, parents, input=None, input_feats=None, compute_point_est=True, num_samples=0,
            use_shape_mode_for_samples=False, shape_eps=None, base_noise=None,
            shape_for_loglik=None, pose_R_for_loglik=None, glob_R_for_loglik=None):
    """humaniflow_model.py:188-340.  Returns the reference's dict; for the log-likelihood mode the two
    lists of distributions are replaced by ``loglik_contexts`` (23 x (B,64)) and, when
    ``pose_R_for_loglik`` is given, ``pose_loglik`` (B,23) = each joint's SO(3) log_prob of its target."""
    nf = cfg.NORM_FLOW
    radius, std, T = nf.COMPACT_SUPPORT_RADIUS, nf.BASE_DIST_STD, nf.NUM_TRANSFORMS
    anc = ancestors_of(parents)
    nb = len(anc)
    if input_feats is None:
        input_feats = resnet_forward(sd, input, cfg.NUM_RESNET_LAYERS)            # :215-216
    B = input_feats.shape[0]
    x = F.elu(_lin(sd, 'fc1', input_feats))                                       # :232
    cam = _lin(sd, 'fc_cam', x) + sd['init_cam']                                  # :237-238
    glob_R = so3.rot6d_to_rotmat(_lin(sd, 'fc_glob', x) + sd['init_glob'])        # :243-245
    sp = _lin(sd, 'fc_shape', x)                                                  # :250-252
    nbeta = cfg.NUM_SMPL_BETAS
    shape_mode, shape_log_std = sp[:, :nbeta], sp[:, nbeta:]
    out = {'cam_wp': cam, 'glob_rotmat': glob_R, 'shape_mode': shape_mode, 'shape_log_std': shape_log_std,
           'input_feats': input_feats}
    loglik = pose_R_for_loglik is not None
    if num_samples > 0:
        if use_shape_mode_for_samples:
            shape_samples = shape_mode[:, None].expand(-1, num_samples, -1)       # :256
        else:
            shape_samples = shape_mode[:, None] + torch.exp(shape_log_std)[:, None] * shape_eps   # :258
        feats_s = image_level_feats(sd, input_feats, shape_samples, glob_R, cam)
        R_s = torch.zeros(B, num_samples, nb, 3, 3)                               # :272 (fp32 buffer)
    if compute_point_est:
        feats_pe = image_level_feats(sd, input_feats, shape_mode, glob_R, cam)
        v_pe = torch.zeros(B, nb, 3)
        R_pe = torch.zeros(B, nb, 3, 3)
    if loglik:
        feats_ll = image_level_feats(sd, input_feats, shape_for_loglik, glob_R_for_loglik, cam)  # :280-283
        ll_ctx, ll = [], torch.zeros(B, nb)
    for j in range(nb):                                                           # :286
        cps = joint_couplings(sd, j, T)
        if compute_point_est:                                                     # :290-301
            ctx = flow_context(sd, j, anc[j], feats_pe, R_pe)
            v = flow.flow_forward(cps, torch.zeros(B, 3), ctx, radius)
            v_pe[:, j] = v
            R_pe[:, j] = so3.batch_rodrigues(v)
        if num_samples > 0:                                                       # :304-311
            ctx = flow_context(sd, j, anc[j], feats_s, R_s)
            R_s[:, :, j] = flow.so3_sample(cps, base_noise[:, :, j], ctx, radius)  # f64 -> f32 store
        if loglik:                                                                # :314-320
            ctx = flow_context(sd, j, anc[j], feats_ll, pose_R_for_loglik)
            ll_ctx.append(ctx)
            ll[:, j] = flow.so3_log_prob(cps, pose_R_for_loglik[:, j].double(), ctx, radius, std)
    if compute_point_est:
        out['pose_axisangle_point_est'] = v_pe
        out['pose_rotmats_point_est'] = R_pe
    if num_samples > 0:
        out['pose_rotmats_samples'] = R_s
        out['shape_samples'] = shape_samples
    if loglik:
        out['loglik_contexts'] = ll_ctx
        out['pose_loglik'] = ll
    return out
