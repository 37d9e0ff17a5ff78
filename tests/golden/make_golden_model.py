"""Golden vectors of the REAL reference class models/humaniflow_model.py::HumaniflowModel.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden_model.py
Writes tests/golden/model_golden.npz.

What runs: the reference's own, unmodified ``HumaniflowModel.__init__`` / ``forward`` (models/humaniflow_model.py:33-340),
``create_conditional_norm_flow`` / ``forward_trans_conditional_norm_flow`` (models/norm_flows/pyro_conditional_norm_flow.py),
``ConditionalSplineCoupling.condition`` (transforms/conditional_spline_coupling_transform.py:35-48), the real
``ConditionalLocalDiffeoTransformedDistribution`` / ``LocalDiffeoTransformedDistribution`` (rsample and log_prob),
``ScaledRadialTanhTransform``, ``ToTransform``, ``SO3ExpCompactTransform``, ``utils/rigid_transform_utils.py``.
This pins oracle/model.py's glue: head order, ancestor order, ``2j+t`` module index, context-first concat site, the (B,N)
broadcast of the image-level features, dtype hand-offs, what the loglik lists are conditioned on.

pyro-ppl 1.7.0 and smplx 0.1.26 are not installable here (no network, not in the wheelhouse), so the names the reference
imports from them are provided below as stand-ins: containers and plumbing with the upstream call signatures, whose
ARITHMETIC is delegated to this repo's restatement (oracle/spline.py, oracle/so3.py::batch_rodrigues).  The spline /
DenseNN / Rodrigues arithmetic itself therefore stays unpinned (stated in DESIGN.md 3); everything around it is the
reference's own code.

Noise: the reference draws from torch's global generator (F8).  ``torch.distributions.normal._standard_normal`` is
replaced for the duration of the call by a queue of pre-drawn tensors, so the real ``rsample`` paths run unchanged
and the draws are known: first ``shape_dist.rsample([N])`` (eps of shape (N,B,10), transposed by the reference to (B,N,10)),
then one (B,N,3) base draw per joint in joint order (scaled by BASE_DIST_STD by the Normal itself).
"""
import os
import sys
import types
from functools import partial

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from detweights import det_input, fill_state_dict  # noqa: E402


def install_stand_ins():
    """Modules named like the upstream packages; arithmetic delegated to oracle/."""
    from torch.distributions import Transform, TransformedDistribution, constraints
    import torch.distributions as td
    from oracle import so3 as oso3
    from oracle import spline as osp

    def mod(name, is_pkg=True):
        m = types.ModuleType(name)
        if is_pkg:
            m.__path__ = []
        sys.modules[name] = m
        return m

    pyro = mod('pyro')
    pd = mod('pyro.distributions')
    pt = mod('pyro.distributions.transforms')
    psc = mod('pyro.distributions.transforms.spline_coupling', False)
    pc = mod('pyro.distributions.conditional', False)
    ptt = mod('pyro.distributions.torch_transform', False)
    pnn = mod('pyro.nn')
    pyro.distributions, pyro.nn = pd, pnn
    pd.transforms, pd.conditional, pd.torch_transform = pt, pc, ptt

    # ---- pyro.distributions.conditional / torch_transform: plumbing only
    class ConditionalDistribution:
        def condition(self, context):
            raise NotImplementedError

    class ConstantConditionalDistribution(ConditionalDistribution):
        def __init__(self, base_dist):
            self.base_dist = base_dist

        def condition(self, context):
            return self.base_dist

    class ConditionalTransform:
        def condition(self, context):
            raise NotImplementedError

    class ConstantConditionalTransform(ConditionalTransform):
        def __init__(self, transform):
            self.transform = transform

        def condition(self, context):
            return self.transform

    class ConditionalTransformModule(ConditionalTransform, nn.Module):
        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)

        def __hash__(self):
            return nn.Module.__hash__(self)

    class TransformModule(Transform, nn.Module):
        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)

        def __hash__(self):
            return nn.Module.__hash__(self)

    class ConditionalTransformedDistribution(ConditionalDistribution):
        def __init__(self, base_dist, transforms):
            self.base_dist = base_dist if isinstance(base_dist, ConditionalDistribution) else ConstantConditionalDistribution(base_dist)
            self.transforms = [t if isinstance(t, ConditionalTransform) else ConstantConditionalTransform(t) for t in transforms]

        def condition(self, context):
            return TransformedDistribution(self.base_dist.condition(context), [t.condition(context) for t in self.transforms],
                                           validate_args=False)

        def clear_cache(self):
            pass

    pc.ConditionalDistribution = ConditionalDistribution
    pc.ConstantConditionalDistribution = ConstantConditionalDistribution
    pc.ConditionalTransform = ConditionalTransform
    pc.ConditionalTransformModule = ConditionalTransformModule
    ptt.TransformModule = TransformModule
    pd.ConditionalTransformedDistribution = ConditionalTransformedDistribution
    pd.ConditionalTransform = ConditionalTransform
    pd.ConditionalDistribution = ConditionalDistribution
    pd.Normal, pd.Independent, pd.constraints = td.Normal, td.Independent, constraints

    # ---- pyro.nn: parameter containers with the upstream layer naming (.layers.{l}); forward is never used, the
    # stand-in SplineCoupling below hands the layers to oracle.spline.dense_nn
    class ConditionalDenseNN(nn.Module):
        def __init__(self, input_dim, context_dim, hidden_dims, param_dims=(1, 1), nonlinearity=None):
            super().__init__()
            dims = [input_dim + context_dim] + list(hidden_dims) + [sum(param_dims)]
            self.layers = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)])
            self.param_dims = list(param_dims)

    class DenseNN(ConditionalDenseNN):
        def __init__(self, input_dim, hidden_dims, param_dims=(1, 1), nonlinearity=None):
            super().__init__(input_dim, 0, hidden_dims, param_dims)

    pnn.ConditionalDenseNN, pnn.DenseNN = ConditionalDenseNN, DenseNN

    # ---- pyro.distributions.transforms
    class Permute(Transform):
        domain = constraints.real_vector
        codomain = constraints.real_vector
        bijective = True

        def __init__(self, permutation, *, dim=-1, cache_size=1):
            super().__init__(cache_size=cache_size)
            self.permutation = permutation
            self.inv_permutation = torch.argsort(permutation)

        def _call(self, x):
            return x.index_select(-1, self.permutation)

        def _inverse(self, y):
            return y.index_select(-1, self.inv_permutation)

        def log_abs_det_jacobian(self, x, y):
            return torch.zeros(x.size()[:-1], dtype=x.dtype, device=x.device)

    class SplineCoupling(Transform):
        """Signature of pyro.distributions.transforms.SplineCoupling as conditional_spline_coupling_transform.py:42-48
        calls it; arithmetic = oracle/spline.py (coupling_forward / coupling_inverse)."""
        domain = constraints.real_vector
        codomain = constraints.real_vector
        bijective = True

        def __init__(self, input_dim, split_dim, hypernet, count_bins=8, bound=3., order='linear', identity=False):
            super().__init__(cache_size=1)
            assert isinstance(hypernet, partial) and order == 'linear' and identity and split_dim == 1 and input_dim == 3
            self.layers = [(l.weight, l.bias) for l in hypernet.func.layers]
            self.context = hypernet.keywords['context']
            self.bound = bound
            assert count_bins == 8

        def _call(self, x):
            y, self._ld = osp.coupling_forward(self.layers, x, self.context, self.bound)
            return y

        def _inverse(self, y):
            x, self._ld = osp.coupling_inverse(self.layers, y, self.context, self.bound)
            return x

        def log_abs_det_jacobian(self, x, y):
            return self._ld

    pt.Permute = Permute
    pt.SplineCoupling = SplineCoupling
    psc.SplineCoupling = SplineCoupling
    pt.spline_coupling = psc

    # ---- smplx.lbs.batch_rodrigues (humaniflow_model.py:6,299)
    smplx = mod('smplx')
    lbs = mod('smplx.lbs', False)
    lbs.batch_rodrigues = oso3.batch_rodrigues
    smplx.lbs = lbs


class NoiseQueue:
    """Replaces torch.distributions.normal._standard_normal by a queue of pre-drawn tensors."""

    def __init__(self, tensors):
        self.q = list(tensors)

    def __enter__(self):
        import torch.distributions.normal as tn
        self._tn, self._orig = tn, tn._standard_normal

        def pop(shape, dtype, device):
            t = self.q.pop(0)
            assert tuple(t.shape) == tuple(shape), (t.shape, shape)
            return t.to(dtype)
        tn._standard_normal = pop
        return self

    def __exit__(self, *a):
        self._tn._standard_normal = self._orig
        assert not self.q, 'unused noise'


def main():
    install_stand_ins()
    sys.path.insert(0, REF)
    import humaniflow_b200 as hb
    from humaniflow_b200.synthetic import SMPL_PARENTS
    from models.humaniflow_model import HumaniflowModel        # the REAL reference class
    from oracle import so3 as oso3

    out = {}
    for layers, feat_dim in ((18, 512), (50, 2048)):
        cfg = hb.get_model_cfg_defaults()
        cfg.NUM_RESNET_LAYERS = layers
        torch.manual_seed(0)
        model = HumaniflowModel('cpu', cfg, list(SMPL_PARENTS)).eval()
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if not k.startswith('image_encoder.')}
        sd = fill_state_dict(shapes, seed=700 + layers)
        for k in sd:            # N(0, 2/fan_in) is ~2.4x torch's default Linear init; halve it so the 23-joint chain is as
            if sd[k].dim() >= 2:   # well-conditioned as a freshly initialised model (fp32 summation-order noise stays ~1e-6)
                sd[k] = sd[k] * 0.5
        sd['init_glob'] = model.state_dict()['init_glob'].clone()
        sd['init_cam'] = model.state_dict()['init_cam'].clone()
        missing = model.load_state_dict(sd, strict=False)
        assert all(k.startswith('image_encoder.') for k in missing.missing_keys) and not missing.unexpected_keys
        tag = 'r%d_' % layers
        out[tag + 'keys'] = np.array(sorted(shapes.keys()))
        out[tag + 'anc_len'] = np.array([len(model.ancestors_dict[j]) for j in range(23)])
        out[tag + 'anc_flat'] = np.array([a for j in range(23) for a in model.ancestors_dict[j]])

        B, N = 5, 4      # not 3: rigid_transform_utils.py:99 calls torch.cross without dim, which for a (3,3) input picks dim 0
        feats = det_input((B, feat_dim), 710 + layers).abs()
        eps_shape = det_input((N, B, 10), 711 + layers)                       # as rsample([N]) draws it
        eps_pose = [det_input((B, N, 3), 720 + layers + j) for j in range(23)]
        ctx_log = []
        orig_ctx = model.compute_flow_context

        def rec_ctx(*a, **k):
            c = orig_ctx(*a, **k)
            ctx_log.append(c.detach().clone())
            return c
        model.compute_flow_context = rec_ctx
        with torch.no_grad(), NoiseQueue([eps_shape] + eps_pose):
            o = model(None, input_feats=feats, compute_point_est=True, num_samples=N)
        # per joint the forward records [point-estimate context (B,64), samples context (B,N,64)]
        out[tag + 'ctx_pe'] = torch.stack(ctx_log[0::2], 1).numpy()
        out[tag + 'ctx_s'] = torch.stack(ctx_log[1::2], 2).numpy()
        for k in ('cam_wp', 'glob_rotmat', 'shape_mode', 'shape_log_std', 'shape_samples', 'pose_axisangle_point_est',
                  'pose_rotmats_point_est', 'pose_rotmats_samples'):
            out[tag + k] = o[k].numpy()
        assert o['pose_rotmats_samples'].dtype == torch.float32 and o['pose_rotmats_samples'].shape == (B, N, 23, 3, 3)
        out[tag + 'scale_of_shape_dist'] = o['shape_dist_for_loglik'].scale.numpy()

        # log-likelihood mode: teacher forcing on given targets (:277-283, :314-320)
        v_t = det_input((B, 23, 3), 760 + layers) * 0.7
        v_t[0, 3] = 0.0                                                       # identity target
        v_t[1, 5] = v_t[1, 5] / v_t[1, 5].norm() * 2.6                        # second pre-image inside the support
        R_t = oso3.so3_exp(v_t.double()).float()
        shape_t = det_input((B, 10), 761 + layers)
        glob_t = oso3.so3_exp(det_input((B, 3), 762 + layers).double() * 0.5).float()
        ctx_log.clear()
        with torch.no_grad():
            o = model(None, input_feats=feats, compute_point_est=False, num_samples=0, compute_for_loglik=True,
                      shape_for_loglik=shape_t, pose_R_for_loglik=R_t, glob_R_for_loglik=glob_t)
            lp_SO3 = torch.stack([o['conditioned_pose_SO3flow_dists_for_loglik'][j].log_prob(R_t[:, j].double()) for j in range(23)], 1)
            v_alg = det_input((B, 23, 3), 763 + layers) * 0.8
            lp_so3 = torch.stack([o['conditioned_pose_so3flow_dists_for_loglik'][j].log_prob(v_alg[:, j]) for j in range(23)], 1)
        assert lp_SO3.dtype == torch.float32 and lp_SO3.shape == (B, 23)
        out[tag + 'ctx_ll'] = torch.stack(ctx_log, 1).numpy()
        out[tag + 'lp_SO3'] = lp_SO3.numpy()
        out[tag + 'lp_so3'] = lp_so3.numpy()
        model.compute_flow_context = orig_ctx
    np.savez_compressed(os.path.join(HERE, 'model_golden.npz'), **out)
    print('model_golden.npz written:', {k: v.shape for k, v in out.items() if k.startswith('r18_')})


if __name__ == '__main__':
    main()
