"""Generate golden vectors by running the REAL reference files found under /root/reference.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Writes small .npz fixtures next to this file.  What is exercised is reference-owned code only:

  so3_golden.npz     utils/rigid_transform_utils.py  (rot6d_to_rotmat, rotmat_to_rot6d, so3_exp, so3_log incl.
                     the near-pi branch, so3_xset, so3_log_abs_det_jacobian)
  rtanh_golden.npz   models/norm_flows/transforms/scaled_radial_tanh_transform.py
  resnet_golden.npz  models/resnet.py  resnet18 / resnet50 (eval), deterministic weights from detweights.py
  logprob_golden.npz models/norm_flows/local_diffeo_transformed_distribution.py + so3_exp_transform.py +
                     to_transform.py + local_diffeo_transform.py: the real LocalDiffeoTransformedDistribution
                     (rsample and log_prob with pre-image logsumexp) over a torch TransformedDistribution whose
                     spline-coupling transforms are THIS repo's restatement (oracle/spline.py) -- pyro is not
                     installable here, so the spline arithmetic itself stays unpinned; what this pins is the
                     reference's SO(3) plumbing, masks, dtypes and term order.
  regressors_sparse.npz  the three real joint regressors shipped in model_files/*.npy as COO triplets.

The pyro names the reference files subclass (`ConditionalDistribution`, `ConstantConditionalDistribution`) are
provided as empty name-only classes below; they contain no arithmetic.
"""
import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from detweights import fill_state_dict, det_input  # noqa: E402


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def pkg(name):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules[name] = m
    return m


def so3_cases():
    rs = np.random.RandomState(7)
    v = rs.standard_normal((64, 3))
    v = v / np.linalg.norm(v, axis=1, keepdims=True)
    ang = np.concatenate([rs.uniform(0, math.pi, 40), [0.0, 1e-12, 1e-9, 1e-6, math.pi / 2, math.pi - 5e-3,
                          math.pi - 1e-3, math.pi - 1e-5, math.pi - 1e-8, math.pi, 3.0, 3.1, 2.0, 1.0],
                          rs.uniform(math.pi - 1e-2, math.pi, 10)])
    return torch.tensor(v * ang[:, None], dtype=torch.float64)


def main():
    sys.path.insert(0, REF)           # for the real `utils` package (rigid_transform_utils, lin_alg_utils)
    rtu = load('utils/rigid_transform_utils.py', 'utils.rigid_transform_utils')

    # ---- so3 ----
    v = so3_cases()
    R = rtu.so3_exp(v)
    logv = rtu.so3_vee(rtu.so3_log(R.clone()))
    xset = rtu.so3_xset(logv, 1)
    lad = rtu.so3_log_abs_det_jacobian(v)
    lad32 = rtu.so3_log_abs_det_jacobian(v.float())
    r6 = det_input((32, 6), 11)
    R6 = rtu.rot6d_to_rotmat(r6.clone())
    back6 = rtu.rotmat_to_rot6d(R6)
    np.savez(os.path.join(HERE, 'so3_golden.npz'), v=v.numpy(), R=R.numpy(), logv=logv.numpy(),
             xset=xset.numpy(), lad=lad.numpy(), lad32=lad32.numpy(), r6=r6.numpy(), R6=R6.numpy(),
             back6=back6.numpy())

    # ---- scaled radial tanh ----
    srt = load('models/norm_flows/transforms/scaled_radial_tanh_transform.py', '_ref_srt')
    radius = 1.5 * math.pi
    t = srt.ScaledRadialTanhTransform(radius=radius)
    x = det_input((128, 3), 13) * torch.tensor(np.random.RandomState(14).uniform(0, 4, (128, 1)).astype(np.float32))
    x[0] = 0.0
    x[1] = 1e-8
    y = t._call(x)
    xi = t._inverse(y)
    ld = t.log_abs_det_jacobian(x, y)
    np.savez(os.path.join(HERE, 'rtanh_golden.npz'), x=x.numpy(), y=y.numpy(), xinv=xi.numpy(), ld=ld.numpy())

    # ---- resnet ----
    resnet = load('models/resnet.py', '_ref_resnet')
    res = {}
    for layers, ctor in ((18, resnet.resnet18), (50, resnet.resnet50)):
        net = ctor(in_channels=18, pretrained=False).eval()
        shapes = {k: v.shape for k, v in net.state_dict().items()}
        sd = fill_state_dict(shapes, seed=100 + layers)
        net.load_state_dict(sd, strict=True)
        inp = det_input((2, 18, 64, 64), 200 + layers, kind='uniform')
        with torch.no_grad():
            res['feats%d' % layers] = net(inp).numpy()
    np.savez(os.path.join(HERE, 'resnet_golden.npz'), **res)

    # ---- real LocalDiffeoTransformedDistribution over an oracle-spline base ----
    pyro = pkg('pyro'); pd = pkg('pyro.distributions'); pc = types.ModuleType('pyro.distributions.conditional')

    class ConditionalDistribution:                       # name-only stand-in, no arithmetic
        pass

    class ConstantConditionalDistribution(ConditionalDistribution):
        def __init__(self, base_dist):
            self.base_dist = base_dist

        def condition(self, context):
            return self.base_dist
    pc.ConditionalDistribution = ConditionalDistribution
    pc.ConstantConditionalDistribution = ConstantConditionalDistribution
    sys.modules['pyro.distributions.conditional'] = pc
    pkg('models'); pkg('models.norm_flows'); tr = pkg('models.norm_flows.transforms')
    ldt = load('models/norm_flows/transforms/local_diffeo_transform.py', 'models.norm_flows.transforms.local_diffeo_transform')
    tr.LocalDiffeoTransform = ldt.LocalDiffeoTransform
    tot = load('models/norm_flows/transforms/to_transform.py', 'models.norm_flows.transforms.to_transform')
    s3t = load('models/norm_flows/transforms/so3_exp_transform.py', 'models.norm_flows.transforms.so3_exp_transform')
    ldd = load('models/norm_flows/local_diffeo_transformed_distribution.py',
               'models.norm_flows.local_diffeo_transformed_distribution')

    from torch.distributions import Transform, constraints, Normal, Independent, TransformedDistribution
    from oracle import spline as osp, flow as oflow

    class OracleCoupling(Transform):                     # oracle spline dressed as a torch Transform
        domain = constraints.real_vector
        codomain = constraints.real_vector
        bijective = True

        def __init__(self, layers, ctx, bound):
            super().__init__(cache_size=1)
            self.layers, self.ctx, self.bound = layers, ctx, bound

        def _call(self, x):
            y, self._ld = osp.coupling_forward(self.layers, x, self.ctx, self.bound)
            return y

        def _inverse(self, y):
            x, self._ld = osp.coupling_inverse(self.layers, y, self.ctx, self.bound)
            return x

        def log_abs_det_jacobian(self, x, y):
            return self._ld

    class Perm(Transform):
        domain = constraints.real_vector
        codomain = constraints.real_vector
        bijective = True

        def __init__(self, perm):
            super().__init__(cache_size=1)
            self.perm = perm
            self.inv_perm = [perm.index(i) for i in range(len(perm))]

        def _call(self, x):
            return x[..., self.perm]

        def _inverse(self, y):
            return y[..., self.inv_perm]

        def log_abs_det_jacobian(self, x, y):
            return torch.zeros(x.shape[:-1], dtype=x.dtype)

    B = 48
    dims = [(65, 64), (64, 32), (32, 32), (32, 62)]
    shapes = {}
    for t_i in range(2):
        for l, (i, o) in enumerate(dims):
            shapes['c%d.%d.weight' % (t_i, l)] = (o, i)
            shapes['c%d.%d.bias' % (t_i, l)] = (o,)
    sd = fill_state_dict(shapes, seed=321)
    for k in sd:                                         # keep the fp32 splines well-conditioned (see tests/test_oracle_invariants.py)
        sd[k] = sd[k] * 0.5
    couplings = [[(sd['c%d.%d.weight' % (t_i, l)], sd['c%d.%d.bias' % (t_i, l)]) for l in range(4)] for t_i in range(2)]
    ctx = det_input((B, 64), 322)
    perms = oflow.permutations(3, 2)
    base = Independent(Normal(torch.zeros(3), torch.ones(3) * 0.6, validate_args=False), 1)
    transforms = []
    for p, c in zip(perms, couplings):
        transforms += [Perm(p), OracleCoupling(c, ctx, radius)]
    transforms.append(srt.ScaledRadialTanhTransform(radius=radius))
    so3flow = TransformedDistribution(base, transforms, validate_args=False)
    dist = ldd.ConditionalLocalDiffeoTransformedDistribution(
        base_dist=so3flow,
        transforms=[tot.ToTransform(dict(dtype=torch.float32), dict(dtype=torch.float64)),
                    s3t.SO3ExpCompactTransform(support_radius=radius)]).condition(ctx)
    # targets: random rotations incl. small angle, > pi/2 (second pre-image inside the support) and near pi
    vt = so3_cases()[:B].clone()
    Rt = rtu.so3_exp(vt)
    with torch.no_grad():
        lp = dist.log_prob(Rt.clone())
        lp_alg = so3flow.log_prob(vt.float() * 0.9)
    # sampling path with injected base noise: replay rsample's transform loop on a fixed z
    z = det_input((B, 3), 323) * 0.6
    with torch.no_grad():
        xs = z
        for tfm in so3flow.transforms:
            xs = tfm(xs)
        v_alg = xs
        for tfm in dist.transforms:
            xs = tfm(xs)
        lp_s = dist.log_prob(xs.clone())
    np.savez(os.path.join(HERE, 'logprob_golden.npz'), Rt=Rt.numpy(), lp=lp.numpy(), vt=vt.numpy(),
             lp_alg=lp_alg.numpy(), z=z.numpy(), v_alg=v_alg.numpy(), R_s=xs.numpy(), lp_s=lp_s.numpy())

    # ---- the three real regressors as sparse triplets ----
    trip = {}
    for key, fn in (('extra', 'J_regressor_extra.npy'), ('cocoplus', 'cocoplus_regressor.npy'),
                    ('h36m', 'J_regressor_h36m.npy')):
        a = np.load(os.path.join(REF, 'model_files', fn))
        r, c = np.nonzero(a)
        trip[key + '_shape'] = np.array(a.shape)
        trip[key + '_row'] = r.astype(np.int32)
        trip[key + '_col'] = c.astype(np.int32)
        trip[key + '_val'] = a[r, c].astype(np.float64)
    np.savez(os.path.join(HERE, 'regressors_sparse.npz'), **trip)
    print('golden vectors written to', HERE)


if __name__ == '__main__':
    main()
