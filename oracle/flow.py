"""Oracle (test infrastructure): per-joint conditional flow on so(3) and its SO(3) push-forward.

Reference-owned structure followed literally (PINNED where noted in oracle/__init__.py):
  models/norm_flows/pyro_conditional_norm_flow.py:21-129   transform list & forward push
  models/norm_flows/transforms/scaled_radial_tanh_transform.py:28-59
  models/norm_flows/transforms/so3_exp_transform.py:30-50, to_transform.py:22-29
  models/norm_flows/local_diffeo_transformed_distribution.py:72-142   rsample / log_prob
Upstream arithmetic (pyro spline coupling, Permute) is in oracle/spline.py ([upstream], unpinned).

A flow for one joint is described by ``couplings``: a list (len NUM_TRANSFORMS) of layer lists
[(W,b) x4] in the order the reference registers them (humaniflow_model.py:111).
"""
import math

import torch

from . import so3
from .spline import coupling_forward, coupling_inverse


def permutations(event_dim, num_transforms):
    """pyro_conditional_norm_flow.py:46-47,60-62: cycle of rotations idx[i:]+idx[:i]."""
    idx = list(range(event_dim))
    return [idx[i % event_dim:] + idx[:i % event_dim] for i in range(num_transforms)]


def radial_tanh_forward(x, radius):
    """scaled_radial_tanh_transform.py:32-39 (input dtype, fp32 on the hot path)."""
    n = x.norm(dim=-1, keepdim=True)
    mask = n > 1e-7
    n = torch.where(mask, n, torch.ones_like(n))
    return torch.where(mask, torch.tanh(n / radius) * (x / n) * radius, x)


def radial_tanh_inverse(y, radius):
    """scaled_radial_tanh_transform.py:41-50: computed in float64, cast back to the input dtype."""
    dt = y.dtype
    y = y.double()
    n = y.norm(dim=-1, keepdim=True)
    mask = n > 1e-7
    n = torch.where(mask, n, torch.ones_like(n))
    return torch.where(mask, torch.atanh(n / radius) * (y / n) * radius, y).to(dt)


def radial_tanh_log_abs_det(x, y, radius):
    """scaled_radial_tanh_transform.py:52-59."""
    xn, yn = x.norm(dim=-1), y.norm(dim=-1)
    ld = 2 * (torch.log(yn) - torch.log(xn)) + torch.log1p(-((yn / radius) ** 2))
    return torch.where(yn > 1e-7, ld, torch.zeros_like(ld))


def flow_forward(couplings, z, context, radius, with_logdet=False):
    """pyro_conditional_norm_flow.py:120-129 / torch TransformedDistribution.rsample:
    z -> [Permute_i -> SplineCoupling_i]* -> ScaledRadialTanh.  z, context fp32."""
    x = z
    logdet = torch.zeros(z.shape[:-1], dtype=z.dtype)
    for perm, layers in zip(permutations(3, len(couplings)), couplings):
        x = x[..., perm]                                   # [upstream] pyro Permute: y = x.index_select(-1, perm)
        x, ld = coupling_forward(layers, x, context, radius)
        logdet = logdet + ld
    y = radial_tanh_forward(x, radius)
    if with_logdet:
        return y, logdet + radial_tanh_log_abs_det(x, y, radius)
    return y


def normal_log_prob(x, std):
    """torch.distributions.Normal.log_prob summed over the event dim (Independent(...,1));
    pyro_conditional_norm_flow.py:49-52."""
    scale = torch.full((), std, dtype=x.dtype)
    lp = -(x ** 2) / (2 * scale ** 2) - scale.log() - math.log(math.sqrt(2 * math.pi))
    return lp.sum(-1)


def algebra_log_prob(couplings, v, context, radius, base_std):
    """log-density on so(3) ~ R^3 of the conditioned flow (torch TransformedDistribution.log_prob over
    the transform list of pyro_conditional_norm_flow.py:54-112).  v fp32 (..., 3)."""
    y = v
    x = radial_tanh_inverse(y, radius)
    lp = 0.0 - radial_tanh_log_abs_det(x, y, radius)
    perms = permutations(3, len(couplings))
    for perm, layers in zip(reversed(perms), reversed(couplings)):
        x, ld = coupling_inverse(layers, x, context, radius)
        lp = lp - ld
        inv = [perm.index(i) for i in range(3)]            # [upstream] Permute._inverse
        x = x[..., inv]
    return lp + normal_log_prob(x, base_std)


def so3_sample(couplings, z, context, radius):
    """local_diffeo_transformed_distribution.py:72-82 with transforms [ToTransform(f32->f64), SO3Exp]
    (humaniflow_model.py:107-109): returns float64 rotation matrices."""
    v = flow_forward(couplings, z, context, radius)
    return so3.so3_exp(v.double())


def so3_log_prob(couplings, R, context, radius, base_std):
    """local_diffeo_transformed_distribution.py:84-142 specialised to the two transforms the model
    uses.  R float64 (B,3,3) -> float32 (B,).  Terms: principal log x and the two pre-images
    x/|x|(|x| -+ 2pi) masked to |.| < support radius (so3_exp_transform.py:33-41), each
    -log|det J_exp| (f64 -> f32) + algebra log-density (f32), combined by logsumexp."""
    x = so3.so3_log(R)                                      # f64
    xset = so3.so3_xset(x)                                  # (2,B,3) f64, NaN when |x|==0
    mask = xset.norm(dim=-1) < radius
    xset = xset.masked_fill(~mask[..., None], 0.0)

    def term(xx):
        ld = so3.so3_log_abs_det_jacobian(xx).float()       # so3_exp_transform.py:43-50
        nxt = torch.zeros(xx.shape[:-1]) + algebra_log_prob(couplings, xx.float(), context, radius, base_std)
        return (-ld) + nxt

    x_term = term(x)
    set_terms = torch.where(mask, term(xset), torch.tensor([float('-inf')]))
    return torch.logsumexp(torch.cat([x_term[None], set_terms]), dim=0)
