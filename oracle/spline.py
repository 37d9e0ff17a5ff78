"""Oracle (test infrastructure): conditional rational-LINEAR spline coupling on R^3.

[upstream, PARITY UNPINNED]  The arithmetic restated here lives in pyro-ppl==1.7.0
(requirements.txt:11), which is not vendored in /root/reference and not installable here:
``pyro.distributions.transforms.spline._monotonic_rational_spline`` (order='linear',
Dolatabadi et al. 2020), ``spline.ConditionalSpline._params``, ``spline_coupling.SplineCoupling``
(identity=True) and ``pyro.nn.dense_nn.ConditionalDenseNN``.  The reference's own call sites that fix
the configuration are cited per function (models/norm_flows/transforms/
conditional_spline_coupling_transform.py:35-78, pyro_conditional_norm_flow.py:46-62).
Constants [upstream]: min_bin_width = min_bin_height = min_derivative = 1e-3, min_lambda = 0.025,
searchsorted eps = 1e-6, boundary derivatives 1 - min_derivative, identity outside [-bound, bound].
"""
import torch
import torch.nn.functional as F

MIN_BIN = 1e-3
MIN_DERIV = 1e-3
MIN_LAMBDA = 0.025
SEARCH_EPS = 1e-6


def dense_nn(layers, x1, context):
    """[upstream] ConditionalDenseNN.forward: input = cat[context, x1] (context FIRST), ReLU between
    layers, last layer linear; output split into param_dims [16,16,14,16] for 3-D events
    (conditional_spline_coupling_transform.py:64-70: split_dim=1, count_bins=8).
    layers: list of (weight, bias)."""
    context = context.expand(x1.shape[:-1] + (context.shape[-1],))   # [upstream] broadcast context over x
    h = torch.cat([context, x1], dim=-1)
    for w, b in layers[:-1]:
        h = F.relu(F.linear(h, w, b))
    w, b = layers[-1]
    return F.linear(h, w, b)


def spline_params(raw, n_dims=2, bins=8):
    """[upstream] ConditionalSpline._params for DenseNN outputs: reshape each block to
    (..., n_dims, bins) (derivatives: bins-1), softmax / softmax / softplus / sigmoid."""
    nb = n_dims * bins
    w, h, d, lam = raw[..., :nb], raw[..., nb:2 * nb], raw[..., 2 * nb:3 * nb - n_dims], raw[..., 3 * nb - n_dims:]
    w = F.softmax(w.reshape(*w.shape[:-1], n_dims, bins), dim=-1)
    h = F.softmax(h.reshape(*h.shape[:-1], n_dims, bins), dim=-1)
    d = F.softplus(d.reshape(*d.shape[:-1], n_dims, bins - 1))
    lam = torch.sigmoid(lam.reshape(*lam.shape[:-1], n_dims, bins))
    return w, h, d, lam


def _knots(lengths, lo, hi):
    """[upstream] spline._calculate_knots: cumsum, left-pad 0, affine to [lo,hi], pin both ends,
    re-derive the lengths from the pinned knots."""
    knots = F.pad(torch.cumsum(lengths, dim=-1), (1, 0), value=0.0)
    knots = (hi - lo) * knots + lo
    knots[..., 0] = lo
    knots[..., -1] = hi
    return knots[..., 1:] - knots[..., :-1], knots


def _pick(x, idx):
    """[upstream] spline._select_bins: clamp the bin index into range, gather on the last dim."""
    idx = idx.clamp(min=0, max=x.size(-1) - 1)
    return x.gather(-1, idx).squeeze(-1)


def rational_linear_spline(inputs, widths, heights, derivs, lambdas, bound, inverse=False):
    """[upstream] spline._monotonic_rational_spline with lambdas (order='linear').
    inputs (..., D); widths/heights/lambdas (..., D, K); derivs (..., D, K-1).
    Returns (outputs, logabsdet) with logabsdet = log|d out / d in| of THIS direction."""
    K = widths.shape[-1]
    lo, hi = -bound, bound
    inside = (inputs >= lo) & (inputs <= hi)

    widths = MIN_BIN + (1.0 - MIN_BIN * K) * widths
    heights = MIN_BIN + (1.0 - MIN_BIN * K) * heights
    derivs = MIN_DERIV + derivs
    widths, cumw = _knots(widths, lo, hi)
    heights, cumh = _knots(heights, lo, hi)
    derivs = F.pad(derivs, (1, 1), value=1.0 - MIN_DERIV)

    edges = (cumh if inverse else cumw) + SEARCH_EPS
    bin_idx = (torch.sum(inputs[..., None] >= edges, dim=-1) - 1)[..., None]

    in_w = _pick(widths, bin_idx)
    in_cw = _pick(cumw, bin_idx)
    in_ch = _pick(cumh, bin_idx)
    in_delta = _pick(heights / widths, bin_idx)
    in_d = _pick(derivs, bin_idx)
    in_d1 = _pick(derivs[..., 1:], bin_idx)
    in_h = _pick(heights, bin_idx)
    lam = _pick((1 - 2 * MIN_LAMBDA) * lambdas + MIN_LAMBDA, bin_idx)

    wa = 1.0
    wb = torch.sqrt(in_d / in_d1) * wa
    wc = (lam * wa * in_d + (1 - lam) * wb * in_d1) / in_delta
    ya = in_ch
    yb = in_h + in_ch
    yc = ((1 - lam) * wa * ya + lam * wb * yb) / ((1 - lam) * wa + lam * wb)

    if inverse:
        left = (inputs <= yc).to(inputs.dtype)
        right = (inputs > yc).to(inputs.dtype)
        num = (lam * wa * (ya - inputs)) * left \
            + ((wc - lam * wb) * inputs + lam * wb * yb - wc * yc) * right
        den = ((wc - wa) * inputs + wa * ya - wc * yc) * left \
            + ((wc - wb) * inputs + wb * yb - wc * yc) * right
        theta = num / den
        out = theta * in_w + in_cw
        dnum = (wa * wc * lam * (yc - ya) * left + wb * wc * (1 - lam) * (yb - yc) * right) * in_w
        lad = torch.log(dnum) - 2 * torch.log(torch.abs(den))
    else:
        theta = (inputs - in_cw) / in_w
        left = (theta <= lam).to(inputs.dtype)
        right = (theta > lam).to(inputs.dtype)
        num = (wa * ya * (lam - theta) + wc * yc * theta) * left \
            + (wc * yc * (1 - theta) + wb * yb * (theta - lam)) * right
        den = (wa * (lam - theta) + wc * theta) * left \
            + (wc * (1 - theta) + wb * (theta - lam)) * right
        out = num / den
        dnum = (wa * wc * lam * (yc - ya) * left + wb * wc * (1 - lam) * (yb - yc) * right) / in_w
        lad = torch.log(dnum) - 2 * torch.log(torch.abs(den))

    out = torch.where(inside, out, inputs)
    lad = torch.where(inside, lad, torch.zeros_like(lad))
    return out, lad


def coupling_forward(layers, x, context, bound):
    """[upstream] SplineCoupling._call with identity=True, split_dim=1: y1 = x1; (y2,y3) = spline(x2,x3)
    with parameters from dense_nn(cat[context, y1]).  Returns (y, log|det J|) (sum over the 2 dims)."""
    x1, x2 = x[..., :1], x[..., 1:]
    w, h, d, lam = spline_params(dense_nn(layers, x1, context))
    y2, lad = rational_linear_spline(x2, w, h, d, lam, bound, inverse=False)
    return torch.cat([x1, y2], dim=-1), lad.sum(-1)


def coupling_inverse(layers, y, context, bound):
    """[upstream] SplineCoupling._inverse (identity=True).  The log-det pyro caches on this path is
    MINUS the inverse-direction logabsdet (Spline._inverse: ``_cache_log_detJ = -log_detJ``);
    returned here as the FORWARD log|det J| evaluated through the inverse formulas, like pyro."""
    y1, y2 = y[..., :1], y[..., 1:]
    w, h, d, lam = spline_params(dense_nn(layers, y1, context))
    x2, lad_inv = rational_linear_spline(y2, w, h, d, lam, bound, inverse=True)
    return torch.cat([y1, x2], dim=-1), -lad_inv.sum(-1)
