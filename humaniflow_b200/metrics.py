"""On-device 3-D error metrics of the evaluation path.

`pointset_errors` returns, per predicted point set, the mean point error against its image's target, plain and after the
reference's two alignments (utils/eval_utils.py:105-125 scale + translation, :62-102 Procrustes).  Everything
metrics/eval_metrics_tracker.py:119-280 accumulates for PVE / PVE-SC / PVE-PA / PVE-T(-SC) / MPJPE(-SC/-PA) and the
"samples_min" variants is a sum or a minimum of these values; `samples_min` is that minimum.  No CPU fallback.
"""
import torch

from . import _lib


def _pse(lib, p, t, B, N, P, out):
    nbytes = lib.hf_pointset_errors_workspace_bytes(B, N)
    ws = torch.empty(nbytes, device=p.device, dtype=torch.uint8)      # torch's caching allocator: a pointer bump, capturable
    _lib.check(lib.hf_pointset_errors_ws(_lib.ptr(p), _lib.ptr(t), B, N, P, _lib.ptr(out), _lib.ptr(ws), nbytes, _lib.stream()))


def pointset_errors(pred, target):
    """pred (B,N,P,3) or (B,P,3) CUDA fp32; target (B,P,3) -> dict of (B,N) (or (B,)) tensors 'plain', 'sc', 'pa':
    mean over the P points of ||pred - target|| without alignment, after scale-and-translation correction and after
    Procrustes alignment."""
    _lib.require_cuda('pointset_errors')
    if not pred.is_cuda:
        raise RuntimeError('humaniflow_b200.metrics: inputs must be CUDA tensors (no CPU fallback)')
    p = _lib.f32c(pred)
    t = _lib.f32c(target).to(p.device)
    squeeze = p.dim() == 3
    if squeeze:
        p = p[:, None]
    B, N, P, _ = p.shape
    if t.shape != (B, P, 3):
        raise ValueError('target shape %s does not match predictions %s' % (tuple(t.shape), tuple(p.shape)))
    out = torch.empty(B, N, 3, device=p.device, dtype=torch.float32)
    with torch.cuda.device(p.device):
        _pse(_lib.load(), p, t, B, N, P, out)
    res = {'plain': out[..., 0], 'sc': out[..., 1], 'pa': out[..., 2]}
    return {k: v[:, 0] for k, v in res.items()} if squeeze else res


def pointset_error_rows(pred, target):
    """pred (B,N,P,3), target (B,P,3) -> (B,6) per-image rows [min over the samples of plain / sc / pa | mean over the samples of
    plain / sc / pa]: hf_pointset_errors followed by ONE reduction launch (hf_samples_reduce) -- the evaluation loop's
    "samples_min" and sample-mean metrics of a batch, ready for a single small device->host copy."""
    _lib.require_cuda('pointset_error_rows')
    if not pred.is_cuda or pred.dim() != 4:
        raise RuntimeError('humaniflow_b200.metrics: pred must be a (B,N,P,3) CUDA tensor (no CPU fallback)')
    p = _lib.f32c(pred)
    t = _lib.f32c(target).to(p.device)
    B, N, P, _ = p.shape
    if t.shape != (B, P, 3):
        raise ValueError('target shape %s does not match predictions %s' % (tuple(t.shape), tuple(p.shape)))
    err = torch.empty(B, N, 3, device=p.device, dtype=torch.float32)
    rows = torch.empty(B, 6, device=p.device, dtype=torch.float32)
    lib = _lib.load()
    with torch.cuda.device(p.device):
        _pse(lib, p, t, B, N, P, err)
        _lib.check(lib.hf_samples_reduce(_lib.ptr(err), B, N, 3, _lib.ptr(rows), _lib.stream()))
    return rows


def samples_min(per_sample_errors):
    """(B,N) per-sample mean errors -> (B,) error of the best sample of every image
    (eval_metrics_tracker.py:201-280: argmin over the samples of the mean error)."""
    return per_sample_errors.min(dim=1).values


def sample_stats(points, target=None, weights=None):
    """points (B,N,P,D) CUDA fp32 with D = 3 (sampled vertices / 3-D joints) or 2 (projected joints); target (B,P,D) or None;
    weights (B,P) visibility flags or None -> dict of (B,) per-frame values:
      'diversity'  mean over (samples, points) of w ||x - mean over the samples||   (eval_metrics_tracker.py:397-433:
                   verts3D_sample_diversity, joints3D_sample_diversity and its (in)visible-joint forms)
      'l2e'        sum w ||x - target|| / (N sum w)                                  (:339-374: (input_)joints2Dsamples-L2E)"""
    _lib.require_cuda('sample_stats')
    if not points.is_cuda:
        raise RuntimeError('humaniflow_b200.metrics: inputs must be CUDA tensors (no CPU fallback)')
    x = _lib.f32c(points)
    B, N, P, D = x.shape
    t = None if target is None else _lib.f32c(target).to(x.device)
    w = None if weights is None else _lib.f32c(weights.to(torch.float32)).to(x.device)
    if t is not None and t.shape != (B, P, D):
        raise ValueError('target shape %s does not match points %s' % (tuple(t.shape), tuple(x.shape)))
    out = torch.empty(B, 2, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().hf_sample_stats(_lib.ptr(x), _lib.ptr(t), _lib.ptr(w), B, N, P, D, _lib.ptr(out), _lib.stream()))
    res = {'diversity': out[:, 0]}
    if t is not None:
        res['l2e'] = out[:, 1]
    return res
