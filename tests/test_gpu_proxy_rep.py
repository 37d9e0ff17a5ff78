"""Proxy-representation builder: fused CUDA kernel through the C-ABI vs the golden vectors of the
real reference modules and the oracle.  fp32; gradient magnitude / heatmaps within 2e-6; the edge map is a per-pixel DECISION
(orientation bin, non-maximum test) on fp32 values whose last bit depends on the convolution's summation order, so a small
fraction of pixels (<= 0.3 %) may decide differently -- every other pixel must match to 2e-6."""
import os

import numpy as np
import pytest
import torch

from oracle import proxy_rep as opr
from util import GOLDEN

pytestmark = pytest.mark.gpu
TOL = 2e-6


def _edge_match(got, ref, max_flip=0.003):
    diff = (got - ref).abs()
    flipped = diff > TOL
    assert flipped.float().mean().item() <= max_flip, flipped.float().mean().item()
    # a flipped pixel is a decision flip: one side is exactly 0, the other the gradient magnitude
    assert ((got[flipped] == 0) | (ref[flipped] == 0)).all()


def test_golden():
    from humaniflow_b200.proxy_rep import CannyEdgeDetector, build_proxy_representation
    g = np.load(os.path.join(GOLDEN, 'proxy_golden.npz'))
    img = torch.tensor(g['img']).cuda()
    for thr, nms in ((0.0, True), (0.2, True), (0.1, False)):
        m = CannyEdgeDetector(non_max_suppression=nms, gaussian_filter_std=1.0, gaussian_filter_size=5, threshold=thr).cuda()
        r = m(img)
        key = 'thresholded_thin_edges' if nms else 'thresholded_grad_magnitude'
        assert r[key].shape == (2, 1, 64, 64)
        _edge_match(r[key].cpu(), torch.tensor(g['edge_%g_%d' % (thr, int(nms))]))
        assert (r['grad_magnitude'].cpu() - torch.tensor(g['mag'])).abs().max().item() <= TOL
        assert (r['grad_orientation'].cpu() != torch.tensor(g['ori'])).float().mean().item() <= 0.003
    j2d = torch.tensor(g['j2d']).cuda()
    rep = build_proxy_representation(img, j2d)
    assert rep.shape == (2, 6, 64, 64)
    assert (rep[:, 1:].cpu() - torch.tensor(g['heat'])).abs().max().item() <= TOL
    _edge_match(rep[:, :1].cpu(), torch.tensor(g['edge_0_1']))


@pytest.mark.parametrize('B,H', [(1, 33), (3, 256), (32, 256)])
def test_against_oracle(B, H):
    """Ragged tiles (H % 32 != 0), the reference resolution and the BASELINE batch; visibility flags applied to the heatmaps."""
    from humaniflow_b200.proxy_rep import build_proxy_representation
    g = torch.Generator().manual_seed(B + H)
    yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(H).float(), indexing='ij')
    img = torch.zeros(B, 3, H, H)
    for b in range(min(B, 4)):
        for c in range(3):
            for _ in range(5):
                cx, cy, s = (torch.rand(3, generator=g) * torch.tensor([H, H, H / 8.0]) + torch.tensor([0., 0., 2.])).tolist()
                img[b, c] += torch.rand(1, generator=g).item() * torch.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))
    img = (img + 0.05 * torch.rand(img.shape, generator=g)).clamp(0, 1)
    j2d = torch.rand(B, 17, 2, generator=g) * H
    vis = torch.rand(B, 17, generator=g) > 0.3
    rep = build_proxy_representation(img.cuda(), j2d.cuda(), vis.cuda())
    assert rep.shape == (B, 18, H, H) and rep.dtype == torch.float32
    rows = list(range(B)) if B <= 3 else [0, 2, 31]
    ref = opr.build_proxy_representation(img[rows], j2d[rows], vis[rows])
    assert (rep[rows, 1:].cpu() - ref[:, 1:]).abs().max().item() <= TOL
    _edge_match(rep[rows, :1].cpu(), ref[:, :1], max_flip=0.01)


def test_feeds_the_model():
    """The device-built proxy representation goes straight into HumaniflowModel.forward."""
    import humaniflow_b200 as hb
    from humaniflow_b200.proxy_rep import build_proxy_representation
    from humaniflow_b200.synthetic import SMPL_PARENTS
    torch.manual_seed(0)
    cfg = hb.get_model_cfg_defaults()
    model = hb.HumaniflowModel('cuda', cfg, SMPL_PARENTS).eval().cuda()
    rgb = torch.rand(2, 3, 256, 256, device='cuda')
    j2d = torch.rand(2, 17, 2, device='cuda') * 256
    out = model(build_proxy_representation(rgb, j2d), num_samples=4)
    assert out['pose_rotmats_samples'].shape == (2, 4, 23, 3, 3) and torch.isfinite(out['pose_rotmats_samples']).all()


def test_no_cpu_fallback():
    from humaniflow_b200.proxy_rep import build_proxy_representation
    with pytest.raises(RuntimeError):
        build_proxy_representation(torch.zeros(1, 3, 8, 8), torch.zeros(1, 17, 2))
