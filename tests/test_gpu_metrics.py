"""3-D error metrics: CUDA path through the C-ABI vs the oracle and the golden vectors of the real
utils/eval_utils.py.  Tolerance 1e-5 relative + 1e-6 m absolute (the reference runs numpy fp32 incl. an fp32 SVD; the
kernel accumulates moments in fp32 per thread / fp64 across the block and solves the 3x3 problem in fp64)."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics as omet
from util import GOLDEN

pytestmark = pytest.mark.gpu


def _close(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= 1e-5 * np.abs(b) + 1e-6)


def test_golden():
    from humaniflow_b200.metrics import pointset_errors, samples_min
    g = np.load(os.path.join(GOLDEN, 'metrics_golden.npz'))
    e = pointset_errors(torch.tensor(g['pred']).cuda(), torch.tensor(g['target']).cuda())
    for k in ('plain', 'sc', 'pa'):
        assert e[k].shape == (3, 4) and _close(e[k].cpu().numpy(), g[k]), (k, e[k].cpu().numpy(), g[k])
    assert _close(samples_min(e['pa']).cpu().numpy(), g['pa'].min(1))
    one = pointset_errors(torch.tensor(g['pred'][:, 0]).cuda(), torch.tensor(g['target']).cuda())       # point-estimate form (B,P,3)
    assert one['sc'].shape == (3,) and _close(one['sc'].cpu().numpy(), g['sc'][:, 0])


@pytest.mark.parametrize('B,N,P', [(2, 3, 14), (1, 1, 17), (4, 25, 6890), (32, 100, 6890)])
def test_against_oracle(B, N, P):
    """Joint sets (MPJPE family), meshes, and the full BASELINE size (spot-checked images)."""
    from humaniflow_b200.metrics import pointset_errors
    g = torch.Generator().manual_seed(B * 100 + N)
    target = torch.randn(B, P, 3, generator=g) * 0.3 + torch.tensor([0.1, -0.2, 2.5])
    pred = target[:, None] * (1 + 0.1 * torch.randn(B, N, 1, 1, generator=g)) + 0.05 * torch.randn(B, N, P, 3, generator=g) \
        + 0.1 * torch.randn(B, N, 1, 3, generator=g)
    e = pointset_errors(pred.cuda(), target.cuda())
    rows = list(range(B)) if B <= 4 else [0, 13, 31]
    ref = omet.pointset_errors(pred[rows].numpy(), target[rows].numpy())
    for k in ('plain', 'sc', 'pa'):
        assert _close(e[k][rows].cpu().numpy(), ref[k]), k


def test_exact_similarity_is_zero_after_procrustes():
    from humaniflow_b200.metrics import pointset_errors
    g = torch.Generator().manual_seed(5)
    target = torch.randn(2, 300, 3, generator=g)
    from oracle import so3
    R = so3.batch_rodrigues(torch.tensor([[0.3, -1.2, 0.4], [2.0, 0.1, -0.5]])).float()
    pred = 1.7 * torch.einsum('bij,bpj->bpi', R, target) + torch.tensor([0.5, -1.0, 2.0])
    e = pointset_errors(pred.cuda(), target.cuda())
    assert e['pa'].abs().max().item() <= 2e-6 and e['plain'].min().item() > 0.5


def test_no_cpu_fallback():
    from humaniflow_b200.metrics import pointset_errors
    with pytest.raises(RuntimeError):
        pointset_errors(torch.zeros(1, 2, 5, 3), torch.zeros(1, 5, 3))


def test_sample_stats_golden_and_oracle():
    """hf_sample_stats (sample diversity, samples-L2E; eval_metrics_tracker.py:339-433) vs golden vectors of the REAL tracker class
    and vs the oracle at the benchmark size (32 images x 100 sampled meshes)."""
    from humaniflow_b200.metrics import sample_stats
    g = np.load(os.path.join(GOLDEN, 'tracker_golden.npz'))
    cu = lambda k: torch.tensor(g[k].astype(np.float32)).cuda()
    vis = cu('in_vis')
    assert _close(sample_stats(cu('verts'))['diversity'].cpu().numpy(), g['verts3D_sample_diversity'])
    assert _close(sample_stats(cu('j3d'))['diversity'].cpu().numpy(), g['joints3D_sample_diversity'])
    assert _close(sample_stats(cu('j3d'), weights=1.0 - vis)['diversity'].cpu().numpy(), g['joints3D_invis_sample_diversity'])
    assert _close(sample_stats(cu('j3d'), weights=vis)['diversity'].cpu().numpy(), g['joints3D_vis_sample_diversity'])
    assert _close(sample_stats(cu('j2d_samples'), cu('tgt_j2d'), cu('tgt_vis'))['l2e'].cpu().numpy(), g['joints2Dsamples_L2E'])
    assert _close(sample_stats(cu('j2d_samples'), cu('in_j2d'), vis)['l2e'].cpu().numpy(), g['input_joints2Dsamples_L2E'])
    rs = np.random.RandomState(5)
    x = (rs.standard_normal((32, 100, 6890, 3)) * 0.05 + rs.standard_normal((32, 1, 6890, 3)) * 0.3).astype(np.float32)
    got = sample_stats(torch.tensor(x).cuda())['diversity'].cpu().numpy()
    assert got.shape == (32,) and _close(got, omet.sample_stats(x)['diversity'])


def test_pointset_error_rows_is_min_and_mean_of_the_per_sample_errors():
    """hf_samples_reduce: (B,N,3) per-sample errors -> (B,6) = [min over samples | mean over samples], against torch on the
    output of hf_pointset_errors itself (exact for the minimum, fp32 summation-order tolerance for the mean)."""
    from humaniflow_b200.metrics import pointset_error_rows, pointset_errors
    g = torch.Generator().manual_seed(11)
    for B, N, P in [(3, 7, 50), (2, 300, 17), (1, 1, 9)]:
        pred = torch.randn(B, N, P, 3, generator=g).cuda()
        tgt = torch.randn(B, P, 3, generator=g).cuda()
        err = pointset_errors(pred, tgt)
        rows = pointset_error_rows(pred, tgt)
        assert rows.shape == (B, 6)
        for k, name in enumerate(('plain', 'sc', 'pa')):
            assert torch.equal(rows[:, k], err[name].min(dim=1).values)
            assert torch.allclose(rows[:, 3 + k], err[name].double().mean(dim=1).float(), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('B,N,P', [(6, 100, 6890), (40, 16, 17), (3, 171, 301)])
def test_three_launch_form_equals_the_fused_kernel(B, N, P):
    """>= 512 point sets with a workspace run as moments / solve / errors launches (hf_pointset_errors_ws); the per-thread
    arithmetic and the block-sum order are those of the fused kernel, so meshes (staged fused path) agree bit for bit and the other
    shapes to fp32 rounding; both agree with the oracle."""
    from humaniflow_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(B + N + P)
    target = torch.randn(B, P, 3, generator=g) * 0.3 + torch.tensor([0.1, -0.2, 2.5])
    pred = target[:, None] * (1 + 0.1 * torch.randn(B, N, 1, 1, generator=g)) + 0.05 * torch.randn(B, N, P, 3, generator=g)
    pd, td = pred.cuda().contiguous(), target.cuda().contiguous()
    fused = torch.empty(B, N, 3, device='cuda')
    split = torch.empty(B, N, 3, device='cuda')
    _lib.check(lib.hf_pointset_errors(_lib.ptr(pd), _lib.ptr(td), B, N, P, _lib.ptr(fused), _lib.stream()))
    nbytes = lib.hf_pointset_errors_workspace_bytes(B, N)
    ws = torch.empty(nbytes, device='cuda', dtype=torch.uint8)
    _lib.check(lib.hf_pointset_errors_ws(_lib.ptr(pd), _lib.ptr(td), B, N, P, _lib.ptr(split), _lib.ptr(ws), nbytes, _lib.stream()))
    torch.cuda.synchronize()
    if P == 6890:
        assert torch.equal(fused, split)
    assert torch.allclose(fused, split, rtol=2e-6, atol=1e-7)
    rows = [0, B - 1]
    ref = omet.pointset_errors(pred[rows].numpy(), target[rows].numpy())
    for i, k in enumerate(('plain', 'sc', 'pa')):
        assert _close(split[rows, :, i].cpu().numpy(), ref[k]), k
