"""Mesh post-processing helpers: CUDA path through the C-ABI vs the oracle and the golden vectors
of the real reference functions.  fp32; tolerance 2e-6 absolute (summation order over the samples differs from torch's
pairwise mean)."""
import os

import numpy as np
import pytest
import torch

from oracle import sampling as osamp
from oracle import smpl as osmpl
from util import GOLDEN, smpl_data

pytestmark = pytest.mark.gpu
TOL = 2e-6


def test_vertex_variance_golden():
    from humaniflow_b200.sampling import compute_vertex_variance_from_samples
    g = {k: torch.tensor(v) for k, v in np.load(os.path.join(GOLDEN, 'sampling_golden.npz')).items()}
    avg, std = compute_vertex_variance_from_samples(g['verts'].cuda())
    assert avg.shape == (300,) and std.shape == (300, 3)
    assert (avg.cpu() - g['avg']).abs().max().item() <= TOL and (std.cpu() - g['std']).abs().max().item() <= TOL


@pytest.mark.parametrize('B,N,V', [(1, 1, 5), (2, 7, 33), (3, 100, 6890), (32, 100, 6890)])
def test_vertex_variance_batched(B, N, V):
    """Ragged vertex chunks (V % 32 != 0), a single sample (zero variance) and the full BASELINE size."""
    from humaniflow_b200.sampling import compute_vertex_variance_from_samples
    g = torch.Generator().manual_seed(B * 1000 + N)
    x = torch.randn(B, 1, V, 3, generator=g) + 0.1 * torch.randn(B, N, V, 3, generator=g)
    avg, std = compute_vertex_variance_from_samples(x.cuda())
    assert avg.shape == (B, V) and std.shape == (B, V, 3)
    rows = range(B) if B <= 3 else [0, 17, 31]
    for b in rows:
        a_ref, s_ref = osamp.compute_vertex_variance_from_samples(x[b])
        assert (avg[b].cpu() - a_ref).abs().max().item() <= TOL and (std[b].cpu() - s_ref).abs().max().item() <= TOL
    if N == 1:
        assert avg.abs().max().item() == 0 and std.abs().max().item() == 0


def test_project_joints2d():
    from humaniflow_b200.sampling import project_joints2d
    g = {k: torch.tensor(v) for k, v in np.load(os.path.join(GOLDEN, 'sampling_golden.npz')).items()}
    got = project_joints2d(g['joints'].cuda(), g['cam'].cuda(), flip_x=False)
    assert (got.cpu() - g['proj']).abs().max().item() <= 1e-6
    got = project_joints2d(g['joints'].cuda(), g['cam'].cuda(), flip_x=False, img_wh=256)
    assert (got.cpu() - g['pix']).abs().max().item() <= 1e-4          # pixels
    got = project_joints2d(g['joints'].cuda(), g['cam'].cuda(), flip_x=True, img_wh=256)
    ref = osamp.project_joints2d(g['joints'], g['cam'], flip_x=True, img_wh=256)
    assert (got.cpu() - ref).abs().max().item() <= 1e-4
    allj = project_joints2d(g['joints'].cuda(), g['cam'].cuda(), joint_ids=None, flip_x=True)
    assert allj.shape == (12, 90, 2)
    assert (allj.cpu() - osamp.project_joints2d(g['joints'], g['cam'], joint_ids=None)).abs().max().item() <= 1e-6


@pytest.mark.parametrize('M', [1, 5, 200])
def test_tpose_matches_forward_with_zero_pose(M):
    """SMPL.tpose == the oracle's forward with identity rotations == our general path with the default pose."""
    import humaniflow_b200 as hb
    data = smpl_data()
    smpl = hb.SMPL.from_arrays(data, create_transl=False).cuda()
    g = torch.Generator().manual_seed(M)
    betas = torch.randn(M, 10, generator=g)
    transl = torch.randn(M, 3, generator=g)
    out = smpl.tpose(betas.cuda(), transl=transl.cuda())
    eye = torch.eye(3).expand(M, 24, 3, 3)
    v_ref, j_ref = osmpl.smpl_forward(data, betas, eye[:, 1:], eye[:, :1], pose2rot=False, transl=transl)
    assert out.vertices.shape == (M, 6890, 3) and out.joints.shape == (M, 90, 3)
    assert (out.vertices.cpu() - v_ref).norm(dim=-1).max().item() <= 1e-5
    assert (out.joints.cpu() - j_ref).norm(dim=-1).max().item() <= 1e-5
    gen = smpl(betas=betas.cuda(), body_pose=eye[:, 1:].cuda().contiguous(), global_orient=eye[:, :1].cuda().contiguous(),
               transl=transl.cuda(), pose2rot=False)
    assert (out.vertices - gen.vertices).norm(dim=-1).max().item() <= 1e-5


def test_no_cpu_fallback():
    from humaniflow_b200.sampling import compute_vertex_variance_from_samples, project_joints2d
    with pytest.raises(RuntimeError):
        compute_vertex_variance_from_samples(torch.zeros(2, 4, 3))
    with pytest.raises(RuntimeError):
        project_joints2d(torch.zeros(2, 90, 3), torch.zeros(1, 3))
