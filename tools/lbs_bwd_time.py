"""CUDA-event timing of the LBS forward and backward (hf_lbs_forward / hf_lbs_backward through the autograd Function)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import humaniflow_b200 as hb  # noqa: E402
from humaniflow_b200.synthetic import synthetic_smpl_data  # noqa: E402

smpl = hb.SMPL.from_arrays(synthetic_smpl_data(seed=0, skinning='body_parts'), create_transl=False).cuda()
for M, verts in ((32, False), (32, True), (3200, False), (3200, True)):
    g = torch.Generator().manual_seed(M)
    betas = torch.randn(M, 10, generator=g).cuda().requires_grad_()
    R = torch.linalg.qr(torch.randn(M, 24, 3, 3, generator=g))[0].cuda().requires_grad_()
    wv = torch.randn(M, 6890, 3, device='cuda')
    wj = torch.randn(M, 90, 3, device='cuda')

    def step():
        out = smpl(betas=betas, body_pose=R[:, 1:], global_orient=R[:, :1], pose2rot=False)
        loss = (out.joints * wj).sum() + ((out.vertices * wv).sum() if verts else 0.)
        betas.grad = R.grad = None
        return out, loss

    for _ in range(3):
        out, loss = step()
        loss.backward()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.
    for _ in range(10):
        ev[0].record()
        out, loss = step()
        ev[1].record()
        loss.backward()
        ev[2].record()
        torch.cuda.synchronize()
        tf += ev[0].elapsed_time(ev[1])
        tb += ev[1].elapsed_time(ev[2])
    print('M=%5d %-22s forward+loss %8.1f us   backward %8.1f us' % (M, 'joints+vertices loss' if verts else 'joints loss', tf * 100, tb * 100))
