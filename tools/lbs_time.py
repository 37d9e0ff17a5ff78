"""Time (CUDA events) the LBS stage at M = 3200 for both synthetic skinning-weight structures; HF_SKIN selects one
(for ncu captures)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import humaniflow_b200 as hb  # noqa: E402
from humaniflow_b200.synthetic import synthetic_smpl_data  # noqa: E402
from oracle import so3  # noqa: E402  (tool, not product)

M = int(os.environ.get('HF_M', 3200))
g = torch.Generator().manual_seed(0)
betas = torch.randn(M, 10, generator=g).cuda()
R = so3.batch_rodrigues((torch.randn(M * 24, 3, generator=g) * 0.3)).view(M, 24, 3, 3).float().cuda().contiguous()
styles = [os.environ['HF_SKIN']] if 'HF_SKIN' in os.environ else ['random', 'body_parts']
iters = int(os.environ.get('HF_ITERS', 20))
for style in styles:
    smpl = hb.SMPL.from_arrays(synthetic_smpl_data(seed=0, skinning=style), create_transl=False).cuda()
    for impl in ([int(os.environ.get('HF_IMPL', 0))] if 'HF_SKIN' in os.environ else [0, 3, 2]):
        smpl.set_impl(impl)
        for _ in range(3):
            smpl.lbs(betas, R)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            smpl.lbs(betas, R)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / iters
        print('skinning=%s impl=%d  M=%d  %.1f us per LBS call  (%.0f GB/s algorithmic)' % (style, impl, M, ms * 1e3, 84664 * M / ms / 1e6))
