"""Time the heads + flow stage (B=32, N=100, ResNet-50 widths) with CUDA events; HF_FLOW_DBG=1 prints the phase cycle counts of CTA 0."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import humaniflow_b200 as hb  # noqa: E402
from humaniflow_b200.synthetic import SMPL_PARENTS  # noqa: E402

B, N = 32, 100
torch.manual_seed(0)
cfg = hb.get_model_cfg_defaults()
cfg.NUM_RESNET_LAYERS = 50
m = hb.HumaniflowModel('cuda', cfg, SMPL_PARENTS).eval().cuda()
g = torch.Generator().manual_seed(1)
feats = torch.randn(B, 2048, generator=g).abs().cuda()
z = (torch.randn(B, N, 23, 3, generator=g) * 0.6).cuda()
se = torch.randn(B, N, 10, generator=g).cuda()
run = lambda: m(None, input_feats=feats, num_samples=N, base_noise=z, shape_eps=se)
iters = int(os.environ.get('HF_ITERS', 20))
for _ in range(3):
    run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(iters):
    run()
b.record()
torch.cuda.synchronize()
print('heads + flow: %.1f us per call' % (a.elapsed_time(b) / iters * 1e3))
