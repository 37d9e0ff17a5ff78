"""Two steps of the bench workload (first one warm-up) for ncu captures; no timing, no CPU baseline."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import humaniflow_b200 as hb  # noqa: E402
from humaniflow_b200.synthetic import SMPL_PARENTS, synthetic_proxy_input, synthetic_smpl_data  # noqa: E402

B, N = int(os.environ.get('HF_B', 32)), int(os.environ.get('HF_N', 100))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
cfg = hb.get_model_cfg_defaults()
cfg.NUM_RESNET_LAYERS = 50
model = hb.HumaniflowModel('cuda', cfg, SMPL_PARENTS).eval().cuda()
smpl = hb.SMPL.from_arrays(synthetic_smpl_data(seed=0, skinning='body_parts'), create_transl=False).cuda()
x = synthetic_proxy_input(B, 18, 256, seed=1).cuda()
g = torch.Generator().manual_seed(2)
z = (torch.randn(B, N, 23, 3, generator=g) * 0.6).cuda()
se = torch.randn(B, N, 10, generator=g).cuda()
from humaniflow_b200.graphs import predict_step  # noqa: E402
from humaniflow_b200.metrics import sample_stats  # noqa: E402
for _ in range(steps):          # the step bench.py times (issued eagerly here: ncu profiles kernel by kernel)
    out, verts, joints = predict_step(model, smpl, x, z, se)
    rows = sample_stats(joints.view(B, N, -1, 3))['diversity']
    torch.cuda.synchronize()
print('done', verts.shape)
