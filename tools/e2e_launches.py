"""One e2e-style step issued eagerly for an ncu launch list: RGB -> staged proxy rep -> model -> LBS -> per-sample error rows."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import humaniflow_b200 as hb  # noqa: E402
from humaniflow_b200.graphs import predict_step  # noqa: E402
from humaniflow_b200.metrics import pointset_error_rows  # noqa: E402
from humaniflow_b200.proxy_rep import build_proxy_representation  # noqa: E402
from humaniflow_b200.synthetic import SMPL_PARENTS, synthetic_smpl_data  # noqa: E402

B, N = 32, 100
torch.manual_seed(0)
cfg = hb.get_model_cfg_defaults()
cfg.NUM_RESNET_LAYERS = 50
model = hb.HumaniflowModel('cuda', cfg, SMPL_PARENTS).eval().cuda()
smpl = hb.SMPL.from_arrays(synthetic_smpl_data(seed=0, skinning='body_parts'), create_transl=False).cuda()
rgb = torch.rand(B, 3, 256, 256, device='cuda')
j2d = torch.rand(B, 17, 2, device='cuda') * 256
g = torch.Generator().manual_seed(2)
z = (torch.randn(B, N, 23, 3, generator=g) * 0.6).cuda()
se = torch.randn(B, N, 10, generator=g).cuda()
tgt = smpl.tpose(torch.zeros(B, 10, device='cuda')).vertices.clone()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    staged = build_proxy_representation(rgb, j2d, encoder=model.image_encoder)
    _, verts, _ = predict_step(model, smpl, staged, z, se)
    rows = pointset_error_rows(verts.view(B, N, -1, 3), tgt)
    torch.cuda.synchronize()
print('done')
