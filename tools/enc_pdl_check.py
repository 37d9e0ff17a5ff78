"""Race check for the encoder's programmatic-dependent-launch mode: alternate two inputs through one encoder and print
a digest of every output; run once per HF_PDL_EARLY setting and diff the two logs (must be identical), also time it."""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import humaniflow_b200 as hb  # noqa: E402
from humaniflow_b200.synthetic import SMPL_PARENTS, synthetic_proxy_input  # noqa: E402

torch.manual_seed(0)
cfg = hb.get_model_cfg_defaults()
cfg.NUM_RESNET_LAYERS = 50
model = hb.HumaniflowModel('cuda', cfg, SMPL_PARENTS).eval().cuda()
enc = model.image_encoder
xs = [synthetic_proxy_input(32, 18, 256, seed=s).cuda() for s in (1, 2)]
outs = []
for i in range(12):
    outs.append(enc(xs[i & 1]))
torch.cuda.synchronize()
for i, o in enumerate(outs):
    print(i, hashlib.sha1(o.cpu().numpy().tobytes()).hexdigest()[:16], float(o.abs().sum()))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(20):
    enc(xs[i & 1])
b.record()
torch.cuda.synchronize()
print('ms per forward %.4f' % (a.elapsed_time(b) / 20), file=sys.stderr)
