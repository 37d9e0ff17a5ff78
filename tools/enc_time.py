"""Time the ResNet-50 encoder forward (B=32, 18ch, 256px) with CUDA events; env toggles of encoder.cu apply."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import humaniflow_b200 as hb  # noqa: E402
from humaniflow_b200.synthetic import SMPL_PARENTS, synthetic_proxy_input  # noqa: E402

B = int(os.environ.get('ENC_B', '32'))
torch.manual_seed(0)
cfg = hb.get_model_cfg_defaults()
cfg.NUM_RESNET_LAYERS = 50
enc = hb.HumaniflowModel('cuda', cfg, SMPL_PARENTS).eval().cuda().image_encoder
x = synthetic_proxy_input(B, 18, 256, seed=1).cuda()
for _ in range(5):
    f = enc(x)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if os.environ.get('ENC_GRAPH'):
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        enc(x)
        torch.cuda.synchronize()
        with torch.cuda.graph(gr, stream=s):
            f = enc(x)
    torch.cuda.synchronize()
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    a.record()
    for _ in range(20):
        gr.replay()
    e.record()
else:
    a.record()
    for _ in range(20):
        f = enc(x)
    e.record()
torch.cuda.synchronize()
ms = a.elapsed_time(e) / 20
print('encoder B=%d: %.3f ms/forward  (%.1f TFLOP/s)  env: %s' % (B, ms, 12.22e9 * B / ms / 1e9,
      {k: v for k, v in os.environ.items() if k.startswith('HF_')}))
