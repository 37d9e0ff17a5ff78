"""Golden vectors for the proxy-representation builder from the REAL reference modules (build container only):
models/canny_edge_detector.py::CannyEdgeDetector and utils/label_conversions.py::convert_2Djoints_to_gaussian_heatmaps_torch.
    python tests/golden/make_golden_proxy.py   ->  tests/golden/proxy_golden.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference')
from models.canny_edge_detector import CannyEdgeDetector  # noqa: E402
from utils.label_conversions import convert_2Djoints_to_gaussian_heatmaps_torch  # noqa: E402

g = torch.Generator().manual_seed(0)
H = 64
# smooth random image (sum of blobs) so that there are real edges, plus a sharp rectangle and noise
yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(H).float(), indexing='ij')
img = torch.zeros(2, 3, H, H)
for b in range(2):
    for c in range(3):
        for _ in range(6):
            cx, cy, s = torch.rand(3, generator=g) * torch.tensor([H, H, 10.0]) + torch.tensor([0., 0., 3.])
            img[b, c] += torch.rand(1, generator=g) * torch.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))
img[0, :, 20:40, 10:30] += 0.5
img = (img + 0.02 * torch.rand(img.shape, generator=g)).clamp(0, 1)
out = {}
with torch.no_grad():
    for thr, nms in ((0.0, True), (0.2, True), (0.1, False)):
        m = CannyEdgeDetector(non_max_suppression=nms, gaussian_filter_std=1.0, gaussian_filter_size=5, threshold=thr)
        r = m(img)
        key = 'thresholded_thin_edges' if nms else 'thresholded_grad_magnitude'
        out['edge_%g_%d' % (thr, int(nms))] = r[key].numpy()
        out['mag'] = r['grad_magnitude'].numpy()
        out['ori'] = r['grad_orientation'].numpy()
j2d = torch.rand(2, 5, 2, generator=g) * (H + 10) - 5
out['heat'] = convert_2Djoints_to_gaussian_heatmaps_torch(j2d, H, std=4.0).numpy()
np.savez_compressed(os.path.join(HERE, 'proxy_golden.npz'), img=img.numpy(), j2d=j2d.numpy(), **out)
print('wrote proxy_golden.npz', {k: v.shape for k, v in out.items()})
