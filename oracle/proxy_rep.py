"""Oracle (test infrastructure): CPU restatement of the reference's proxy-representation builder.

PINNED by tests/golden/proxy_golden.npz, produced by the real models/canny_edge_detector.py and
utils/label_conversions.py in the build container (tests/golden/make_golden_proxy.py).
"""
import numpy as np
import torch
import torch.nn.functional as F


def gaussian_taps(size=5, std=1.0):
    """scipy.signal.windows.gaussian normalised (canny_edge_detector.py:23-24)."""
    n = np.arange(0, size) - (size - 1.0) / 2.0
    w = np.exp(-n ** 2 / (2 * std * std))
    return torch.tensor((w / w.sum()).astype(np.float32))


def canny(img, threshold=0.0, nms=True, std=1.0, size=5):
    """models/canny_edge_detector.py:104-166.  img (B,C,H,W) fp32 -> dict like the reference module."""
    g = gaussian_taps(size, std)
    sob = torch.tensor([[1., 0., -1.], [2., 0., -2.], [1., 0., -1.]])
    B, C = img.shape[:2]
    gx = torch.zeros(B, 1, *img.shape[2:])
    gy = torch.zeros(B, 1, *img.shape[2:])
    for c in range(C):
        bl = F.conv2d(F.conv2d(img[:, [c]], g.view(1, 1, 1, size), padding=(0, size // 2)), g.view(1, 1, size, 1), padding=(size // 2, 0))
        gx += F.conv2d(bl, sob.view(1, 1, 3, 3), padding=1)
        gy += F.conv2d(bl, sob.t().contiguous().view(1, 1, 3, 3), padding=1)
    gx, gy = gx / C, gy / C
    mag = (gx ** 2 + gy ** 2) ** 0.5
    ori = torch.atan2(gy, gx) * (180.0 / np.pi) + 180.0
    ori = torch.round(ori / 45.0) * 45.0
    tmag = mag.clone()
    tmag[mag < threshold] = 0.0
    out = {'grad_magnitude': mag, 'grad_orientation': ori, 'thresholded_grad_magnitude': tmag}
    if nms:
        offs = [(0, 1), (1, 1), (1, 0), (1, -1), (0, -1), (-1, -1), (-1, 0), (-1, 1)]      # neighbour of filter_0 .. filter_315
        filt = torch.zeros(8, 1, 3, 3)
        for k, (dy, dx) in enumerate(offs):
            filt[k, 0, 1, 1] = 1.0
            filt[k, 0, 1 + dy, 1 + dx] = -1.0
        d = F.conv2d(mag, filt, padding=1)
        idx = (ori / 45) % 8
        thin = mag.clone()
        for p in range(4):
            oriented = ((idx == p) * 1 + (idx == p + 4) * 1)
            is_max = (torch.stack([d[:, p], d[:, p + 4]]).min(dim=0)[0] > 0.0).unsqueeze(1)
            thin[((is_max == 0) * 1 * oriented) > 0] = 0.0
        tthin = thin.clone()
        tthin[thin < threshold] = 0.0
        out['thin_edges'] = thin
        out['thresholded_thin_edges'] = tthin
    return out


def heatmaps(joints2D, img_wh, std=4.0):
    """utils/label_conversions.py:106-125."""
    xx, yy = torch.meshgrid(torch.arange(img_wh), torch.arange(img_wh), indexing='ij')
    xx = xx[None, None].float()
    yy = yy[None, None].float()
    u = joints2D[:, :, 0, None, None]
    v = joints2D[:, :, 1, None, None]
    return torch.exp(-(((xx - v) / std) ** 2) / 2 - (((yy - u) / std) ** 2) / 2)


def build_proxy_representation(rgb, joints2D, vis=None, nms=True, threshold=0.0, std=1.0, heat_std=4.0):
    """predict_humaniflow.py:96-110."""
    e = canny(rgb, threshold, nms, std)
    edge = e['thresholded_thin_edges'] if nms else e['thresholded_grad_magnitude']
    h = heatmaps(joints2D, rgb.shape[-1], heat_std)
    if vis is not None:
        h = h * vis[:, :, None, None].float()
    return torch.cat([edge, h], dim=1).float()
