"""Input proxy representation on the device: CUDA mirror of the reference's
``models/canny_edge_detector.py::CannyEdgeDetector`` + ``utils/label_conversions.py::
convert_2Djoints_to_gaussian_heatmaps_torch`` as ONE fused kernel (`hf_proxy_rep`).

    edge = CannyEdgeDetector(non_max_suppression=True, gaussian_filter_std=1.0, gaussian_filter_size=5, threshold=0.0)
    edge(rgb)['thresholded_thin_edges']                        # (B,1,H,W), like the reference module
    build_proxy_representation(rgb, joints2D, joints_vis)      # (B,18,H,W) = cat[edges, heatmaps * vis]  (predict_humaniflow.py:96-110)

No CPU fallback: CUDA tensors only.
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _lib


def _gaussian_taps(size, std):
    """scipy.signal.windows.gaussian(size, std) normalised to sum 1, as float32 (canny_edge_detector.py:23-24,32)."""
    n = np.arange(0, size) - (size - 1.0) / 2.0
    w = np.exp(-n ** 2 / (2 * std * std))
    return (w / w.sum()).astype(np.float32)


def _run(rgb, joints2D, vis, taps, threshold, nms, heat_std, debug=False):
    _lib.require_cuda('proxy representation')
    if not rgb.is_cuda:
        raise RuntimeError('humaniflow_b200.proxy_rep: inputs must be CUDA tensors (no CPU fallback)')
    x = _lib.f32c(rgb)
    B, C, H, W = x.shape
    j = None if joints2D is None else _lib.f32c(joints2D).to(x.device)
    J = 0 if j is None else j.shape[1]
    v = None if vis is None else _lib.f32c(vis.to(torch.float32)).to(x.device)
    out = torch.empty(B, 1 + J, H, W, device=x.device, dtype=torch.float32)
    mag = torch.empty(B, 1, H, W, device=x.device) if debug else None
    ori = torch.empty(B, 1, H, W, device=x.device) if debug else None
    g = (ctypes.c_float * 5)(*[float(t) for t in taps])
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().hf_proxy_rep(_lib.ptr(x), _lib.ptr(j), _lib.ptr(v), B, C, H, W, J, ctypes.cast(g, ctypes.c_void_p),
                                            float(threshold), int(bool(nms)), float(heat_std), _lib.ptr(out), _lib.ptr(mag),
                                            _lib.ptr(ori), _lib.stream()))
    return out, mag, ori


class CannyEdgeDetector(nn.Module):
    """models/canny_edge_detector.py:11-166, same constructor; ``forward`` returns the dict entries the callers read:
    'thresholded_thin_edges' (non_max_suppression=True) or 'thresholded_grad_magnitude', plus 'grad_magnitude' and
    'grad_orientation'."""

    def __init__(self, non_max_suppression=True, gaussian_filter_std=1.0, gaussian_filter_size=5, threshold=0.2):
        super().__init__()
        if gaussian_filter_size != 5:
            raise ValueError('humaniflow_b200.CannyEdgeDetector: only the 5-tap Gaussian of the reference configuration is built')
        self.threshold = threshold
        self.non_max_suppression = non_max_suppression
        self.register_buffer('gaussian_taps', torch.tensor(_gaussian_taps(gaussian_filter_size, gaussian_filter_std)))

    def forward(self, img):
        out, mag, ori = _run(img, None, None, self.gaussian_taps.tolist(), self.threshold, self.non_max_suppression, 1.0, debug=True)
        key = 'thresholded_thin_edges' if self.non_max_suppression else 'thresholded_grad_magnitude'
        return {key: out, 'grad_magnitude': mag, 'grad_orientation': ori}


def build_proxy_representation(rgb, joints2D, joints_vis=None, edge_nms=True, edge_threshold=0.0, edge_gaussian_std=1.0,
                               heatmap_std=4.0, encoder=None):
    """rgb (B,3,H,W) in [0,1], joints2D (B,17,2) pixel (column,row), joints_vis (B,17) bool/float or None ->
    (B,18,H,W) fp32 = cat[edge map, heatmaps * vis]; defaults = configs/humaniflow_config.py:29-34.

    ``encoder=model.image_encoder``: the kernel writes straight into that encoder's staged stem input (bf16 NHWC inside the
    encoder's workspace) and a ``StagedInput`` handle is returned instead of a tensor; ``model(handle, ...)`` then skips the
    fp32 NCHW intermediate and the layout pass ("fused into the encoder's first-layer staging")."""
    taps = _gaussian_taps(5, edge_gaussian_std)
    if encoder is None:
        out, _, _ = _run(rgb, joints2D, joints_vis, taps, edge_threshold, edge_nms, heatmap_std)
        return out
    from .resnet import StagedInput
    _lib.require_cuda('proxy representation')
    if not rgb.is_cuda:
        raise RuntimeError('humaniflow_b200.proxy_rep: inputs must be CUDA tensors (no CPU fallback)')
    x = _lib.f32c(rgb)
    B, C, H, W = x.shape
    j = _lib.f32c(joints2D).to(x.device)
    J = j.shape[1]
    if 1 + J != encoder.in_channels:
        raise ValueError('encoder expects %d channels, proxy representation has %d' % (encoder.in_channels, 1 + J))
    v = None if joints_vis is None else _lib.f32c(joints_vis.to(torch.float32)).to(x.device)
    p, (Hp, Wp, Cp, top, left) = encoder.stem_input(B, H, W, x.device)
    g = (ctypes.c_float * 5)(*[float(t) for t in taps])
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().hf_proxy_rep_staged(_lib.ptr(x), _lib.ptr(j), _lib.ptr(v), B, C, H, W, J, ctypes.cast(g, ctypes.c_void_p),
                                                   float(edge_threshold), int(bool(edge_nms)), float(heatmap_std), p, Hp, Wp, Cp,
                                                   top, left, _lib.stream()))
    return StagedInput(encoder, B, H, W, x.device)
