"""ResNet-18/50 trunk (FC removed) as a parameter container + an op program for the sm_100a encoder.

Mirrors /root/reference/models/resnet.py (same constructor arguments used by the reference,
``resnet18(in_channels, pretrained=False)`` / ``resnet50(...)``, same state-dict key names as torchvision)
but ``forward`` runs the tcgen05/TMA implicit-GEMM kernels through the C-ABI (``hf_encoder_*``):
eval-mode BatchNorm folded into bf16 weights + fp32 bias, bf16 NHWC activations, fp32 accumulation.
Training-mode BatchNorm (batch statistics) and backward are out of scope of this path.
"""
import ctypes
import math
import os

import torch
from torch import nn

from . import _lib

STEM_CIN = 32


class StagedInput:
    """Handle of an encoder input that already sits in the encoder's staged stem layout (written there by a producer
    kernel, see ``ResNet.stem_input``).  Accepted by ``ResNet.forward`` / ``HumaniflowModel.forward`` in place of the image."""

    def __init__(self, encoder, B, H, W, device):
        self.encoder, self.B, self.H, self.W, self.device = encoder, B, H, W, torch.device(device)
        self.shape = (B, encoder.in_channels, H, W)


class _Block(nn.Module):
    def __init__(self, kind, inplanes, planes, stride, downsample):
        super().__init__()
        self.kind = kind
        self.stride = stride
        if kind == 'basic':      # models/resnet.py:40-78
            self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
            self.bn1 = nn.BatchNorm2d(planes)
            self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
            self.bn2 = nn.BatchNorm2d(planes)
        else:                    # models/resnet.py:81-122 (stride on the 3x3)
            self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
            self.bn1 = nn.BatchNorm2d(planes)
            self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
            self.bn2 = nn.BatchNorm2d(planes)
            self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
            self.bn3 = nn.BatchNorm2d(planes * 4)
        if downsample is not None:
            self.downsample = downsample


class ResNet(nn.Module):
    def __init__(self, kind, layers, in_channels):
        super().__init__()
        if in_channels > STEM_CIN:
            raise ValueError('the stem kernel supports up to %d input channels' % STEM_CIN)
        self.kind = kind
        self.in_channels = in_channels
        self.expansion = 1 if kind == 'basic' else 4
        self.inplanes = 64
        self.conv1 = nn.Conv2d(in_channels, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.layer1 = self._make_layer(64, layers[0], 1)
        self.layer2 = self._make_layer(128, layers[1], 2)
        self.layer3 = self._make_layer(256, layers[2], 2)
        self.layer4 = self._make_layer(512, layers[3], 2)
        self.feat_dim = 512 * self.expansion
        for m in self.modules():      # models/resnet.py:160-165
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
        self._enc = None
        self._ws = None
        self._packed_version = None

    def _make_layer(self, planes, blocks, stride):
        downsample = None
        if stride != 1 or self.inplanes != planes * self.expansion:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes * self.expansion, 1, stride, bias=False),
                                       nn.BatchNorm2d(planes * self.expansion))
        layers = [_Block(self.kind, self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * self.expansion
        for _ in range(1, blocks):
            layers.append(_Block(self.kind, self.inplanes, planes, 1, None))
        return nn.Sequential(*layers)

    # ---- packing: fold BN, reorder to (cout,k,k,cin) bf16, build the op program ----
    @staticmethod
    def _fold(conv, bn, cin_pad=None):
        w = conv.weight.detach().double().cpu()
        scale = bn.weight.detach().double().cpu() / torch.sqrt(bn.running_var.detach().double().cpu() + bn.eps)
        bias = bn.bias.detach().double().cpu() - bn.running_mean.detach().double().cpu() * scale
        w = (w * scale[:, None, None, None]).permute(0, 2, 3, 1)         # (cout,kh,kw,cin)
        if cin_pad is not None and cin_pad > w.shape[3]:
            w = torch.nn.functional.pad(w, (0, cin_pad - w.shape[3]))
        wb = w.float().contiguous().to(torch.bfloat16).contiguous()
        return wb.view(torch.int16), bias.float().contiguous()

    def _version(self):
        return tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))

    def _build(self, device):
        lib = _lib.load()
        ops, weights, biases = [], [], []
        free = []
        nbuf = [0]

        def alloc():
            if free:
                return free.pop()
            nbuf[0] += 1
            return nbuf[0] - 1

        def conv_op(conv, bn, src, dst, res, relu, cin_pad=None, branch=None):
            """branch = (conv1x1, bn, src2): the block's downsample path fused into this op (one launch, fp32 accumulation of
            both convolutions in the same accumulator, no bf16 round trip of the branch output)."""
            w, b = self._fold(conv, bn, cin_pad)
            cin = w.shape[3]
            src2, cin2, stride2 = -1, 0, 1
            if branch is not None:
                w2, b2 = self._fold(branch[0], branch[1])
                src2, cin2, stride2 = branch[2], w2.shape[3], branch[0].stride[0]
                w = torch.cat([w.reshape(w.shape[0], -1), w2.reshape(w2.shape[0], -1)], dim=1).contiguous()   # rows [K1 | K2]
                b = (b + b2).contiguous()
            weights.append(w)
            biases.append(b)
            ops.append(_lib.EncOp(_lib.OP_CONV, src, dst, res, cin, w.shape[0], conv.kernel_size[0],
                                  conv.stride[0], conv.padding[0], int(relu), len(weights) - 1, src2, cin2, stride2))

        x = alloc()
        conv_op(self.conv1, self.bn1, -1, x, -1, True, cin_pad=STEM_CIN)
        y = alloc()
        ops.append(_lib.EncOp(_lib.OP_MAXPOOL, x, y, -1, 64, 64, 3, 2, 1, 0, -1))
        free.append(x)
        x = y
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in layer:
                t1 = alloc()
                conv_op(blk.conv1, blk.bn1, x, t1, -1, True)
                last_in = t1
                if blk.kind != 'basic':
                    t2 = alloc()
                    conv_op(blk.conv2, blk.bn2, t1, t2, -1, True)
                    free.append(t1)
                    last_in = t2
                # identity blocks add the input as residual; blocks with a downsample path fuse its 1x1 conv into the last op
                branch = (blk.downsample[0], blk.downsample[1], x) if hasattr(blk, 'downsample') else None
                res = x if branch is None else -1
                out = alloc()
                if blk.kind == 'basic':
                    conv_op(blk.conv2, blk.bn2, last_in, out, res, True, branch=branch)
                else:
                    conv_op(blk.conv3, blk.bn3, last_in, out, res, True, branch=branch)
                free.append(last_in)
                free.append(x)
                x = out
        ops.append(_lib.EncOp(_lib.OP_AVGPOOL, x, x, -1, self.feat_dim, self.feat_dim, 0, 1, 0, 0, -1))
        op_arr = (_lib.EncOp * len(ops))(*ops)
        w_arr = (ctypes.c_void_p * len(weights))(*[w.data_ptr() for w in weights])
        b_arr = (ctypes.c_void_p * len(biases))(*[b.data_ptr() for b in biases])
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.hf_encoder_create(ctypes.byref(h), op_arr, len(ops), w_arr, b_arr, len(weights),
                                             self.in_channels, STEM_CIN, self.feat_dim))
        return h

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._drop()
        return out

    def _drop(self):
        if getattr(self, '_enc', None) is not None:
            try:
                _lib.load().hf_encoder_destroy(self._enc)
            except Exception:
                pass
        self._enc = None
        self._ws = None
        self._packed_version = None

    def set_impl(self, impl):
        """0 = tcgen05 implicit GEMM (product path), 1 = SIMT direct convolution (debug cross-check)."""
        self._impl = impl
        if self._enc is not None:
            _lib.check(_lib.load().hf_encoder_set_impl(self._enc, impl))

    def _ready(self, dev, B, H, W):
        """Packed weights on `dev` + a workspace large enough for (B,H,W)."""
        _lib.require_cuda('ResNet.forward')
        if self.training:
            raise RuntimeError('humaniflow_b200 encoder implements eval-mode BatchNorm only; call .eval()')
        lib = _lib.load()
        ver = self._version()
        if self._enc is None or ver != self._packed_version:
            self._drop()
            self._enc = self._build(dev)
            self._packed_version = ver
            if getattr(self, '_impl', 0):
                _lib.check(lib.hf_encoder_set_impl(self._enc, self._impl))
        with torch.cuda.device(dev):
            nbytes = lib.hf_encoder_workspace_bytes(self._enc, B, H, W)
            if self._ws is None or self._ws.numel() < nbytes or self._ws.device != dev:
                self._ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
                if os.environ.get('HF_POISON_WS'):      # debugging aid: make any read of uninitialised scratch visible
                    self._ws.fill_(0x7f)
        return lib

    def stem_input(self, B, H, W, device):
        """(pointer, (Hp, Wp, Cp, top, left)) of the staged stem input inside the encoder's workspace: bf16 NHWC with a zero
        border.  A producer kernel fills the interior (``proxy_rep.build_proxy_representation(..., encoder=...)``), then
        ``forward(StagedInput)`` runs the trunk without the NCHW -> NHWC conversion pass."""
        dev = torch.device(device)
        lib = self._ready(dev, B, H, W)
        p = ctypes.c_void_p()
        dims = (ctypes.c_int * 5)()
        with torch.cuda.device(dev):
            _lib.check(lib.hf_encoder_stem_input(self._enc, B, H, W, _lib.ptr(self._ws), self._ws.numel(), _lib.stream(),
                                                 ctypes.byref(p), dims))
        return p, tuple(dims)

    def forward(self, x):
        """(B,C,H,W) fp32 (or bf16 / fp16: half the host->device bytes) CUDA tensor, or a ``StagedInput`` -> (B, feat_dim) fp32.
        models/resnet.py:202-217 in eval mode."""
        if isinstance(x, StagedInput):
            if x.encoder is not self:
                raise ValueError('StagedInput belongs to another encoder')
            lib = self._ready(x.device, x.B, x.H, x.W)
            feats = torch.empty(x.B, self.feat_dim, device=x.device, dtype=torch.float32)
            with torch.cuda.device(x.device):
                _lib.check(lib.hf_encoder_forward_staged(self._enc, x.B, x.H, x.W, _lib.ptr(feats), _lib.ptr(self._ws),
                                                         self._ws.numel(), _lib.stream()))
            return feats
        if not x.is_cuda:
            raise RuntimeError('humaniflow_b200 encoder: input must be a CUDA tensor (no CPU fallback)')
        dev = x.device
        B, C, H, W = x.shape
        if C != self.in_channels:
            raise ValueError('expected %d input channels, got %d' % (self.in_channels, C))
        lib = self._ready(dev, B, H, W)
        half = x.dtype in (torch.bfloat16, torch.float16)
        x = x.detach().to(torch.bfloat16).contiguous() if half else _lib.f32c(x)
        with torch.cuda.device(dev):
            feats = torch.empty(B, self.feat_dim, device=dev, dtype=torch.float32)
            fn = lib.hf_encoder_forward_bf16 if half else lib.hf_encoder_forward
            _lib.check(fn(self._enc, _lib.ptr(x), B, H, W, _lib.ptr(feats), _lib.ptr(self._ws), self._ws.numel(), _lib.stream()))
        return feats

    def __del__(self):
        try:
            self._drop()
        except Exception:
            pass


def resnet18(in_channels, pretrained=False, **kwargs):
    if pretrained:
        raise ValueError('pretrained weights are not downloadable here')
    return ResNet('basic', [2, 2, 2, 2], in_channels)


def resnet50(in_channels, pretrained=False, **kwargs):
    if pretrained:
        raise ValueError('pretrained weights are not downloadable here')
    return ResNet('bottleneck', [3, 4, 6, 3], in_channels)
