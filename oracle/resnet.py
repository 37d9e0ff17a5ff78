"""Oracle (test infrastructure): ResNet-18/50 trunk (FC removed) in eval mode, functional form.

Follows /root/reference/models/resnet.py:40-122 (BasicBlock / Bottleneck, stride on the 3x3 conv)
and :202-217 (forward).  PINNED by tests/golden/resnet_golden.npz (real reference class run on
deterministic weights).  Takes a state dict with torchvision key names under ``prefix``.
"""
import torch.nn.functional as F

BLOCKS = {18: ('basic', [2, 2, 2, 2]), 50: ('bottleneck', [3, 4, 6, 3])}


def _bn(sd, name, x):
    return F.batch_norm(x, sd[name + '.running_mean'], sd[name + '.running_var'],
                        sd[name + '.weight'], sd[name + '.bias'], training=False, eps=1e-5)


def resnet_forward(sd, x, num_layers, prefix='image_encoder.'):
    """(B,C,H,W) fp32 -> (B, 512|2048) fp32."""
    kind, counts = BLOCKS[num_layers]
    g = lambda k: sd[prefix + k]
    sdp = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    x = F.conv2d(x, g('conv1.weight'), stride=2, padding=3)
    x = F.relu(_bn(sdp, 'bn1', x))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li, n in enumerate(counts, start=1):
        for bi in range(n):
            p = 'layer%d.%d.' % (li, bi)
            stride = 2 if (li > 1 and bi == 0) else 1
            identity = x
            if kind == 'basic':
                out = F.conv2d(x, sdp[p + 'conv1.weight'], stride=stride, padding=1)
                out = F.relu(_bn(sdp, p + 'bn1', out))
                out = F.conv2d(out, sdp[p + 'conv2.weight'], padding=1)
                out = _bn(sdp, p + 'bn2', out)
            else:
                out = F.relu(_bn(sdp, p + 'bn1', F.conv2d(x, sdp[p + 'conv1.weight'])))
                out = F.conv2d(out, sdp[p + 'conv2.weight'], stride=stride, padding=1)
                out = F.relu(_bn(sdp, p + 'bn2', out))
                out = _bn(sdp, p + 'bn3', F.conv2d(out, sdp[p + 'conv3.weight']))
            if p + 'downsample.0.weight' in sdp:
                identity = _bn(sdp, p + 'downsample.1',
                               F.conv2d(x, sdp[p + 'downsample.0.weight'], stride=stride))
            x = F.relu(out + identity)
    return x.mean(dim=(2, 3))
