"""Drop-in for the reference's models/segmentation_models/ResnetUnet.py: torchvision ResNet-50 encoder + U-Net decoder
(same file / class / submodule names => identical state_dict keys, 393 entries).

Differences forced by the environment, not by the design:
  * the reference builds `models.resnet50(weights=ResNet50_Weights.DEFAULT)` (ResnetUnet.py:32), which downloads a
    checkpoint; offline that raises, so this module falls back to random init with a warning (load the reference's
    checkpoint with load_state_dict to get the pretrained encoder);
  * `freeze=True` (the reference default) runs the encoder through forward-only kernels under no_grad; `freeze=False`
    (ResnetUnet.py:29-30) trains it through the res_* autograd ops of ops_resnet.py.
"""
import warnings

import torch
import torch.nn as nn

from ... import ops, ops_resnet as R
from ...blocks import basic_block, check_image, conv_bn_act  # noqa: F401  (basic_block re-exported)


def _bn_update(bn, stats, y):
    if bn.training:
        n, h, w, _ = y.shape
        ops.bn_update_running_(stats, n * h * w, float(bn.momentum), bn.running_mean, bn.running_var,
                               bn.num_batches_tracked)


def _enc_conv_bn(x, conv, bn, relu, identity=None):
    """torchvision conv (bias=False) + BatchNorm (+identity) (+ReLU), forward only."""
    y, stats = R.enc_conv_bn(x, conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, identity,
                             conv.stride[0], bn.training, float(bn.eps), relu)
    _bn_update(bn, stats, y)
    return y


def _bottleneck(blk, x):
    """torchvision.models.resnet.Bottleneck.forward (v1.5: the stride sits on conv2)"""
    o = _enc_conv_bn(x, blk.conv1, blk.bn1, True)
    o = _enc_conv_bn(o, blk.conv2, blk.bn2, True)
    idt = x if blk.downsample is None else _enc_conv_bn(x, blk.downsample[0], blk.downsample[1], False)
    return _enc_conv_bn(o, blk.conv3, blk.bn3, True, identity=idt)


def _res_conv_bn(x, conv, bn, relu, identity=None):
    """the same layer with autograd (trainable encoder)"""
    y, _z, _coef, stats = R.res_conv_bn(x, conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, identity,
                                        conv.stride[0], bn.training, float(bn.eps), relu)
    _bn_update(bn, stats.detach(), y)
    return y


def _res_bottleneck(blk, x):
    o = _res_conv_bn(x, blk.conv1, blk.bn1, True)
    o = _res_conv_bn(o, blk.conv2, blk.bn2, True)
    idt = x if blk.downsample is None else _res_conv_bn(x, blk.downsample[0], blk.downsample[1], False)
    return _res_conv_bn(o, blk.conv3, blk.bn3, True, identity=idt)


class DecoderBlock(nn.Module):
    """ConvTranspose2d(k2,s2) -> cat([up, skip]) -> basic_block — reference ResnetUnet.py:17-27."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.basic_block = basic_block(in_channels, out_channels)
        self.up_sample = nn.ConvTranspose2d(in_channels - out_channels, in_channels - out_channels, 2, 2)

    def forward(self, down, skip):
        x = R.conv_transpose2x2(down, self.up_sample.weight, self.up_sample.bias)
        return self.basic_block((x, skip))          # cat([x, skip], dim=1): up-sampled FIRST (ResnetUnet.py:25)


class ResNetUnet(nn.Module):
    """reference ResnetUnet.py:29-83"""

    def __init__(self, n_classes=1, freeze=True):
        super().__init__()
        import torchvision.models as models
        try:
            backbone = models.resnet50(weights=models.ResNet50_Weights.DEFAULT)
        except Exception as e:   # offline: no checkpoint download
            warnings.warn(f"ResNet-50 pretrained weights unavailable ({type(e).__name__}); using random init")
            backbone = models.resnet50(weights=None)

        self.encoder1 = nn.Sequential(backbone.conv1, backbone.bn1, backbone.relu)
        self.maxpool = backbone.maxpool
        self.encoder2 = backbone.layer1
        self.encoder3 = backbone.layer2
        self.encoder4 = backbone.layer3
        self.encoder5 = backbone.layer4
        self._frozen = bool(freeze)
        if freeze:
            self._freeze_backbone()

        self.decoder5 = DecoderBlock(2048 + 1024, 1024)
        self.decoder4 = DecoderBlock(1024 + 512, 512)
        self.decoder3 = DecoderBlock(512 + 256, 256)
        self.decoder2 = DecoderBlock(256 + 64, 64)
        self.decoder1 = nn.Sequential(
            nn.ConvTranspose2d(64, 32, kernel_size=2, stride=2),
            nn.Conv2d(32, 32, kernel_size=3, padding=1),
            nn.BatchNorm2d(32),
            nn.ReLU(inplace=True),
        )
        self.out = nn.Conv2d(32, n_classes, kernel_size=1)

    def _freeze_backbone(self):
        for layer in (self.encoder1, self.encoder2, self.encoder3, self.encoder4, self.encoder5):
            for param in layer.parameters():
                param.requires_grad = False

    def _encode_trainable(self, x):
        """freeze=False: the encoder is part of the autograd graph"""
        conv1, bn1 = self.encoder1[0], self.encoder1[1]
        e1, _z, _coef, stats, _xcol = R.res_stem(x, conv1.weight, bn1.weight, bn1.bias, bn1.running_mean,
                                                 bn1.running_var, bn1.training, float(bn1.eps))
        _bn_update(bn1, stats.detach(), e1)
        feats = [e1]
        t = R.res_maxpool3x3s2(e1)
        for layer in (self.encoder2, self.encoder3, self.encoder4, self.encoder5):
            for blk in layer:
                t = _res_bottleneck(blk, t)
            feats.append(t)
        return feats

    def _encode(self, x):
        enc_params = [p for layer in (self.encoder1, self.encoder2, self.encoder3, self.encoder4, self.encoder5)
                      for p in layer.parameters()]
        if torch.is_grad_enabled() and any(p.requires_grad for p in enc_params):
            return self._encode_trainable(x)
        with torch.no_grad():
            conv1, bn1 = self.encoder1[0], self.encoder1[1]
            z, stats = R.enc_stem(x, conv1.weight)
            n, h, w, _ = z.shape
            coef = ops.bn_finalize_(stats, n * h * w, bn1.weight, bn1.bias, bn1.running_mean, bn1.running_var,
                                    bn1.num_batches_tracked, bn1.training, float(bn1.momentum), float(bn1.eps))
            e1 = ops.bn_apply(z, coef, bn1.weight, bn1.bias, True, bn1.training)
            feats = [e1]
            t = R.enc_maxpool3x3s2(e1)
            for layer in (self.encoder2, self.encoder3, self.encoder4, self.encoder5):
                for blk in layer:
                    t = _bottleneck(blk, t)
                feats.append(t)
        return feats

    def forward(self, x):
        x = check_image(x)
        from ... import ops_fp32
        if ops_fp32.active(self):                 # fp32 parity mode (inference): b200seg.precision("fp32")
            return ops_fp32.resnet_unet(self, x)
        e1, e2, e3, e4, e5 = self._encode(x)
        d5 = self.decoder5(e5, e4)
        d4 = self.decoder4(d5, e3)
        d3 = self.decoder3(d4, e2)
        d2 = self.decoder2(d3, e1)
        d1 = R.conv_transpose2x2(d2, self.decoder1[0].weight, self.decoder1[0].bias)
        d1 = conv_bn_act(d1, self.decoder1[1], self.decoder1[2])
        return ops.head(d1, self.out.weight, self.out.bias)
