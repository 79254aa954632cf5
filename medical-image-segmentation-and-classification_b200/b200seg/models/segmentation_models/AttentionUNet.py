"""Drop-in for the reference's models/segmentation_models/AttentionUNet.py (same file name, class names,
constructor signatures, submodule names => identical state_dict keys; Appendix B of SURVEY.md).
forward() runs the b200seg sm_100a kernels; see b200seg/blocks.py for the block implementations."""
import torch
import torch.nn as nn

from ... import ops
from ...blocks import AttentionGate, UpConv, basic_block, check_image  # noqa: F401  (re-exported names)


class AttentionUNet(nn.Module):
    """Attention U-Net — reference AttentionUNet.py:56-121.  `in_channel` is stored but ignored (first block is
    hard-wired to 3 input channels), exactly like the reference (AttentionUNet.py:57-62)."""

    def __init__(self, in_channel=3, out_channel=1):
        super().__init__()
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.max_pool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.conv1 = basic_block(3, 64)
        self.conv2 = basic_block(64, 128)
        self.conv3 = basic_block(128, 256)
        self.conv4 = basic_block(256, 512)
        self.conv5 = basic_block(512, 1024)

        self.up5 = UpConv(1024, 512)
        self.att5 = AttentionGate(F_g=512, F_l=512, F_int=256)
        self.up_conv5 = basic_block(1024, 512)

        self.up4 = UpConv(512, 256)
        self.att4 = AttentionGate(F_g=256, F_l=256, F_int=128)
        self.up_conv4 = basic_block(512, 256)

        self.up3 = UpConv(256, 128)
        self.att3 = AttentionGate(F_g=128, F_l=128, F_int=64)
        self.up_conv3 = basic_block(256, 128)

        self.up2 = UpConv(128, 64)
        self.att2 = AttentionGate(F_g=64, F_l=64, F_int=32)
        self.up_conv2 = basic_block(128, 64)

        self.out = nn.Conv2d(64, out_channel, kernel_size=1, stride=1, padding=0)

    def features(self, x: torch.Tensor):
        """Everything up to the head, on internal NHWC bf16 activations."""
        # (pooled, skip): the pass-through output is what the decoder consumes, so the skip's decoder-side gradient is
        # added inside the pool-backward kernel (ops.maxpool2x2_pass)
        pool = ops.maxpool2x2_pass
        x1 = self.conv1._internal(x)
        p1, x1 = pool(x1)
        x2 = self.conv2(p1)
        p2, x2 = pool(x2)
        x3 = self.conv3(p2)
        p3, x3 = pool(x3)
        x4 = self.conv4(p3)
        p4, x4 = pool(x4)
        x5 = self.conv5(p4)

        d5 = self.up5(x5)
        a5, d5 = self.att5.gate_pass(g=d5, x=x4)      # cat((x4, d5), dim=1): skip first (ref :101)
        d5 = self.up_conv5((a5, d5))
        d4 = self.up4(d5)
        a4, d4 = self.att4.gate_pass(g=d4, x=x3)
        d4 = self.up_conv4((a4, d4))
        d3 = self.up3(d4)
        a3, d3 = self.att3.gate_pass(g=d3, x=x2)
        d3 = self.up_conv3((a3, d3))
        d2 = self.up2(d3)
        a2, d2 = self.att2.gate_pass(g=d2, x=x1)
        d2 = self.up_conv2((a2, d2))
        return d2

    def forward(self, x: torch.Tensor):
        from ... import ops_fp32
        if ops_fp32.active(self):                 # fp32 parity mode (inference): b200seg.precision("fp32")
            return ops_fp32.attention_unet(self, check_image(x))
        d2 = self.features(check_image(x))
        return ops.head(d2, self.out.weight, self.out.bias)    # raw logits, fp32 NCHW
