"""Drop-in for the reference's models/segmentation_models/R2AttU_Net.py: R2U_Net with attention gates on the skips
(same file / class / submodule names => identical state_dict keys; the reference file also defines an unused
`basic_block`, re-exported here for import compatibility)."""
import torch
import torch.nn as nn

from ... import ops
from ...blocks import (AttentionGate, Recurrent_block, RRCNN_block, UpConv, basic_block,  # noqa: F401
                       check_image)


class R2AttU_Net(nn.Module):
    """reference R2AttU_Net.py:88-158 (ctor default t=5, R2AttU_Net.py:89)."""

    def __init__(self, in_channels=3, out_channels=1, t=5):
        super().__init__()
        self.max_pool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.upsample = nn.Upsample(scale_factor=2)

        self.RRCNN1 = RRCNN_block(in_channels=in_channels, out_channels=64, t=t)
        self.RRCNN2 = RRCNN_block(in_channels=64, out_channels=128, t=t)
        self.RRCNN3 = RRCNN_block(in_channels=128, out_channels=256, t=t)
        self.RRCNN4 = RRCNN_block(in_channels=256, out_channels=512, t=t)
        self.RRCNN5 = RRCNN_block(in_channels=512, out_channels=1024, t=t)

        self.up5 = UpConv(in_channels=1024, out_channels=512)
        self.att5 = AttentionGate(F_g=512, F_l=512, F_int=256)
        self.up_RRCNN5 = RRCNN_block(in_channels=1024, out_channels=512, t=t)

        self.up4 = UpConv(in_channels=512, out_channels=256)
        self.att4 = AttentionGate(F_g=256, F_l=256, F_int=128)
        self.up_RRCNN4 = RRCNN_block(in_channels=512, out_channels=256, t=t)

        self.up3 = UpConv(in_channels=256, out_channels=128)
        self.att3 = AttentionGate(F_g=128, F_l=128, F_int=64)
        self.up_RRCNN3 = RRCNN_block(in_channels=256, out_channels=128, t=t)

        self.up2 = UpConv(in_channels=128, out_channels=64)
        self.att2 = AttentionGate(F_g=64, F_l=64, F_int=32)
        self.up_RRCNN2 = RRCNN_block(in_channels=128, out_channels=64, t=t)

        self.conv_1x1 = nn.Conv2d(64, out_channels, kernel_size=1, stride=1, padding=0)

    def features(self, x: torch.Tensor):
        # (pooled, skip): the pass-through output is what the decoder consumes, so the skip's decoder-side gradient is
        # added inside the pool-backward kernel (ops.maxpool2x2_pass)
        pool = ops.maxpool2x2_pass
        x1 = self.RRCNN1._internal(x if self.RRCNN1.conv_1x1.in_channels <= 4 else ops.to_nhwc(x))
        p1, x1 = pool(x1)
        x2 = self.RRCNN2(p1)
        p2, x2 = pool(x2)
        x3 = self.RRCNN3(p2)
        p3, x3 = pool(x3)
        x4 = self.RRCNN4(p3)
        p4, x4 = pool(x4)
        x5 = self.RRCNN5(p4)

        d5 = self.up5(x5)
        a5, d5 = self.att5.gate_pass(g=d5, x=x4)     # R2AttU_Net.py:137-139
        d5 = self.up_RRCNN5((a5, d5))
        d4 = self.up4(d5)
        a4, d4 = self.att4.gate_pass(g=d4, x=x3)
        d4 = self.up_RRCNN4((a4, d4))
        d3 = self.up3(d4)
        a3, d3 = self.att3.gate_pass(g=d3, x=x2)
        d3 = self.up_RRCNN3((a3, d3))
        d2 = self.up2(d3)
        a2, d2 = self.att2.gate_pass(g=d2, x=x1)
        d2 = self.up_RRCNN2((a2, d2))
        return d2

    def forward(self, x):
        from ... import ops_fp32
        if ops_fp32.active(self):                 # fp32 parity mode (inference): b200seg.precision("fp32")
            return ops_fp32.r2_net(self, check_image(x), gates=True)
        d2 = self.features(check_image(x))
        return ops.head(d2, self.conv_1x1.weight, self.conv_1x1.bias)
