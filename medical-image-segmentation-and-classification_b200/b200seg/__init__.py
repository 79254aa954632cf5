"""b200seg — B200-native (sm_100a) forward/backward of the U-Net-family segmentation models of
bababyVN/medical-image-segmentation-and-classification, behind the reference's own nn.Module surface.

Layout
  _lib.py      ctypes binding of libb200seg.so (include/b200seg.h)
  kernels.py   tensor-level wrappers (pointers + dims + stream)
  ops.py       torch.library custom ops with autograd
  models/segmentation_models/{AttentionUNet,R2U_Net,R2AttU_Net,ResnetUnet}.py   drop-in modules
  utils/helpers.py   train() / iou() / get_seg_model() mirrors of the reference's utils/helpers.py
  ddp.py       bucketed NCCL gradient all-reduce
"""
__version__ = "0.1.0"


def __getattr__(name):
    # b200seg.precision("fp32") / set_precision / get_precision: the fp32 parity mode switch (ops_fp32.py), resolved
    # lazily so that importing the package stays free of torch.library / ctypes work
    if name in ("precision", "set_precision", "get_precision"):
        from . import ops_fp32
        return getattr(ops_fp32, name)
    raise AttributeError(f"module 'b200seg' has no attribute {name!r}")
