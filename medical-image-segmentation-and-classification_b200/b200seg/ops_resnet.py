"""torch.library ops that only ResNetUnet needs (reference models/segmentation_models/ResnetUnet.py).

* encoder ops (torchvision resnet50 layers, ResnetUnet.py:32-43) are FORWARD-ONLY: the reference freezes the encoder
  by default (ResnetUnet.py:30,45-46,60-66), so autograd never records it.  They have no autograd formula; asking for
  gradients through them raises.
* conv_transpose2x2 = nn.ConvTranspose2d(k=2, s=2) (ResnetUnet.py:21,53) with full backward: four 1x1 tcgen05 GEMMs
  whose TMA-store epilogue scatters into the 2x grid (pixel shuffle), dgrad = 2x2/stride-2 conv of dY, wgrad = tcgen05
  MN-major GEMM between the coarse input and the strided fine-grid dY.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor
from torch.library import custom_op

from . import kernels as K

_F64 = torch.float64


def _c(t):
    return t if t is None or t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------------------------------------
# encoder (forward only)
# ----------------------------------------------------------------------------------------------------------
@custom_op("b200seg::enc_stem", mutates_args=())
def enc_stem(x: Tensor, weight: Tensor) -> Tuple[Tensor, Tensor]:
    """backbone.conv1: 7x7/s2/p3, 3->64, no bias, from the fp32 NCHW image; returns (z, stats)"""
    x4 = K.image_to_nhwc4(_c(x))
    z = K.stem7x7_fprop(x4, K.pack_small_weight(weight))
    stats = torch.zeros((2, weight.shape[0]), dtype=_F64, device=x.device)
    K.channel_stats(z, stats)
    return z, stats


@enc_stem.register_fake
def _(x, weight):
    n, _, h, w = x.shape
    return x.new_empty((n, h // 2, w // 2, weight.shape[0]), dtype=torch.bfloat16), x.new_empty((2, 64), dtype=_F64)


@custom_op("b200seg::enc_maxpool3x3s2", mutates_args=())
def enc_maxpool3x3s2(x: Tensor) -> Tensor:
    """backbone.maxpool = MaxPool2d(3, 2, 1)"""
    return K.maxpool3x3s2_fwd(_c(x))


@enc_maxpool3x3s2.register_fake
def _(x):
    n, h, w, c = x.shape
    return x.new_empty((n, h // 2, w // 2, c))


@custom_op("b200seg::enc_conv_bn", mutates_args=())
def enc_conv_bn(x: Tensor, weight: Tensor, gamma: Tensor, beta: Tensor, running_mean: Tensor, running_var: Tensor,
                identity: Optional[Tensor], stride: int, training: bool, eps: float,
                relu: bool) -> Tuple[Tensor, Tensor]:
    """bias-free conv (1x1 or 3x3, stride 1|2) -> BatchNorm (batch stats when training) -> [+identity] -> [ReLU];
    returns (y, stats) — the caller applies the running-stat update."""
    cout, cin, k, _ = weight.shape
    wf, _ = K.packed(weight)
    stats = torch.zeros((2, cout), dtype=_F64, device=x.device) if training else \
        torch.empty((0,), dtype=_F64, device=x.device)
    z = K.conv_igemm(_c(x), wf, cout, k, stats=stats if training else None, stride=stride)
    n, h, w, _ = z.shape
    if training:
        coef = K.bn_finalize(stats, n * h * w, gamma, beta, eps, 0.0, None, None, None)
    else:
        coef = K.bn_eval_coeffs(gamma, beta, running_mean, running_var, eps)
    return K.bn_apply(z, coef, relu=relu, addend=_c(identity)), stats


@enc_conv_bn.register_fake
def _(x, weight, gamma, beta, rm, rv, identity, stride, training, eps, relu):
    n, h, w, _ = x.shape
    cout = weight.shape[0]
    return (x.new_empty((n, h // stride, w // stride, cout)),
            x.new_empty((2, cout) if training else (0,), dtype=_F64))


# ----------------------------------------------------------------------------------------------------------
# ConvTranspose2d(kernel 2, stride 2) with backward
# ----------------------------------------------------------------------------------------------------------
@custom_op("b200seg::conv_transpose2x2", mutates_args=())
def conv_transpose2x2(x: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    """y[n, 2h+i, 2w+j, co] = sum_ci x[n,h,w,ci] W[ci,co,i,j] + b[co];  weight is [Cin, Cout, 2, 2]"""
    cin, cout = weight.shape[0], weight.shape[1]
    x = _c(x)
    n, h, w, _ = x.shape
    wf, _ = K.packed(weight, "convT_f")                                          # [4][Cout][Cin]
    y = K.new_act(n, 2 * h, 2 * w, cout, x.device)
    for i in range(2):
        for j in range(2):
            t = i * 2 + j
            K.conv_igemm(x, wf[t:t + 1], cout, 1, bias=bias, out=y, out_mul=2, out_off=(i, j))
    return y


@conv_transpose2x2.register_fake
def _(x, weight, bias):
    n, h, w, _ = x.shape
    return x.new_empty((n, 2 * h, 2 * w, weight.shape[1]))


@custom_op("b200seg::conv_transpose2x2_bwd", mutates_args=())
def conv_transpose2x2_bwd(dy: Tensor, x: Tensor, weight: Tensor, need_dx: bool,
                          has_bias: bool) -> Tuple[Tensor, Tensor, Tensor]:
    cin, cout = weight.shape[0], weight.shape[1]
    dy = _c(dy)
    dev = dy.device
    if need_dx:
        # dx[n,h,w,ci] = sum_{i,j,co} dy[n,2h+i,2w+j,co] W[ci,co,i,j]: a 2x2 / stride-2 conv of dy whose
        # "output channels" are ci — W already has the [rows=ci][K=co][2][2] shape the fprop packing expects
        wd, _ = K.packed(weight, "convT_d")                                       # [4][Cin][Cout]
        dx = K.conv_igemm(dy, wd, cin, 2, stride=2, dgrad=True)
    else:
        dx = torch.empty((0,), dtype=torch.bfloat16, device=dev)
    with K.wgrad_stream(x, dy, allow=K.grad_is_stolen(weight)):
        dw = K.conv_wgrad(x, dy, 2, x_stride=2, out=K.grad_slot(weight, (cin, 4, cout)))   # [Cin][4][Cout]
    db = K.channel_sum(dy) if has_bias else torch.empty((0,), device=dev)
    return dx, dw, db


def _ct_setup(ctx, inputs, output):
    x, weight, bias = inputs
    ctx.save_for_backward(_c(x), weight)
    ctx.has_bias = bias is not None


def _ct_backward(ctx, dy):
    x, weight = ctx.saved_tensors
    cin, cout = weight.shape[0], weight.shape[1]
    dx, dw, db = conv_transpose2x2_bwd(dy, x, weight, ctx.needs_input_grad[0], ctx.has_bias)
    return (dx if ctx.needs_input_grad[0] else None,
            dw.view(cin, 2, 2, cout).permute(0, 3, 1, 2),
            db if ctx.has_bias else None)


conv_transpose2x2.register_autograd(_ct_backward, setup_context=_ct_setup)
