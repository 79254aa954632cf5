"""torch.library ops that only ResNetUnet needs (reference models/segmentation_models/ResnetUnet.py).

* frozen-encoder ops (enc_*): the reference freezes the torchvision resnet50 encoder by default (ResnetUnet.py:30,45-46,
  60-66), so autograd never records it; these are forward-only and run under no_grad.
* trainable-encoder ops (res_*): ResNetUnet(freeze=False) (ResnetUnet.py:29-30) trains the encoder too — the 7x7 stem as
  an im2col GEMM on tcgen05 (fprop + wgrad), MaxPool2d(3,2,1) with backward, and bias-free conv (1x1 / 3x3, stride 1|2)
  -> BatchNorm -> [+identity] -> [ReLU] as ONE autograd node (the strided convolutions' dgrad / wgrad run on the
  stride-1 tensor-core kernels over a zero-inserted dY).
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


# ----------------------------------------------------------------------------------------------------------
# trainable encoder: ResNetUnet(freeze=False)
# ----------------------------------------------------------------------------------------------------------
@custom_op("b200seg::res_stem", mutates_args=())
def res_stem(x: Tensor, weight: Tensor, gamma: Tensor, beta: Tensor, running_mean: Tensor, running_var: Tensor,
             training: bool, eps: float) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """backbone.conv1 (7x7/s2/p3, no bias) -> bn1 -> ReLU from the fp32 NCHW image, as an im2col GEMM on tcgen05;
    returns (y, z, coef, stats, xcol)"""
    cout, cin, k, _ = weight.shape
    xcol = K.stem_im2col(_c(x), k, 2, k // 2, K.stem_cols(weight))
    wf, _ = K.packed(weight, "stem")
    stats = torch.zeros((2, cout), dtype=_F64, device=x.device) if training else \
        torch.empty((0,), dtype=_F64, device=x.device)
    z = K.conv_igemm(xcol, wf, cout, 1, stats=stats if training else None)
    n, h, w, _ = z.shape
    if training:
        coef = K.bn_finalize(stats, n * h * w, gamma, beta, eps, 0.0, None, None, None)
    else:
        coef = K.bn_eval_coeffs(gamma, beta, running_mean, running_var, eps)
    return K.bn_apply(z, coef, relu=True), z, coef, stats, xcol


@res_stem.register_fake
def _(x, weight, gamma, beta, rm, rv, training, eps):
    n, _, h, w = x.shape
    cout = weight.shape[0]
    y = x.new_empty((n, h // 2, w // 2, cout), dtype=torch.bfloat16)
    cols = 32 if weight.shape[2] == 3 else (weight.shape[2] ** 2 * weight.shape[1] + 7) // 8 * 8
    return (y, torch.empty_like(y), x.new_empty((4, cout)), x.new_empty((2, cout) if training else (0,), dtype=_F64),
            x.new_empty((n, h // 2, w // 2, cols), dtype=torch.bfloat16))


@custom_op("b200seg::res_stem_bwd", mutates_args=())
def res_stem_bwd(dy: Tensor, xcol: Tensor, weight: Tensor, z: Tensor, coef: Tensor, gamma: Tensor,
                 training: bool) -> Tuple[Tensor, Tensor, Tensor]:
    cout, cin, k, _ = weight.shape
    dz, dgamma, dbeta = K.bn_bwd(_c(dy), z, coef, gamma, relu=True, training=training)
    dwk = K.stem_weight_grad(K.conv_wgrad(dz, xcol, 1), cin, taps=k * k)            # [cout, taps, cin]
    return dwk.reshape(cout, k, k, cin).permute(0, 3, 1, 2).contiguous(), dgamma, dbeta


def _res_stem_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    _x, weight, gamma, _b, _rm, _rv, ctx.training, _eps = inputs
    _y, z, coef, _stats, xcol = output
    ctx.save_for_backward(xcol, weight, z, coef, gamma)


def _res_stem_backward(ctx, dy, *_unused):
    xcol, weight, z, coef, gamma = ctx.saved_tensors
    dw, dgamma, dbeta = res_stem_bwd(dy, xcol, weight, z, coef, gamma, ctx.training)
    return None, dw, dgamma, dbeta, None, None, None, None


res_stem.register_autograd(_res_stem_backward, setup_context=_res_stem_setup)


@custom_op("b200seg::res_maxpool3x3s2", mutates_args=())
def res_maxpool3x3s2(x: Tensor) -> Tensor:
    """backbone.maxpool = MaxPool2d(3, 2, 1), with backward"""
    return K.maxpool3x3s2_fwd(_c(x))


@res_maxpool3x3s2.register_fake
def _(x):
    n, h, w, c = x.shape
    return x.new_empty((n, h // 2, w // 2, c))


@custom_op("b200seg::res_maxpool3x3s2_bwd", mutates_args=())
def res_maxpool3x3s2_bwd(dy: Tensor, x: Tensor) -> Tensor:
    return K.maxpool3x3s2_bwd(_c(dy), x)


res_maxpool3x3s2.register_autograd(lambda ctx, dy: res_maxpool3x3s2_bwd(dy, ctx.saved_tensors[0]),
                                   setup_context=lambda ctx, inputs, output: ctx.save_for_backward(_c(inputs[0])))


@custom_op("b200seg::res_conv_bn", mutates_args=())
def res_conv_bn(x: Tensor, weight: Tensor, gamma: Tensor, beta: Tensor, running_mean: Tensor, running_var: Tensor,
                identity: Optional[Tensor], stride: int, training: bool, eps: float,
                relu: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """bias-free conv (1x1 or 3x3, stride 1|2) -> BatchNorm -> [+identity] -> [ReLU] with full backward;
    returns (y, z, coef, stats)"""
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
    return K.bn_apply(z, coef, relu=relu, addend=_c(identity)), z, coef, stats


@res_conv_bn.register_fake
def _(x, weight, gamma, beta, rm, rv, identity, stride, training, eps, relu):
    n, h, w, _ = x.shape
    cout = weight.shape[0]
    y = x.new_empty((n, h // stride, w // stride, cout))
    return (y, torch.empty_like(y), x.new_empty((4, cout), dtype=torch.float32),
            x.new_empty((2, cout) if training else (0,), dtype=_F64))


@custom_op("b200seg::res_conv_bn_bwd", mutates_args=())
def res_conv_bn_bwd(dy: Tensor, x: Tensor, weight: Tensor, z: Tensor, coef: Tensor, gamma: Tensor,
                    y: Optional[Tensor], stride: int, relu: bool, training: bool,
                    need_dx: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """-> (dx, dw [Cout, taps, Cin], dgamma, dbeta, d_identity); y is given iff the node had an identity input"""
    cout, cin, k, _ = weight.shape
    dev = dy.device
    dy = _c(dy)
    if y is not None:
        # relu(bn(z) + identity): the ReLU mask depends on the identity too, so it is taken from the saved OUTPUT
        g = K.relu_mask(dy, y) if relu else dy
        dz, dgamma, dbeta = K.bn_bwd(g, z, coef, gamma, relu=False, training=training)
        d_idt = g
    else:
        dz, dgamma, dbeta = K.bn_bwd(dy, z, coef, gamma, relu=relu, training=training)
        d_idt = torch.empty((0,), dtype=torch.bfloat16, device=dev)
    # a stride-2 convolution's dY lives on the coarse grid: put it on the input grid (zeros in between) and the
    # ordinary stride-1 dgrad / wgrad kernels compute exactly the strided gradients
    dzu = K.zero_insert2x(dz) if stride == 2 else dz
    if need_dx:
        _, wd = K.packed(weight, want_dgrad=True)
        dx = K.conv_igemm(dzu, wd, cin, k, dgrad=True)
    else:
        dx = torch.empty((0,), dtype=torch.bfloat16, device=dev)
    with K.wgrad_stream(dzu, x, allow=K.grad_is_stolen(weight)):
        dw = K.conv_wgrad(dzu, x, k, out=K.grad_slot(weight, (cout, k * k, cin)))
    return dx, dw, dgamma, dbeta, d_idt


def _res_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    x, weight, gamma, _b, _rm, _rv, identity, ctx.stride, ctx.training, _eps, ctx.relu = inputs
    y, z, coef, _stats = output
    ctx.has_identity = identity is not None
    ctx.save_for_backward(_c(x), weight, z, coef, gamma, y if ctx.has_identity else None)


def _res_backward(ctx, dy, *_unused):
    x, weight, z, coef, gamma, y = ctx.saved_tensors
    need = ctx.needs_input_grad
    dx, dw, dgamma, dbeta, d_idt = res_conv_bn_bwd(dy, x, weight, z, coef, gamma, y, ctx.stride, ctx.relu, ctx.training,
                                                   bool(need[0]))
    cout, cin, k, _ = weight.shape
    return (dx if need[0] else None, dw.view(cout, k, k, cin).permute(0, 3, 1, 2), dgamma, dbeta, None, None,
            d_idt if (ctx.has_identity and need[6]) else None, None, None, None, None)


res_conv_bn.register_autograd(_res_backward, setup_context=_res_setup)
