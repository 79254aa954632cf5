"""Inference fast path (reference callers utils/pipeline.py:340-357, utils/tester.py:264-289,
validation loop utils/helpers.py:345-360): with BatchNorm in eval mode and autograd off, every
[Conv2d -> BatchNorm2d -> ReLU] collapses into ONE tcgen05 convolution whose weights carry gamma/sqrt(var+eps) and
whose epilogue adds the folded bias and applies the ReLU (Appendix D.3).  UpConv uses the 16-tap folded weights, so
the up-sampled tensor is never built either.  Forward-only ops: no autograd formula.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor
from torch.library import custom_op

from . import kernels as K
from .ops import _PHASES, _c


@custom_op("b200seg::conv_bn_act_infer", mutates_args=())
def conv_bn_act_infer(x0: Tensor, x1: Optional[Tensor], weight: Tensor, bias: Optional[Tensor], gamma: Tensor,
                      beta: Tensor, running_mean: Tensor, running_var: Tensor, eps: float, relu: bool,
                      addend: Optional[Tensor]) -> Tensor:
    """y = act(conv(x; W*s) + (b*s + t)) [+ addend]  with s = gamma/sqrt(running_var+eps), t = beta - running_mean*s"""
    cout, cin, k, _ = weight.shape
    coef = K.bn_eval_coeffs(gamma, beta, running_mean, running_var, eps)
    wf, bf = K.pack_weights_folded(weight, bias, coef)
    return K.conv_igemm(_c(x0), wf, cout, k, x1=_c(x1), bias=bf, relu=relu, addend=_c(addend),
                        add_after_act=addend is not None)


@conv_bn_act_infer.register_fake
def _(x0, x1, weight, bias, gamma, beta, rm, rv, eps, relu, addend):
    n, h, w, _ = x0.shape
    return x0.new_empty((n, h, w, weight.shape[0]))


@custom_op("b200seg::upconv_bn_act_infer", mutates_args=())
def upconv_bn_act_infer(x: Tensor, weight: Tensor, bias: Optional[Tensor], gamma: Tensor, beta: Tensor,
                        running_mean: Tensor, running_var: Tensor, eps: float, relu: bool) -> Tensor:
    cout = weight.shape[0]
    x = _c(x)
    n, h, w, _ = x.shape
    coef = K.bn_eval_coeffs(gamma, beta, running_mean, running_var, eps)
    wf, bf = K.pack_weights_folded(weight, bias, coef, upfold=True)
    y = K.new_act(n, 2 * h, 2 * w, cout, x.device)
    for ph, (a, b) in enumerate(_PHASES):
        K.conv_igemm(x, wf[ph], cout, 2, bias=bf, relu=relu, out=y, out_mul=2, out_off=(a, b), pad=(1 - a, 1 - b))
    return y


@upconv_bn_act_infer.register_fake
def _(x, weight, bias, gamma, beta, rm, rv, eps, relu):
    n, h, w, _ = x.shape
    return x.new_empty((n, 2 * h, 2 * w, weight.shape[0]))


def inference_mode(bn: torch.nn.BatchNorm2d) -> bool:
    """folded path applies: eval-mode BN with running statistics and no autograd recording"""
    return (not bn.training) and bn.running_mean is not None and not torch.is_grad_enabled()
