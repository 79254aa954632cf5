"""The whole AttentionGate (AttentionUNet.py:29-54 / R2AttU_Net.py:61-86) as ONE autograd node.

forward : W_g, W_x 1x1 GEMMs on tcgen05 (BN statistics in their epilogues) -> finalize x2 -> gate_psi_fwd (BN, BN, add,
          ReLU, psi dot, statistics of q) -> finalize -> gate_apply_fwd (BN, sigmoid, multiply)
backward: gate_apply_bwd -> gate_psi_bwd_reduce -> gate_psi_bwd_apply (also the two conv-bias gradients) ->
          dgrad(W_g), dgrad(W_x) with the direct x-gradient added in the GEMM epilogue -> wgrad x2
Compared with composing conv2d + gate_mid ops this removes one full-size gradient accumulation per gate (x receives a
single gradient) and the separate bias-gradient reductions.  Functional: the three BatchNorms' running statistics are
updated by the caller from the returned (sum, sumsq) buffers.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
from torch import Tensor
from torch.library import custom_op

from . import kernels as K
from .ops import _c, _dw_as_param_grad, bn_update_running_

_F64 = torch.float64


@custom_op("b200seg::attention_gate", mutates_args=())
def attention_gate(g: Tensor, x: Tensor, wg: Tensor, bg: Tensor, wx: Tensor, bx: Tensor,
                   gamma_g: Tensor, beta_g: Tensor, rm_g: Tensor, rv_g: Tensor,
                   gamma_x: Tensor, beta_x: Tensor, rm_x: Tensor, rv_x: Tensor,
                   wpsi: Tensor, bpsi: Tensor, gamma_1: Tensor, beta_1: Tensor, rm_1: Tensor, rv_1: Tensor,
                   training: bool, eps: float
                   ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    g, x = _c(g), _c(x)
    n, h, w, _ = x.shape
    npix = n * h * w
    fint = wg.shape[0]
    dev = x.device

    def stats_buf():
        # (several of these are outputs of one op: custom-op outputs may not share a storage, so no arena slices here)
        return torch.zeros((2, fint), dtype=_F64, device=dev) if training else torch.empty((0,), dtype=_F64, device=dev)

    stats_g, stats_x = stats_buf(), stats_buf()
    wgf, _ = K.packed(wg)
    wxf, _ = K.packed(wx)
    g1p = K.conv_igemm(g, wgf, fint, 1, bias=bg, stats=stats_g if training else None)
    x1p = K.conv_igemm(x, wxf, fint, 1, bias=bx, stats=stats_x if training else None)
    if training:
        coef_g = K.bn_finalize(stats_g, npix, gamma_g, beta_g, eps, 0.0, None, None, None)
        coef_x = K.bn_finalize(stats_x, npix, gamma_x, beta_x, eps, 0.0, None, None, None)
    else:
        coef_g = K.bn_eval_coeffs(gamma_g, beta_g, rm_g, rv_g, eps)
        coef_x = K.bn_eval_coeffs(gamma_x, beta_x, rm_x, rv_x, eps)
    q, qstats = K.gate_psi_fwd(g1p, x1p, coef_g, coef_x, wpsi, bpsi)
    if training:
        coef_1 = K.bn_finalize(qstats, npix, gamma_1, beta_1, eps, 0.0, None, None, None)
    else:
        coef_1 = K.bn_eval_coeffs(gamma_1, beta_1, rm_1, rv_1, eps)
    out, psi = K.gate_apply_fwd(x, q, coef_1)
    return out, q, psi, g1p, x1p, coef_g, coef_x, coef_1, stats_g, stats_x, qstats


@attention_gate.register_fake
def _(g, x, wg, *rest):
    n, h, w, _ = x.shape
    fint = wg.shape[0]
    training = rest[-2]
    f32 = torch.float32
    st = (2, fint) if training else (0,)
    return (torch.empty_like(x), x.new_empty((n, h, w)), x.new_empty((n, h, w)), x.new_empty((n, h, w, fint)),
            x.new_empty((n, h, w, fint)), x.new_empty((4, fint), dtype=f32), x.new_empty((4, fint), dtype=f32),
            x.new_empty((4, 1), dtype=f32), x.new_empty(st, dtype=_F64), x.new_empty(st, dtype=_F64),
            x.new_empty((2,), dtype=_F64))


@custom_op("b200seg::attention_gate_bwd", mutates_args=())
def attention_gate_bwd(dout: Tensor, g: Tensor, x: Tensor, wg: Tensor, wx: Tensor, psi: Tensor, q: Tensor,
                       g1p: Tensor, x1p: Tensor, coef_g: Tensor, gamma_g: Tensor, coef_x: Tensor, gamma_x: Tensor,
                       coef_1: Tensor, gamma_1: Tensor, wpsi: Tensor, training: bool, need_dg: bool, need_dx: bool,
                       dg_add: Optional[Tensor] = None
                       ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    fint = wg.shape[0]
    dev = dout.device
    dx, dsig, sums1 = K.gate_apply_bwd(_c(dout), x, psi, q, coef_1)          # dx = dout * psi (direct path)
    dg1p, dx1p, dgb, dbn1, dwpsi, dbpsi, dbias = K.gate_psi_bwd(dsig, sums1, q, g1p, x1p, coef_g, gamma_g, coef_x,
                                                                gamma_x, coef_1, gamma_1, wpsi, training=training)
    if need_dg:
        _, wgd = K.packed(wg, want_dgrad=True)
        # dg_add: the gradient g received from its other consumer (the concat), added in the GEMM epilogue
        dg = K.conv_igemm(dg1p, wgd, g.shape[3], 1, dgrad=True, addend=_c(dg_add) if dg_add is not None else None)
    else:
        dg = torch.empty((0,), dtype=torch.bfloat16, device=dev)
    if need_dx:
        _, wxd = K.packed(wx, want_dgrad=True)
        K.conv_igemm(dx1p, wxd, x.shape[3], 1, addend=dx, out=dx, dgrad=True)     # dx += dx1p . W_x (epilogue add)
    # gradient slots (data-parallel buckets): both weights usually live in ONE flat bucket, and two outputs of a custom
    # op may not share a storage — so slot-resident gradients are written as a side effect and returned as empty
    # placeholders; the autograd formula below hands autograd fresh views of the slots
    slot_g, slot_x = K.grad_slot(wg, (fint, 1, g.shape[3])), K.grad_slot(wx, (fint, 1, x.shape[3]))
    with K.wgrad_stream(dg1p, dx1p, g, x, allow=K.grad_is_stolen(wg) and K.grad_is_stolen(wx)):
        dwg = K.conv_wgrad(dg1p, g, 1, out=slot_g)
        dwx = K.conv_wgrad(dx1p, x, 1, out=slot_x)
    if slot_g is not None:
        dwg = torch.empty((0,), device=dev)
    if slot_x is not None:
        dwx = torch.empty((0,), device=dev)
    return dg, dx, dwg, dwx, dgb, dbn1, dwpsi, dbpsi, dbias


def _setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    (g, x, wg, _bg, wx, _bx, gamma_g, _b, _rm, _rv, gamma_x, _b2, _rm2, _rv2, wpsi, _bpsi, gamma_1, _b1, _rm1, _rv1,
     ctx.training, _eps) = inputs
    out, q, psi, g1p, x1p, coef_g, coef_x, coef_1, *_ = output
    ctx.save_for_backward(_c(g), _c(x), wg, wx, psi, q, g1p, x1p, coef_g, gamma_g, coef_x, gamma_x, coef_1, gamma_1,
                          wpsi)


def _backward(ctx, dout, *_unused):
    g, x, wg, wx, psi, q, g1p, x1p, coef_g, gamma_g, coef_x, gamma_x, coef_1, gamma_1, wpsi = ctx.saved_tensors
    need = ctx.needs_input_grad
    dg, dx, dwg, dwx, dgb, dbn1, dwpsi, dbpsi, dbias = attention_gate_bwd(
        dout, g, x, wg, wx, psi, q, g1p, x1p, coef_g, gamma_g, coef_x, gamma_x, coef_1, gamma_1, wpsi, ctx.training,
        bool(need[0]), bool(need[1]))
    fint = wg.shape[0]
    if dwg.numel() == 0:
        dwg = K.grad_slot(wg, (fint, 1, wg.shape[1]))
    if dwx.numel() == 0:
        dwx = K.grad_slot(wx, (fint, 1, wx.shape[1]))
    return (dg if need[0] else None, dx if need[1] else None,
            _dw_as_param_grad(dwg, wg), dbias[0], _dw_as_param_grad(dwx, wx), dbias[1],
            dgb[0], dgb[1], None, None, dgb[2], dgb[3], None, None,
            dwpsi.view_as(wpsi), dbpsi, dbn1[0:1], dbn1[1:2], None, None, None, None)


attention_gate.register_autograd(_backward, setup_context=_setup)


class _GatePass(torch.autograd.Function):
    """(gate(g, x), g): the UpConv output g feeds the gate AND the concat (AttentionUNet.py:99-101).  Handing the concat
    the pass-through output makes this node g's only consumer, so the concat-side gradient arrives here and rides the
    W_g dgrad's epilogue as its addend instead of autograd's separate accumulation pass over the full-size tensor."""

    @staticmethod
    def forward(ctx, g, x, *rest):
        (wg, _bg, wx, _bx, gamma_g, _b, _rm, _rv, gamma_x, _b2, _rm2, _rv2, wpsi, _bpsi, gamma_1, _b1, _rm1, _rv1,
         training, _eps) = rest
        out, q, psi, g1p, x1p, coef_g, coef_x, coef_1, stats_g, stats_x, qstats = attention_gate(g, x, *rest)
        ctx.save_for_backward(_c(g), _c(x), wg, wx, psi, q, g1p, x1p, coef_g, gamma_g, coef_x, gamma_x, coef_1, gamma_1,
                              wpsi)
        ctx.training = training
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(stats_g, stats_x, qstats)
        return out, g, stats_g, stats_x, qstats

    @staticmethod
    def backward(ctx, dout, dg_pass, *_unused):
        g, x, wg, wx, psi, q, g1p, x1p, coef_g, gamma_g, coef_x, gamma_x, coef_1, gamma_1, wpsi = ctx.saved_tensors
        need = ctx.needs_input_grad
        if dout is None:                      # the gate's own output was not used: only the pass-through gradient flows
            return (dg_pass,) + (None,) * 21
        dg, dx, dwg, dwx, dgb, dbn1, dwpsi, dbpsi, dbias = attention_gate_bwd(
            dout, g, x, wg, wx, psi, q, g1p, x1p, coef_g, gamma_g, coef_x, gamma_x, coef_1, gamma_1, wpsi, ctx.training,
            bool(need[0]), bool(need[1]), dg_pass if need[0] else None)
        fint = wg.shape[0]
        if dwg.numel() == 0:
            dwg = K.grad_slot(wg, (fint, 1, wg.shape[1]))
        if dwx.numel() == 0:
            dwx = K.grad_slot(wx, (fint, 1, wx.shape[1]))
        return (dg if need[0] else None, dx if need[1] else None,
                _dw_as_param_grad(dwg, wg), dbias[0], _dw_as_param_grad(dwx, wx), dbias[1],
                dgb[0], dgb[1], None, None, dgb[2], dgb[3], None, None,
                dwpsi.view_as(wpsi), dbpsi, dbn1[0:1], dbn1[1:2], None, None, None, None)


_GATE_PASS = os.environ.get("B200SEG_GATE_PASS", "1") != "0"      # 0: autograd accumulates g's two gradients (A/B switch)


def attention_gate_module(gate, g: Tensor, x: Tensor, with_pass: bool = False):
    """gate: an AttentionGate module (W_g, W_x, psi Sequentials with the reference's layout).
    with_pass=True returns (out, g_pass): g_pass is g as the tensor the concat should consume (see _GatePass)."""
    bn_g, bn_x, bn_1 = gate.W_g[1], gate.W_x[1], gate.psi[1]
    if os.environ.get("B200SEG_FOLD_BN", "1") != "0" and os.environ.get("B200SEG_GATE_FUSED", "1") != "0":
        from . import ops_infer
        if ops_infer.inference_mode(bn_g) and ops_infer.gate_fusable(gate, g, x):
            out = ops_infer.attention_gate_fused(gate, g, x)           # eval + no_grad: the whole gate in ONE kernel
            return (out, g) if with_pass else out
    training = bn_g.training
    args = (g, x, gate.W_g[0].weight, gate.W_g[0].bias, gate.W_x[0].weight, gate.W_x[0].bias,
            bn_g.weight, bn_g.bias, bn_g.running_mean, bn_g.running_var,
            bn_x.weight, bn_x.bias, bn_x.running_mean, bn_x.running_var,
            gate.psi[0].weight, gate.psi[0].bias, bn_1.weight, bn_1.bias, bn_1.running_mean,
            bn_1.running_var, training, float(bn_1.eps))
    g_pass = g
    if with_pass and _GATE_PASS and torch.is_grad_enabled() and g.requires_grad:
        out, g_pass, stats_g, stats_x, qstats = _GatePass.apply(*args)
    else:
        res = attention_gate(*args)
        out, stats_g, stats_x, qstats = res[0], res[8], res[9], res[10]
    if training:
        n, h, w, _ = x.shape
        for bn, st in ((bn_g, stats_g), (bn_x, stats_x), (bn_1, qstats)):
            bn_update_running_(st.detach(), n * h * w, float(bn.momentum), bn.running_mean, bn.running_var,
                               bn.num_batches_tracked)
    return (out, g_pass) if with_pass else out
