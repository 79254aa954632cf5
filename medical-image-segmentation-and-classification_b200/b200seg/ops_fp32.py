"""fp32 parity mode: the INFERENCE path of the four models with fp32 activation storage and fp32 accumulation
(BASELINE.json north_star "fp32-accumulate mode within 1e-4"; csrc/fp32.cu).

The reference runs its segmentation model in fp32 at inference time (utils/pipeline.py:340-357: no autocast, then
`sigmoid(logits) > 0.5`); under bf16 storage logits move by ~1e-3 and threshold pixels can flip.  With

    with b200seg.precision("fp32"):        # or b200seg.set_precision("fp32")
        logits = model(x)                  # model.eval(), torch.no_grad()

every model's forward() routes here: each [Conv2d -> BatchNorm2d(eval) -> ReLU] is ONE fp32 convolution with the
BatchNorm folded into weights / bias, the concat is elided (two K sources), Recurrent_block's `x + x1` rides the conv
epilogue, the attention gate's tail is one kernel.  Forward / eval mode only: training stays on the bf16 tensor-core
path (the reference itself trains under autocast, helpers.py:321).

Parity (tests/test_gpu_fp32.py): op level <= 1e-5, eval end-to-end logits <= 1e-4 vs the fp64 oracle for all four
models, and the uint8 masks of predict_mask equal to the real reference's fp32 masks stored in tests/golden/.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import weakref

import torch

from . import _lib
from . import kernels as K
from ._lib import F32ConvArgs, call
from .kernels import _p, _stream

_STATE = {"precision": "bf16"}


def set_precision(mode: str) -> None:
    if mode not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' (tensor-core path) or 'fp32' (parity mode, inference only)")
    _STATE["precision"] = mode


def get_precision() -> str:
    return _STATE["precision"]


@contextlib.contextmanager
def precision(mode: str):
    old = _STATE["precision"]
    set_precision(mode)
    try:
        yield
    finally:
        _STATE["precision"] = old


def active(module: torch.nn.Module) -> bool:
    """fp32 mode applies to this forward() call; raises if the mode is on but the call is not an inference call"""
    if _STATE["precision"] != "fp32":
        return False
    if module.training or torch.is_grad_enabled():
        raise RuntimeError("b200seg fp32 parity mode covers the inference path only: call model.eval() and run under "
                           "torch.no_grad() (training uses the bf16 tensor-core path)")
    return True


# ----------------------------------------------------------------------------------------------------------
# kernels wrappers (NHWC fp32 tensors [N, H, W, C])
# ----------------------------------------------------------------------------------------------------------
def _nhwc(t):
    assert t.dtype == torch.float32 and t.dim() == 4 and t.is_cuda and t.stride(3) == 1
    n, h, w, c = t.shape
    ld = t.stride(2)
    assert t.stride(1) == w * ld and t.stride(0) == h * w * ld
    return n, h, w, c, ld


_PACKED = {}


def _packed(weight, bias, bn, transposed=False):
    """([taps][cin][cout] fp32 with eval-mode BN folded, [cout] bias) cached per weight; invalidated by any update of
    the tensors involved (torch version counters, or the fused optimizer's parameter epoch)"""
    tensors = (weight, bias) + ((bn.weight, bn.bias, bn.running_mean, bn.running_var) if bn is not None else ())
    stamp = tuple((t.data_ptr(), t._version) if t is not None else None for t in tensors) + (K.param_epoch(),)
    key = (weight.data_ptr(), transposed)
    hit = _PACKED.get(key)
    if hit is not None and hit[0]() is weight and hit[1] == stamp:
        return hit[2], hit[3]
    view = weight.permute(1, 0, 2, 3) if transposed else weight      # ConvTranspose2d stores [Cin, Cout, k, k]
    cout, cin, kh, kw = view.shape
    assert kh == kw
    dev = weight.device
    if bn is not None:
        coef = K.bn_eval_coeffs(bn.weight, bn.bias, bn.running_mean, bn.running_var, float(bn.eps))
        scale, shift = coef[2], coef[3]
    else:
        scale = shift = None
    wp = torch.empty((kh * kw, cin, cout), dtype=torch.float32, device=dev)
    need_bias = bias is not None or bn is not None
    bp = torch.empty((cout,), dtype=torch.float32, device=dev) if need_bias else None
    s = view.stride()
    call("b2_f32_pack_weights", _p(view), cout, cin, kh, s[0], s[1], s[2], s[3], _p(scale), _p(shift), _p(bias), _p(wp),
         _p(bp), _stream())
    if len(_PACKED) > 4096:
        for k in [k for k, v in _PACKED.items() if v[0]() is None]:
            del _PACKED[k]
    _PACKED[key] = (weakref.ref(weight), stamp, wp, bp)
    return wp, bp


def conv(x, wp, bias, ksize, x1=None, stride=1, pad=None, relu=False, addend=None, add_after_act=False, out=None,
         out_mul=1, out_off=(0, 0), tap=None):
    """fp32 convolution; wp [taps][c0+c1][cout]; `tap`: use only that tap of wp as a 1x1 kernel (ConvTranspose phases)"""
    n, hi, wi, c0, ld0 = _nhwc(x)
    c1 = ld1 = 0
    if x1 is not None:
        n1, h1, w1, c1, ld1 = _nhwc(x1)
        assert (n1, h1, w1) == (n, hi, wi)
    taps, ctot, cout = wp.shape
    assert ctot == c0 + c1
    if tap is not None:
        wp = wp[tap:tap + 1]
        ksize = 1
    assert wp.shape[0] == ksize * ksize and wp.is_contiguous()
    pad = (ksize // 2, ksize // 2) if pad is None else pad
    ho = (hi + 2 * pad[0] - ksize) // stride + 1
    wo = (wi + 2 * pad[1] - ksize) // stride + 1
    y = out if out is not None else torch.empty((n, ho * out_mul, wo * out_mul, cout), dtype=torch.float32, device=x.device)
    ny, hy, wy, cy, ldy = _nhwc(y)
    assert (ny, hy, wy, cy) == (n, ho * out_mul, wo * out_mul, cout)
    a = F32ConvArgs()
    a.x0, a.x1 = x.data_ptr(), (x1.data_ptr() if x1 is not None else None)
    a.c0, a.c1, a.ldx0, a.ldx1 = c0, c1, ld0, ld1
    a.n, a.hi, a.wi, a.ho, a.wo = n, hi, wi, ho, wo
    a.ksize, a.stride, a.pad_h, a.pad_w = ksize, stride, pad[0], pad[1]
    a.w = wp.data_ptr()
    a.bias = bias.data_ptr() if bias is not None else None
    if addend is not None:
        na, ha, wa, ca, lda = _nhwc(addend)
        assert (na, ha, wa, ca) == (n, ho, wo, cout) and out_mul == 1
        a.addend, a.ldadd = addend.data_ptr(), lda
    a.add_after_act, a.relu, a.cout = int(add_after_act), int(relu), cout
    a.y, a.ldy, a.out_mul, a.out_off_h, a.out_off_w = y.data_ptr(), ldy, out_mul, out_off[0], out_off[1]
    call("b2_f32_conv", C.byref(a), _stream())
    return y


def maxpool(x, ksize=2, stride=2, pad=0):
    n, h, w, c, ld = _nhwc(x)
    assert ld == c
    ho, wo = (h + 2 * pad - ksize) // stride + 1, (w + 2 * pad - ksize) // stride + 1
    y = torch.empty((n, ho, wo, c), dtype=torch.float32, device=x.device)
    call("b2_f32_maxpool", _p(x), n, h, w, c, ksize, stride, pad, _p(y), _stream())
    return y


def upsample2x(x):
    n, h, w, c, ld = _nhwc(x)
    assert ld == c
    y = torch.empty((n, 2 * h, 2 * w, c), dtype=torch.float32, device=x.device)
    call("b2_f32_upsample2x", _p(x), n, h, w, c, _p(y), _stream())
    return y


def to_nhwc(x):
    """NCHW fp32 -> NHWC fp32"""
    x = x.float().contiguous()
    n, c, h, w = x.shape
    y = torch.empty((n, h, w, c), dtype=torch.float32, device=x.device)
    call("b2_f32_layout", _p(x), n, c, h * w, 0, _p(y), _stream())
    return y


def to_nchw(x):
    n, h, w, c, ld = _nhwc(x)
    assert ld == c
    y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    call("b2_f32_layout", _p(x), n, c, h * w, 1, _p(y), _stream())
    return y


# ----------------------------------------------------------------------------------------------------------
# blocks (module objects with the reference's layout)
# ----------------------------------------------------------------------------------------------------------
def conv_bn_relu(x, cv, bn, relu=True, x1=None, addend=None, add_after_act=False):
    wp, bp = _packed(cv.weight, cv.bias, bn)
    return conv(x, wp, bp, cv.kernel_size[0], x1=x1, stride=cv.stride[0], pad=(cv.padding[0], cv.padding[1]), relu=relu,
                addend=addend, add_after_act=add_after_act)


def basic_block(blk, x, x1=None):
    """[conv3x3, BN, ReLU] x 2 — AttentionUNet.py:4-13, ResnetUnet.py:5-14"""
    return conv_bn_relu(conv_bn_relu(x, blk[0], blk[1], x1=x1), blk[3], blk[4])


def up_conv(up, x):
    """Upsample(x2 nearest) -> conv3x3 -> BN -> ReLU — AttentionUNet.py:15-27"""
    return conv_bn_relu(upsample2x(x), up.up[1], up.up[2])


def attention_gate(att, g, x):
    """AttentionUNet.py:29-54"""
    g1 = conv_bn_relu(g, att.W_g[0], att.W_g[1], relu=False)
    x1 = conv_bn_relu(x, att.W_x[0], att.W_x[1], relu=False)
    bn1 = att.psi[1]
    coef = K.bn_eval_coeffs(bn1.weight, bn1.bias, bn1.running_mean, bn1.running_var, float(bn1.eps))
    n, h, w, c, ld = _nhwc(x)
    fint = g1.shape[3]
    out = torch.empty((n, h, w, c), dtype=torch.float32, device=x.device)
    wpsi = att.psi[0].weight.detach().reshape(-1)
    call("b2_f32_gate_tail", _p(g1), _p(x1), fint, _p(wpsi), _p(att.psi[0].bias), _p(coef[2]), _p(coef[3]), _p(x), ld,
         c, n * h * w, _p(out), c, _stream())
    return out


def recurrent_block(rb, x):
    """R2U_Net.py:4-20: t+1 applications of one conv+BN+ReLU; `x + x1` is produced by the conv epilogue"""
    cv, bn = rb.conv[0], rb.conv[1]
    if rb.t < 1:
        raise ValueError("Recurrent_block needs t >= 1")
    s = conv_bn_relu(x, cv, bn, addend=x, add_after_act=True)          # x + f(x)
    for _ in range(rb.t - 1):
        s = conv_bn_relu(s, cv, bn, addend=x, add_after_act=True)      # x + f(x + x1)
    return conv_bn_relu(s, cv, bn)                                     # f(x + x1)


def rrcnn_block(blk, x, x1=None):
    """R2U_Net.py:36-48: x0 = conv1x1(x); x0 + RB(RB(x0))"""
    wp, bp = _packed(blk.conv_1x1.weight, blk.conv_1x1.bias, None)
    x0 = conv(x, wp, bp, 1, x1=x1)
    r = recurrent_block(blk.RCNN[0], x0)
    cv, bn = blk.RCNN[1].conv[0], blk.RCNN[1].conv[1]
    rb = blk.RCNN[1]
    s = conv_bn_relu(r, cv, bn, addend=r, add_after_act=True)
    for _ in range(rb.t - 1):
        s = conv_bn_relu(s, cv, bn, addend=r, add_after_act=True)
    return conv_bn_relu(s, cv, bn, addend=x0, add_after_act=True)      # x0 + RB(RB(x0)): the residual rides the last conv


def conv_transpose2x2(ct, x):
    """ConvTranspose2d(k2, s2) as four 1x1 convolutions with pixel-shuffle placement — ResnetUnet.py:21,53"""
    wp, bp = _packed(ct.weight, ct.bias, None, transposed=True)
    n, h, w, _, _ = _nhwc(x)
    cout = wp.shape[2]
    y = torch.empty((n, 2 * h, 2 * w, cout), dtype=torch.float32, device=x.device)
    for i in range(2):
        for j in range(2):
            conv(x, wp, bp, 1, pad=(0, 0), out=y, out_mul=2, out_off=(i, j), tap=i * 2 + j)
    return y


def head(cv, x):
    wp, bp = _packed(cv.weight, cv.bias, None)
    return to_nchw(conv(x, wp, bp, 1, pad=(0, 0)))


# ----------------------------------------------------------------------------------------------------------
# model forwards
# ----------------------------------------------------------------------------------------------------------
def attention_unet(m, x):
    """AttentionUNet.forward — AttentionUNet.py:86-121"""
    a = to_nhwc(x)
    x1 = basic_block(m.conv1, a)
    x2 = basic_block(m.conv2, maxpool(x1))
    x3 = basic_block(m.conv3, maxpool(x2))
    x4 = basic_block(m.conv4, maxpool(x3))
    x5 = basic_block(m.conv5, maxpool(x4))
    d = x5
    for up, att, blk, skip in ((m.up5, m.att5, m.up_conv5, x4), (m.up4, m.att4, m.up_conv4, x3),
                               (m.up3, m.att3, m.up_conv3, x2), (m.up2, m.att2, m.up_conv2, x1)):
        d = up_conv(up, d)
        s = attention_gate(att, d, skip)
        d = basic_block(blk, s, x1=d)                      # cat((skip, d), dim=1)
    return head(m.out, d)


def r2_net(m, x, gates):
    """R2U_Net.forward (R2U_Net.py:78-111) / R2AttU_Net.forward (R2AttU_Net.py:118-158)"""
    a = to_nhwc(x)
    x1 = rrcnn_block(m.RRCNN1, a)
    x2 = rrcnn_block(m.RRCNN2, maxpool(x1))
    x3 = rrcnn_block(m.RRCNN3, maxpool(x2))
    x4 = rrcnn_block(m.RRCNN4, maxpool(x3))
    x5 = rrcnn_block(m.RRCNN5, maxpool(x4))
    d = x5
    for lvl, skip in ((5, x4), (4, x3), (3, x2), (2, x1)):
        d = up_conv(getattr(m, f"up{lvl}"), d)
        s = attention_gate(getattr(m, f"att{lvl}"), d, skip) if gates else skip
        d = rrcnn_block(getattr(m, f"up_RRCNN{lvl}"), s, x1=d)
    return head(m.conv_1x1, d)


def _bottleneck(blk, x):
    """torchvision Bottleneck (v1.5: the stride sits on conv2)"""
    o = conv_bn_relu(x, blk.conv1, blk.bn1)
    o = conv_bn_relu(o, blk.conv2, blk.bn2)
    idt = x if blk.downsample is None else conv_bn_relu(x, blk.downsample[0], blk.downsample[1], relu=False)
    return conv_bn_relu(o, blk.conv3, blk.bn3, addend=idt)            # relu(bn3(conv3(o)) + identity)


def resnet_unet(m, x):
    """ResNetUnet.forward — ResnetUnet.py:68-83"""
    a = to_nhwc(x)
    e1 = conv_bn_relu(a, m.encoder1[0], m.encoder1[1])
    t = maxpool(e1, 3, 2, 1)
    feats = [e1]
    for layer in (m.encoder2, m.encoder3, m.encoder4, m.encoder5):
        for blk in layer:
            t = _bottleneck(blk, t)
        feats.append(t)
    e1, e2, e3, e4, e5 = feats
    d = e5
    for dec, skip in ((m.decoder5, e4), (m.decoder4, e3), (m.decoder3, e2), (m.decoder2, e1)):
        up = conv_transpose2x2(dec.up_sample, d)
        d = basic_block(dec.basic_block, up, x1=skip)      # cat([up, skip], dim=1)
    d = conv_transpose2x2(m.decoder1[0], d)
    d = conv_bn_relu(d, m.decoder1[1], m.decoder1[2])
    return head(m.out, d)
