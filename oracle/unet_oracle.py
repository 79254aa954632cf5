"""ORACLE — test infrastructure only.  Never imported by the product path (b200seg/*); only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.

A plain-PyTorch *functional* restatement of the reference's segmentation forward passes, written against a
state_dict (no nn.Module classes), so it can run on the GPU box where /root/reference does not exist.  It is
pinned against the real reference modules by oracle/make_golden.py (run in the build container, where
/root/reference is importable) through the fixtures in tests/golden/; tests/test_oracle_golden.py re-checks
those fixtures on every CPU run.  The reference itself holds no tests or golden vectors.

Every function cites the reference lines it restates (paths relative to the reference root).
dtype / device follow the tensors in `sd` and `x` (fp64 = ground truth, fp32 = "reference result";
wrap the call in torch.autocast(bf16) for the reference's own reduced-precision noise floor).
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

BN_EPS = 1e-5       # nn.BatchNorm2d default used everywhere in the reference
BN_MOMENTUM = 0.1


class _Ctx:
    """Carries the state dict, train/eval flag and collects the running-stat side effects."""

    def __init__(self, sd, training):
        self.sd = sd
        self.training = training
        self.new_buffers = OrderedDict()

    def bn(self, x, prefix):
        """nn.BatchNorm2d forward (AttentionUNet.py:7,10,21,34,38,42; R2U_Net.py:11,28; ResnetUnet.py:8,11,55).
        Train mode: batch statistics, running stats updated with momentum 0.1 / unbiased variance, once per CALL
        (a Recurrent_block calls its single BN t+1 times, R2U_Net.py:15-20)."""
        sd = self.sd
        rm = self.new_buffers.get(prefix + ".running_mean", sd[prefix + ".running_mean"]).clone()
        rv = self.new_buffers.get(prefix + ".running_var", sd[prefix + ".running_var"]).clone()
        y = F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], self.training, BN_MOMENTUM, BN_EPS)
        if self.training:
            self.new_buffers[prefix + ".running_mean"] = rm
            self.new_buffers[prefix + ".running_var"] = rv
            k = prefix + ".num_batches_tracked"
            self.new_buffers[k] = self.new_buffers.get(k, sd[k]) + 1
        return y

    def conv(self, x, prefix, padding=0, stride=1):
        return F.conv2d(x, self.sd[prefix + ".weight"], self.sd.get(prefix + ".bias"), stride=stride, padding=padding)


def _basic_block(c, x, p):
    """basic_block: [conv3x3 p1 -> BN -> ReLU] x2   (AttentionUNet.py:4-13, ResnetUnet.py:5-14)"""
    x = torch.relu(c.bn(c.conv(x, p + ".0", 1), p + ".1"))
    return torch.relu(c.bn(c.conv(x, p + ".3", 1), p + ".4"))


def _up_conv(c, x, p):
    """UpConv: nearest x2 -> conv3x3 -> BN -> ReLU   (AttentionUNet.py:15-27, R2U_Net.py:22-34)"""
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    return torch.relu(c.bn(c.conv(x, p + ".up.1", 1), p + ".up.2"))


def _attention_gate(c, g, x, p):
    """AttentionGate.forward (AttentionUNet.py:48-54, R2AttU_Net.py:80-86):
    x * sigmoid(BN1(psi(relu(BN(W_g g) + BN(W_x x)))))"""
    g1 = c.bn(c.conv(g, p + ".W_g.0"), p + ".W_g.1")
    x1 = c.bn(c.conv(x, p + ".W_x.0"), p + ".W_x.1")
    a = torch.relu(g1 + x1)
    psi = torch.sigmoid(c.bn(c.conv(a, p + ".psi.0"), p + ".psi.1"))
    return x * psi


def attention_unet_forward(sd, x, training=True):
    """AttentionUNet.forward (AttentionUNet.py:86-121).  Returns (logits, new_buffers)."""
    c = _Ctx(sd, training)
    pool = lambda t: F.max_pool2d(t, 2, 2)                      # AttentionUNet.py:61
    x1 = _basic_block(c, x, "conv1")
    x2 = _basic_block(c, pool(x1), "conv2")
    x3 = _basic_block(c, pool(x2), "conv3")
    x4 = _basic_block(c, pool(x3), "conv4")
    x5 = _basic_block(c, pool(x4), "conv5")
    d = x5
    for lvl, skip in ((5, x4), (4, x3), (3, x2), (2, x1)):
        d = _up_conv(c, d, f"up{lvl}")
        s = _attention_gate(c, d, skip, f"att{lvl}")
        d = _basic_block(c, torch.cat((s, d), dim=1), f"up_conv{lvl}")   # skip FIRST (AttentionUNet.py:101)
    return c.conv(d, "out"), c.new_buffers


def _recurrent_block(c, x, p, t):
    """Recurrent_block.forward (R2U_Net.py:15-20): one shared conv+BN+ReLU applied t+1 times."""
    f = lambda v: torch.relu(c.bn(c.conv(v, p + ".conv.0", 1), p + ".conv.1"))
    x1 = None
    for i in range(t):
        if i == 0:
            x1 = f(x)
        x1 = f(x + x1)
    return x1


def _rrcnn_block(c, x, p, t):
    """RRCNN_block.forward (R2U_Net.py:45-48): x0 = conv1x1(x); x0 + RB(RB(x0))"""
    x0 = c.conv(x, p + ".conv_1x1")
    x1 = _recurrent_block(c, _recurrent_block(c, x0, p + ".RCNN.0", t), p + ".RCNN.1", t)
    return x0 + x1


def _r2_forward(sd, x, t, training, gates):
    c = _Ctx(sd, training)
    pool = lambda v: F.max_pool2d(v, 2, 2)
    x1 = _rrcnn_block(c, x, "RRCNN1", t)
    x2 = _rrcnn_block(c, pool(x1), "RRCNN2", t)
    x3 = _rrcnn_block(c, pool(x2), "RRCNN3", t)
    x4 = _rrcnn_block(c, pool(x3), "RRCNN4", t)
    x5 = _rrcnn_block(c, pool(x4), "RRCNN5", t)
    d = x5
    for lvl, skip in ((5, x4), (4, x3), (3, x2), (2, x1)):
        d = _up_conv(c, d, f"up{lvl}")
        s = _attention_gate(c, d, skip, f"att{lvl}") if gates else skip
        d = _rrcnn_block(c, torch.cat((s, d), dim=1), f"up_RRCNN{lvl}", t)
    return c.conv(d, "conv_1x1"), c.new_buffers


def r2u_net_forward(sd, x, t=5, training=True):
    """R2U_Net.forward (R2U_Net.py:78-111); ctor default t=5 (R2U_Net.py:51)."""
    return _r2_forward(sd, x, t, training, gates=False)


def r2attu_net_forward(sd, x, t=5, training=True):
    """R2AttU_Net.forward (R2AttU_Net.py:118-158); ctor default t=5 (R2AttU_Net.py:89)."""
    return _r2_forward(sd, x, t, training, gates=True)


def _bottleneck(c, x, p, stride, has_down):
    """torchvision.models.resnet.Bottleneck (v1.5: stride on the 3x3) — third-party dependency of
    ResnetUnet.py:32 (torchvision pinned 0.24.0 in requirements.txt; same arithmetic in 0.26.0)."""
    idt = x
    o = torch.relu(c.bn(c.conv(x, p + ".conv1"), p + ".bn1"))
    o = torch.relu(c.bn(c.conv(o, p + ".conv2", 1, stride), p + ".bn2"))
    o = c.bn(c.conv(o, p + ".conv3"), p + ".bn3")
    if has_down:
        idt = c.bn(c.conv(x, p + ".downsample.0", 0, stride), p + ".downsample.1")
    return torch.relu(o + idt)


def _res_layer(c, x, p, blocks, stride):
    x = _bottleneck(c, x, f"{p}.0", stride, True)
    for i in range(1, blocks):
        x = _bottleneck(c, x, f"{p}.{i}", 1, False)
    return x


def _decoder_block(c, down, skip, p):
    """DecoderBlock.forward (ResnetUnet.py:23-27): ConvT2x2s2 -> cat([up, skip]) -> basic_block"""
    up = F.conv_transpose2d(down, c.sd[p + ".up_sample.weight"], c.sd[p + ".up_sample.bias"], stride=2)
    return _basic_block(c, torch.cat([up, skip], dim=1), p + ".basic_block")


def resnet_unet_forward(sd, x, training=True):
    """ResNetUnet.forward (ResnetUnet.py:68-83); encoder = torchvision resnet50 layers (ResnetUnet.py:32-43).
    model.train() puts the (frozen) encoder BNs in train mode too."""
    c = _Ctx(sd, training)
    e1 = torch.relu(c.bn(c.conv(x, "encoder1.0", 3, 2), "encoder1.1"))
    p1 = F.max_pool2d(e1, 3, 2, 1)
    e2 = _res_layer(c, p1, "encoder2", 3, 1)
    e3 = _res_layer(c, e2, "encoder3", 4, 2)
    e4 = _res_layer(c, e3, "encoder4", 6, 2)
    e5 = _res_layer(c, e4, "encoder5", 3, 2)
    d5 = _decoder_block(c, e5, e4, "decoder5")
    d4 = _decoder_block(c, d5, e3, "decoder4")
    d3 = _decoder_block(c, d4, e2, "decoder3")
    d2 = _decoder_block(c, d3, e1, "decoder2")
    d1 = F.conv_transpose2d(d2, sd["decoder1.0.weight"], sd["decoder1.0.bias"], stride=2)
    d1 = torch.relu(c.bn(c.conv(d1, "decoder1.1", 1), "decoder1.2"))
    return c.conv(d1, "out"), c.new_buffers


FORWARDS = {
    "AttentionUNet": attention_unet_forward,
    "R2U_Net": r2u_net_forward,
    "R2AttU_Net": r2attu_net_forward,
    "ResNetUnet": resnet_unet_forward,
}

# block-level entry points for the level-(ii) parity tests
def basic_block(sd, x, prefix, training=True):
    c = _Ctx(sd, training)
    return _basic_block(c, x, prefix), c.new_buffers


def up_conv(sd, x, prefix, training=True):
    c = _Ctx(sd, training)
    return _up_conv(c, x, prefix), c.new_buffers


def attention_gate(sd, g, x, prefix, training=True):
    c = _Ctx(sd, training)
    return _attention_gate(c, g, x, prefix), c.new_buffers


def rrcnn_block(sd, x, prefix, t=2, training=True):
    c = _Ctx(sd, training)
    return _rrcnn_block(c, x, prefix, t), c.new_buffers


def recurrent_block(sd, x, prefix, t=2, training=True):
    """Recurrent_block on its own (R2U_Net.py:4-20)"""
    c = _Ctx(sd, training)
    return _recurrent_block(c, x, prefix, t), c.new_buffers


def decoder_block(sd, down, skip, prefix, training=True):
    """DecoderBlock on its own (ResnetUnet.py:17-27)"""
    c = _Ctx(sd, training)
    return _decoder_block(c, down, skip, prefix), c.new_buffers


# ----------------------------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------------------------
def bce_with_logits(z, t):
    """nn.BCEWithLogitsLoss() mean reduction (utils/helpers.py:244-246,327), softplus form."""
    return (torch.clamp(z, min=0) - z * t + torch.log1p(torch.exp(-z.abs()))).mean()


def dice_loss(z, t, smooth=1.0):
    """DiceLoss (utils/clip_seg_finetuner.py:40-58): global over the flattened batch."""
    p = torch.sigmoid(z).reshape(-1)
    tt = t.reshape(-1)
    inter = (p * tt).sum()
    return 1 - (2.0 * inter + smooth) / (p.sum() + tt.sum() + smooth)


def combined_loss(z, t, w_bce=0.5, w_dice=0.5):
    """CombinedLoss (utils/clip_seg_finetuner.py:61-74)."""
    return w_bce * bce_with_logits(z, t) + w_dice * dice_loss(z, t)


def iou(pred, target, thresh=0.5):
    """iou() of utils/helpers.py:223-227: ((p>0.5)&t).sum / ((p>0.5)|t).sum + 1e-7 in the denominator."""
    p = pred > thresh
    tb = target > 0.5
    inter = (p & tb).sum().double()
    union = (p | tb).sum().double()
    return float(inter / (union + 1e-7))


def segmentation_metrics(pred, target, threshold=0.5):
    """calculate_segmentation_metrics of utils/tester.py:158-193 (with calculate_iou :92-111, calculate_dice :114-134,
    calculate_pixel_accuracy :137-155) for ONE sample: pred = probabilities after sigmoid.  Returns the same dict
    (values in percent)."""
    pb = (pred > threshold).double()
    tb = (target > threshold).double()
    inter = (pb * tb).sum()
    union = ((pb + tb) > 0).double().sum()
    iou_ = (inter + 1e-7) / (union + 1e-7)
    dice = (2.0 * inter + 1e-7) / (pb.sum() + tb.sum() + 1e-7)
    pix = (pb == tb).double().sum() / tb.numel()
    tp = float(inter)
    fp = float((pb * (1 - tb)).sum())
    fn = float(((1 - pb) * tb).sum())
    precision = (tp + 1e-7) / (tp + fp + 1e-7)
    recall = (tp + 1e-7) / (tp + fn + 1e-7)
    f1 = 2 * (precision * recall) / (precision + recall + 1e-7)
    return {"iou": float(iou_) * 100, "dice": float(dice) * 100, "pixel_accuracy": float(pix) * 100,
            "precision": precision * 100, "recall": recall * 100, "f1": f1 * 100}


def train_step_grads(name, sd, x, target, training=True, loss="bce", **fw_kwargs):
    """forward + loss + backward w.r.t. every floating-point parameter in `sd` (helpers.py:321-329 without AMP).
    Returns (logits, loss, grads: dict name -> tensor, new_buffers)."""
    params = OrderedDict()
    work = OrderedDict()
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            p = v.detach().clone().requires_grad_(True)
            params[k] = p
            work[k] = p
        else:
            work[k] = v
    logits, newb = FORWARDS[name](work, x, training=training, **fw_kwargs)
    lf = bce_with_logits if loss == "bce" else combined_loss
    lval = lf(logits.to(target.dtype) if logits.dtype != target.dtype else logits, target)
    grads = torch.autograd.grad(lval, list(params.values()), allow_unused=True)
    return logits.detach(), lval.detach(), OrderedDict(zip(params.keys(), grads)), newb


def state_dict_from_layout(keys, shapes, dtypes, seed=0, float_dtype=torch.float32):
    """A state_dict with the reference's layout (key order / shapes as stored in tests/golden/<model>.npz by
    make_golden.py from the real reference modules; shapes are 'd0,d1,..' strings, '' for scalars) and the
    deterministic synthetic fill — lets the CPU arm of bench.py build its weights without touching the CUDA package's
    modules.  Floating-point entries get `float_dtype` (the reference's parameters are fp32), integer ones int64."""
    from .synthetic import fill_state_dict_
    sd = OrderedDict()
    for k, shp, dt in zip(keys, shapes, dtypes):
        dims = tuple(int(d) for d in str(shp).split(",") if d != "")
        sd[str(k)] = torch.zeros(dims, dtype=float_dtype if "float" in str(dt) else torch.int64)
    return fill_state_dict_(sd, seed)
