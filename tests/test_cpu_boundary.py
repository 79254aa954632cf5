"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/b200seg.h declares,
the ctypes table covers the header, the modules have the reference's exact state_dict layout (from the golden
fixtures written by the real reference), and nothing silently falls back to a CPU path."""
import re
import warnings
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"


def _header_symbols():
    text = (ROOT / "include" / "b200seg.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from b200seg import _lib
    lib = _lib.load()
    names = _header_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200seg.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in b200seg/_lib.py"
    assert set(_lib.SIGNATURES) <= set(names), set(_lib.SIGNATURES) - set(names)
    assert lib.b2_abi_version() == 1


def test_no_gpu_means_error_not_fallback():
    from b200seg import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    rc = lib.b2_arch_check()
    assert rc < 0 and lib.b2_last_error()
    from b200seg.models.segmentation_models import AttentionUNet
    m = AttentionUNet()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 64, 64))


@pytest.mark.parametrize("case,cls,kw", [("AttentionUNet", "AttentionUNet", {}), ("R2U_Net", "R2U_Net", {"t": 2}),
                                         ("R2AttU_Net", "R2AttU_Net", {"t": 2}), ("R2U_Net_t5", "R2U_Net", {}),
                                         ("ResNetUnet", "ResNetUnet", {})])
def test_state_dict_layout_matches_reference(case, cls, kw):
    from b200seg.models import segmentation_models as M
    g = np.load(GOLD / f"{case}.npz")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = getattr(M, cls)(**kw)
    sd = m.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    for (k, v), shp, dt in zip(sd.items(), g["shapes"], g["dtypes"]):
        assert ",".join(map(str, v.shape)) == str(shp), k
        # the fixture was taken from the reference module after .double(): float64 there == the module's float32
        assert str(v.dtype).replace("torch.", "") == str(dt).replace("float64", "float32"), k
    assert [int(p.requires_grad) for p in m.parameters()] == [int(r) for r in g["requires_grad"]]


def test_factory_and_ops_registered():
    from b200seg.utils import helpers
    import b200seg.ops  # noqa: F401
    import b200seg.ops_resnet  # noqa: F401
    for name, cls in (("attentionunet", "AttentionUNet"), ("r2unet", "R2U_Net"), ("r2attunet", "R2AttU_Net")):
        assert type(helpers.get_seg_model(name)).__name__ == cls
    with pytest.raises(ValueError):
        helpers.get_seg_model("nope")
    for op in ("conv2d", "conv_bn_act", "bn_apply", "gate_mid", "head", "seg_loss", "maxpool2x2", "upsample2x",
               "conv_transpose2x2", "enc_conv_bn"):
        assert hasattr(torch.ops.b200seg, op), op


def test_fake_kernels_give_reference_shapes():
    """register_fake implementations (shape/dtype propagation without a GPU)"""
    import b200seg.ops as ops
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        x = torch.empty(2, 32, 32, 64, dtype=torch.bfloat16, device="cuda")
        w = torch.empty(128, 64, 3, 3, device="cuda")
        g = torch.empty(128, device="cuda")
        y, z, coef, stats, x4 = ops.conv_bn_act(x, None, w, None, g, g, g, g, True, 1e-5, True)
        assert y.shape == (2, 32, 32, 128) and y.dtype == torch.bfloat16 and coef.shape == (4, 128)
        assert ops.maxpool2x2(y).shape == (2, 16, 16, 128)
        assert ops.upsample2x(y).shape == (2, 64, 64, 128)
        hw = torch.empty(1, 128, 1, 1, device="cuda")
        assert ops.head(y, hw, None).shape == (2, 1, 32, 32)


def test_host_side_helpers_degrade_gracefully_without_a_gpu():
    """kernels.zeros_scratch / step_begin / wgrad_stream are no-ops (plain torch.zeros, same stream) on a machine
    without CUDA and outside a backward pass — they never touch the library."""
    import torch
    from b200seg import kernels as K
    K.step_begin()
    z = K.zeros_scratch((2, 8), torch.float64, torch.device("cpu"))
    assert z.shape == (2, 8) and z.dtype == torch.float64 and float(z.abs().sum()) == 0.0
    K.set_wgrad_overlap(True)
    try:
        with K.wgrad_stream(z):
            pass
        assert K.wgrad_side_stream() is None
    finally:
        K.set_wgrad_overlap(False)


def test_library_is_native_sm100a_code():
    """The shipped .so holds hand-written sm_100a SASS — tcgen05 MMAs (UTCHMMA, also the 2-CTA form), TMA loads / stores
    (UTMALDG / UTMASTG), TMEM loads (LDTM), the programmatic-dependent-launch pair (ACQBULK = griddepcontrol.wait,
    PREEXIT = griddepcontrol.launch_dependents) — and links no vendor kernel library."""
    import shutil
    import subprocess
    from b200seg import _lib
    so = Path(_lib.load()._name)
    ldd = subprocess.run(["ldd", str(so)], capture_output=True, text=True).stdout.lower()
    for vendor in ("cudnn", "cublas", "nccl", "cutlass"):
        assert vendor not in ldd, f"libb200seg.so links {vendor}"
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    r = subprocess.run([cuobjdump, "-sass", str(so)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    sass = r.stdout
    assert "sm_100a" in sass
    count = {k: len(re.findall(r"\b" + k, sass)) for k in ("UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "LDTM",
                                                            "ACQBULK", "PREEXIT")}
    assert count["UTCHMMA"] >= 30 and count["UTCHMMA.2CTA"] >= 1, count
    assert count["UTMALDG"] >= 40 and count["UTMASTG"] >= 10 and count["LDTM"] >= 10, count
    assert count["ACQBULK"] >= 20 and count["PREEXIT"] >= 20, count
