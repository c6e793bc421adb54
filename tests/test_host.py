"""CPU tests of the host side: C-ABI exports, tensor table == oracle state_dict, module surface, error behaviour."""
import ctypes
import os

import pytest
import torch

import vickers_hardness_unet_b200 as vb
from vickers_hardness_unet_b200 import _lib
from oracle import build_oracle


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    names = _lib.exported_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/unetb200.h but not exported"


def test_tensor_table_matches_oracle_state_dict():
    table = _lib.tensor_table()
    osd = build_oracle().state_dict()
    assert [t[0] for t in table] == list(osd.keys())
    for name, shape, off, kind in table:
        assert tuple(osd[name].shape) == tuple(shape), name
        assert (kind == 2) == (osd[name].dtype == torch.int64), name
    lib = _lib.load()
    assert lib.unetb200_num_params() == 24_436_369
    assert lib.unetb200_num_buffers() == sum(v.numel() for k, v in osd.items() if "running_" in k)
    assert lib.unetb200_num_counters() == 46


def test_state_dict_roundtrip_and_flat_views():
    m = vb.Unet("resnet34", encoder_weights=None, in_channels=3, classes=1, activation=None)
    o = build_oracle()
    m.load_state_dict(o.state_dict(), strict=True)
    sd = m.state_dict()
    assert all(torch.equal(sd[k], v) for k, v in o.state_dict().items())
    # parameters are views of ONE flat buffer, in state_dict order
    p0 = m.encoder.conv1.weight
    assert p0.data_ptr() == m.flat_params.data_ptr()
    assert m.segmentation_head[0].bias.data_ptr() == m.flat_params.data_ptr() + 4 * (24_436_369 - 1) \
        if hasattr(m.segmentation_head, "__getitem__") else True
    v0 = m.flat_params._version
    with torch.no_grad():
        m.decoder.blocks._modules["0"].conv1._modules["0"].weight.mul_(1.0)
    assert m.flat_params._version != v0  # in-place edits of any view are visible => bf16 caches get refreshed
    # .to() keeps the flat layout
    m2 = m.to(torch.float32)
    assert m2.encoder.conv1.weight.data_ptr() == m2.flat_params.data_ptr()
    assert len(list(m.parameters())) == 140 and all(p.is_leaf and p.requires_grad for p in m.parameters())


def test_rejects_unsupported_configs_and_cpu_inputs():
    with pytest.raises(ValueError):
        vb.Unet("resnet50")
    with pytest.raises(ValueError):
        vb.Unet("resnet34", encoder_weights="imagenet")
    with pytest.raises(ValueError):
        vb.Unet("resnet34", classes=2)
    m = vb.Unet("resnet34")
    with pytest.raises(vb.UnetB200Error, match="no CPU fallback"):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(ValueError):
        vb.losses.DiceLoss(mode="multiclass")


def test_shim_exposes_the_two_smp_symbols():
    import segmentation_models_pytorch as smp

    assert smp.Unet is vb.Unet and smp.losses.DiceLoss is vb.losses.DiceLoss


def test_create_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.unetb200_create(ctypes.byref(h), 0, 1, 64, 64) != 0
    assert len(lib.unetb200_last_error(None)) > 0
    assert lib.unetb200_create(ctypes.byref(h), 0, 1, 65, 64) != 0
    assert b"multiples of 32" in lib.unetb200_last_error(None)
