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


def test_fp16_operand_library_exports_the_same_abi():
    """libunetb200_f16.so (same kernels, IEEE-half storage, inference) is loaded beside the bf16 library and exports
    every symbol include/unetb200.h declares."""
    a, b = _lib.load("bf16"), _lib.load("fp16")
    assert a is not b
    for n in _lib.exported_symbols():
        assert hasattr(b, n), n
    assert b.unetb200_num_params() == a.unetb200_num_params() == 24_436_369
    with pytest.raises(ValueError):
        _lib.load("fp8")
    with pytest.raises(ValueError):
        vb.Unet("resnet34", precision="int8")
    assert vb.Unet("resnet34", precision="fp16").precision == "fp16"


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
        vb.Unet("resnet34", encoder_weights="ssl")
    with pytest.raises(ValueError):
        vb.Unet("resnet34", classes=2)
    m = vb.Unet("resnet34")
    with pytest.raises(vb.UnetB200Error, match="no CPU fallback"):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(ValueError):
        vb.losses.DiceLoss(mode="multiclass")


def test_imagenet_encoder_weights_from_a_local_torchvision_checkpoint(tmp_path, monkeypatch):
    """train.py:753,595 passes encoder_weights="imagenet": accepted when a torchvision-format resnet34 file is on local
    disk (what smp would have downloaded), a FileNotFoundError with the path hint otherwise."""
    import torchvision

    monkeypatch.setenv("TORCH_HOME", str(tmp_path / "empty_hub"))
    monkeypatch.delenv("UNETB200_RESNET34_WEIGHTS", raising=False)
    with pytest.raises(FileNotFoundError, match="UNETB200_RESNET34_WEIGHTS"):
        vb.Unet("resnet34", encoder_weights="imagenet")
    torch.manual_seed(7)
    tv = torchvision.models.resnet34(weights=None)
    with torch.no_grad():
        for b in tv.buffers():
            if b.dtype == torch.float32:
                b.uniform_(0.5, 1.5)
    f = tmp_path / "resnet34-local.pth"
    torch.save(tv.state_dict(), f)
    monkeypatch.setenv("UNETB200_RESNET34_WEIGHTS", str(f))
    m = vb.Unet("resnet34", encoder_weights="imagenet", in_channels=3, classes=1, activation=None)
    sd = m.state_dict()
    for k, v in tv.state_dict().items():
        if not k.startswith("fc."):
            assert torch.equal(sd["encoder." + k], v), k
    bad = tmp_path / "bad.pth"
    torch.save({"conv1.weight": torch.zeros(1)}, bad)
    monkeypatch.setenv("UNETB200_RESNET34_WEIGHTS", str(bad))
    with pytest.raises(RuntimeError, match="not a torchvision resnet34"):
        vb.Unet("resnet34", encoder_weights="imagenet")


def test_weight_cache_key_sees_in_place_updates_after_to():
    """after `.to()` every Parameter has its own version counter, so the key that decides whether the
    bf16 operand caches are stale must include the per-parameter versions (stock torch.optim.AdamW updates in place)."""
    m = vb.Unet("resnet34").to(torch.float32)   # _apply -> _reflatten -> `p.data = view`

    class _Ctx:  # stands in for the native context: records whether a re-pack was requested
        calls = 0
        handle = None

        class lib:  # noqa: N801
            @staticmethod
            def unetb200_load_weights_ex(*a):
                _Ctx.calls += 1
                return 0

        @staticmethod
        def check(rc, what):
            assert rc == 0

    m._sync_weights(_Ctx, 0)
    assert _Ctx.calls == 1
    m._sync_weights(_Ctx, 0)
    assert _Ctx.calls == 1                      # nothing changed: no re-pack
    with torch.no_grad():
        m.encoder.conv1.weight.add_(1.0)        # what optimizer.step() does
    m._sync_weights(_Ctx, 0)
    assert _Ctx.calls == 2
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    for p in m.parameters():
        p.grad = torch.ones_like(p)
    opt.step()
    m._sync_weights(_Ctx, 0)
    assert _Ctx.calls == 3
    with torch.no_grad():
        m.encoder.bn1.running_var.mul_(2.0)     # buffers too (eval-mode fold)
    m.eval()
    m._sync_weights(_Ctx, 0)
    assert _Ctx.calls == 4


def test_deepcopy_and_pickle_keep_the_flat_layout():
    """EMA / best-model snapshots: copy.deepcopy(model) and torch.save(model) must work and must not share (or
    double-free) the native context; the copy's parameters are again views of ONE flat array."""
    import copy
    import io

    m = vb.Unet("resnet34")
    m._ctx = object()            # pretend a forward created the native context (a ctypes handle cannot be pickled)
    m._packed_version = (1, 2, 3)
    c = copy.deepcopy(m)
    assert c._ctx is None and c._packed_version is None and m._ctx is not None
    assert c.encoder.conv1.weight.data_ptr() == c.flat_params.data_ptr() != m.flat_params.data_ptr()
    assert all(torch.equal(a, b) for a, b in zip(c.state_dict().values(), m.state_dict().values()))
    with torch.no_grad():
        c.segmentation_head._modules["0"].bias.fill_(3.0)
    assert float(c.flat_params[-1]) == 3.0 and float(m.flat_params[-1]) == 0.0
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    r = torch.load(buf, weights_only=False)
    assert r._ctx is None and r.encoder.conv1.weight.data_ptr() == r.flat_params.data_ptr()
    assert all(torch.equal(a, b) for a, b in zip(r.state_dict().values(), m.state_dict().values()))


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
