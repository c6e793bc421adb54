"""GPU parity tests (B200): CUDA path through the C ABI vs the fp32 oracle.

Per-kernel tests are point-wise (one bf16 output ulp).  Whole-model tests on RANDOM-INIT weights compare against the
fp32 oracle AND against the oracle with the CUDA path's bf16 rounding points emulated in fp32
(oracle/bf16_emulation.py): such a network is ill-conditioned (rounding only weights and input to bf16 moves the fp32
oracle by 0.08 mean-abs), so BASELINE.json's north_star tolerance (2e-2 max-abs / 1e-3 mean-abs / IoU 0.999 vs fp32) is
evaluated where it was meant to hold — on a TRAINED network and the reference's own micrographs — in
tests/test_gpu_trained.py, which also persists every distance (profiles/parity_r2.json).
"""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import vickers_hardness_unet_b200 as vb
from vickers_hardness_unet_b200 import _lib
from oracle import build_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _iou(a, b):
    a, b = a.bool(), b.bool()
    inter = (a & b).flatten(1).sum(1).float()
    union = (a | b).flatten(1).sum(1).float()
    return float(((inter + 1e-7) / (union + 1e-7)).mean())  # /root/reference/train.py:262-281


@pytest.fixture(scope="module")
def models():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    o = build_oracle(42)
    m = vb.Unet("resnet34", encoder_weights=None, in_channels=3, classes=1, activation=None)
    m.load_state_dict(o.state_dict(), strict=True)
    return o, m.to("cuda")


def _calibrate(o, m, size=128):
    """Fill BN running statistics (momentum=None -> cumulative average) so eval activations are O(1)."""
    o.train()
    for mod in o.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.reset_running_stats()
            mod.momentum = None
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for _ in range(2):
            o(torch.randn(4, 3, size, size, generator=g))
    o.eval()
    m.load_state_dict(o.state_dict(), strict=True)
    m.eval()


@pytest.mark.parametrize("cfg", [
    # N, H, W, cin, cout, k, stride, residual, relu
    (2, 32, 32, 64, 64, 3, 1, True, True),
    (1, 16, 48, 128, 256, 3, 2, False, True),
    (3, 8, 8, 256, 512, 1, 2, False, False),
    (1, 64, 64, 16, 16, 3, 1, False, True),
    (2, 24, 40, 32, 64, 3, 1, True, False),
])
def test_conv_kernel_parity(cfg):
    N, H, W, cin, cout, k, stride, res, relu = cfg
    lib = _lib.load()
    ctx = _lib.Context(0, 1, 32, 32)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N, cin, H, W, device="cuda", generator=g)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5
    sc = torch.rand(cout, device="cuda", generator=g) + 0.5
    sh = torch.randn(cout, device="cuda", generator=g) * 0.1
    Ho, Wo = H // stride, W // stride
    r = torch.randn(N, cout, Ho, Wo, device="cuda", generator=g) if res else None
    xb = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    rb = r.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16) if res else None
    out = torch.empty(N, Ho, Wo, cout, device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(cout, 2, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ctx.check(lib.unetb200_conv_nhwc(ctx.handle, xb.data_ptr(), w.data_ptr(), sc.data_ptr(), sh.data_ptr(),
                                     rb.data_ptr() if res else None, int(relu), out.data_ptr(), stats.data_ptr(),
                                     N, H, W, cin, cout, k, stride, st), "conv_nhwc")
    torch.cuda.synchronize()
    assert ctx.device_error_flag() == 0
    # fp32 reference over the same bf16-rounded operands
    ref = F.conv2d(xb.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), None, stride, k // 2)
    ref = ref * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    if res:
        ref = ref + rb.float().permute(0, 3, 1, 2)
    if relu:
        ref = ref.relu()
    got = out.float().permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item()
    assert err <= 2 ** -7 * ref.abs().max().item() + 1e-3, err  # one bf16 output rounding
    s_ref = torch.stack([got.sum((0, 2, 3)), (got * got).sum((0, 2, 3))], 1)
    assert torch.allclose(stats, s_ref, rtol=1e-3, atol=1e-2)
    ctx.close()


@pytest.mark.parametrize("cfg", [(2, 32, 32, 64, 64, 3, 1, True, True), (1, 16, 48, 128, 256, 3, 2, False, True)])
def test_conv_kernel_parity_fp16_operand_build(cfg):
    """The same kernel-level check through libunetb200_f16.so: IEEE-half operands and outputs (one half ulp = 2^-10)."""
    N, H, W, cin, cout, k, stride, res, relu = cfg
    lib = _lib.load("fp16")
    ctx = _lib.Context(0, 1, 32, 32, "fp16")
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N, cin, H, W, device="cuda", generator=g)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5
    sc = torch.rand(cout, device="cuda", generator=g) + 0.5
    sh = torch.randn(cout, device="cuda", generator=g) * 0.1
    Ho, Wo = H // stride, W // stride
    r = torch.randn(N, cout, Ho, Wo, device="cuda", generator=g) if res else None
    xh = x.permute(0, 2, 3, 1).contiguous().to(torch.float16)
    rh = r.permute(0, 2, 3, 1).contiguous().to(torch.float16) if res else None
    out = torch.empty(N, Ho, Wo, cout, device="cuda", dtype=torch.float16)
    st = torch.cuda.current_stream().cuda_stream
    ctx.check(lib.unetb200_conv_nhwc(ctx.handle, xh.data_ptr(), w.data_ptr(), sc.data_ptr(), sh.data_ptr(),
                                     rh.data_ptr() if res else None, int(relu), out.data_ptr(), None,
                                     N, H, W, cin, cout, k, stride, st), "conv_nhwc")
    torch.cuda.synchronize()
    assert ctx.device_error_flag() == 0
    ref = F.conv2d(xh.float().permute(0, 3, 1, 2), w.to(torch.float16).float(), None, stride, k // 2)
    ref = ref * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    if res:
        ref = ref + rh.float().permute(0, 3, 1, 2)
    if relu:
        ref = ref.relu()
    err = (out.float().permute(0, 3, 1, 2) - ref).abs().max().item()
    assert err <= 2 ** -10 * ref.abs().max().item() + 2e-4, err
    ctx.close()


def _report(tag, got, emu, ref):
    d_ge, d_gr, d_er = (got - emu).abs(), (got - ref).abs(), (emu - ref).abs()
    print(f"\n[{tag}] |logit| max {ref.abs().max():.3f} mean {ref.abs().mean():.3f} | "
          f"cuda-vs-bf16emu max {d_ge.max():.5f} mean {d_ge.mean():.6f} IoU {_iou(got >= 0, emu >= 0):.5f} | "
          f"cuda-vs-fp32 max {d_gr.max():.5f} mean {d_gr.mean():.6f} IoU {_iou(got >= 0, ref >= 0):.5f} | "
          f"bf16emu-vs-fp32 max {d_er.max():.5f} mean {d_er.mean():.6f} IoU {_iou(emu >= 0, ref >= 0):.5f}")
    return d_ge, d_gr, d_er


def test_golden_fixture_eval(models):
    """Committed oracle vectors (tests/golden): raw random-init eval model (|logit| up to ~25)."""
    from oracle.bf16_emulation import Bf16EmulatedUnet
    o, m = models
    o = build_oracle(42).eval()
    m.load_state_dict(o.state_dict(), strict=True)
    m.eval()
    g = np.load(os.path.join(GOLD, "oracle_golden.npz"))
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        got = m(x.cuda()).cpu()
        emu = Bf16EmulatedUnet(o)(x)
    ref = torch.from_numpy(g["logits_eval"])
    d_ge, d_gr, d_er = _report("golden eval", got, emu, ref)
    # the CUDA result must sit as close to the fp32 oracle as bf16 operands allow, and closer still to the emulation
    assert d_gr.mean().item() <= 1.25 * d_er.mean().item() + 1e-4
    assert d_ge.mean().item() <= d_er.mean().item()


@pytest.mark.parametrize("shape", [(2, 512, 512), (1, 256, 384), (3, 64, 64), (1, 96, 160), (1, 1024, 1024)])
def test_model_parity(models, shape):
    """CUDA path vs (a) the bf16-rounding-point emulation of the oracle and (b) the fp32 oracle.

    north_star tolerance (2e-2 max-abs / 1e-3 mean-abs / IoU 0.999 vs fp32) is NOT attainable with bf16 operands on this
    random-init network: rounding only weights+input to bf16 in the fp32 oracle already exceeds it (see
    oracle/bf16_emulation.py); the numbers are printed for the record and the asserted bar is "as close to fp32 as the
    bf16 emulation is, and much closer to the emulation itself".
    """
    from oracle.bf16_emulation import Bf16EmulatedUnet
    o, m = models
    _calibrate(o, m)
    N, H, W = shape
    g = torch.Generator().manual_seed(100 + H)
    x = torch.randn(N, 3, H, W, generator=g)
    with torch.no_grad():
        ref = o(x)
        emu = Bf16EmulatedUnet(o)(x)
        got = m(x.cuda()).cpu()
        mask = m.predict_mask(x.cuda(), 0.5).cpu()
    d_ge, d_gr, d_er = _report(f"parity {shape}", got, emu, ref)
    assert got.shape == ref.shape == (N, 1, H, W)
    assert d_gr.mean().item() <= 1.25 * d_er.mean().item() + 1e-4
    assert d_gr.max().item() <= 2.0 * d_er.max().item() + 1e-3
    assert d_ge.mean().item() <= d_er.mean().item()
    assert _iou(got >= 0, ref >= 0) >= _iou(emu >= 0, ref >= 0) - 0.02
    assert torch.equal(mask > 0, torch.sigmoid(got) >= 0.5)
    assert m._ctx.device_error_flag() == 0


def test_structured_input(models):
    """Dark diamond on textured background (fg ~5 %, like data/masks), ImageNet-normalised as infer_pth_gui.py:47-48."""
    from oracle.bf16_emulation import Bf16EmulatedUnet
    o, m = models
    _calibrate(o, m)
    H = W = 256
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    g = torch.Generator().manual_seed(5)
    imgs = []
    for i in range(2):
        cx, cy, r = 100 + 40 * i, 128, 36 + 8 * i
        diamond = ((xx - cx).abs() + (yy - cy).abs() < r).float()
        base = 0.6 + 0.1 * torch.randn(H, W, generator=g) - 0.45 * diamond
        rgb = torch.stack([base, base * 0.98, base * 1.02]).clamp(0, 1)
        mean = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
        std = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)
        imgs.append((rgb - mean) / std)
    x = torch.stack(imgs)
    with torch.no_grad():
        ref = o(x)
        emu = Bf16EmulatedUnet(o)(x)
        got = m(x.cuda()).cpu()
    d_ge, d_gr, d_er = _report("structured", got, emu, ref)
    assert d_gr.mean().item() <= 1.25 * d_er.mean().item() + 1e-4
    assert d_ge.mean().item() <= d_er.mean().item()


def test_host_buffer_entry_point(models):
    o, m = models
    _calibrate(o, m)
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        ref = m(x.cuda()).cpu()
    ctx = m._ctx
    xh = x.contiguous().pin_memory()
    out = torch.empty(2, 1, 64, 64).pin_memory()
    mask = torch.empty(2, 1, 64, 64, dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()
    ctx.check(ctx.lib.unetb200_infer_host(ctx.handle, xh.data_ptr(), out.data_ptr(), None, mask.data_ptr(), 0.5, 2),
              "infer_host")
    assert torch.equal(out, ref)
    assert torch.equal(mask > 0, ref >= 0)


def test_errors_are_loud(models):
    _, m = models
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 100, 64, device="cuda"))
    with pytest.raises(vb.UnetB200Error):
        m(torch.zeros(1, 3, 64, 64))
    with torch.inference_mode():
        y = m.eval()(torch.zeros(1, 3, 64, 64, device="cuda"))
    assert y.shape == (1, 1, 64, 64)
    with torch.autocast("cuda", dtype=torch.float16), torch.no_grad():
        y2 = m(torch.zeros(1, 3, 64, 64, device="cuda"))
    assert y2.dtype == torch.float32 and torch.equal(y, y2)


# ----------------------------------------------------------------------------------------------- host-buffer entry points
def test_host_submit_wait_two_slots_match_device_call(models):
    """unetb200_infer_host_submit/_wait: two requests in flight, each slot returns the result of ITS request,
    bit-identical to the device-tensor call on the same input."""
    o, m = models
    _calibrate(o, m, 64)
    g = torch.Generator().manual_seed(5)
    xa = torch.randn(3, 3, 64, 96, generator=g).pin_memory()
    xb = torch.randn(3, 3, 64, 96, generator=g).pin_memory()
    want = [m.predict_mask(x.cuda(), return_prob=True) for x in (xa, xb)]
    mk = [torch.empty(3, 1, 64, 96, dtype=torch.uint8).pin_memory() for _ in range(2)]
    pr = [torch.empty(3, 1, 64, 96).pin_memory() for _ in range(2)]
    lg = [torch.empty(3, 1, 64, 96).pin_memory() for _ in range(2)]
    for rep in range(3):  # slots are reusable
        m.submit_host(0, xa, mask_out=mk[0], prob_out=pr[0], logits_out=lg[0])
        m.submit_host(1, xb, mask_out=mk[1], prob_out=pr[1], logits_out=lg[1])
        with pytest.raises(_lib.UnetB200Error):
            m.submit_host(1, xb, mask_out=mk[1])  # still in flight
        m.wait_host(0)
        m.wait_host(1)
        for s in range(2):
            assert torch.equal(mk[s], want[s][0].cpu())
            assert torch.equal(pr[s], want[s][1].cpu())
            assert torch.equal(torch.sigmoid(lg[s]) >= 0.5, mk[s] > 0)
    assert torch.equal(m.predict_mask_host(xa), want[0][0].cpu())
    assert m._ctx.device_error_flag() == 0


def test_uint8_frames_equal_host_preprocessing(models):
    """uint8 BGR HWC frames through the fused pre-processing pack == the reference's host pre-processing
    (infer_pth_gui.py:46-48: BGR->RGB, /255, (x-mean)/std, HWC->CHW) followed by the fp32-tensor call."""
    o, m = models
    _calibrate(o, m, 64)
    g = torch.Generator().manual_seed(6)
    frames = torch.randint(0, 256, (2, 64, 96, 3), dtype=torch.uint8, generator=g)
    mean = torch.tensor(vb.Unet.IMAGENET_MEAN)
    std = torch.tensor(vb.Unet.IMAGENET_STD)
    rgb = frames.flip(-1).float() / 255.0
    x = ((rgb - mean) / std).permute(0, 3, 1, 2).contiguous()
    ref = m(x.cuda()).cpu()
    lg = torch.empty(2, 1, 64, 96).pin_memory()
    mk = torch.empty(2, 1, 64, 96, dtype=torch.uint8).pin_memory()
    m.submit_host(0, frames.pin_memory(), mask_out=mk, logits_out=lg, bgr=True)
    m.wait_host(0)
    # the only difference is fp32 rounding of (v/255 - mean) * (1/std) vs (v/255 - mean) / std before the bf16 pack
    d = (lg - ref).abs()
    print(f"uint8 path vs host pre-processing: max {float(d.max()):.3e} mean {float(d.mean()):.3e}")
    assert float(d.max()) <= 2e-2 and float(d.mean()) <= 1e-3
    assert _iou(mk > 0, ref >= 0) >= 0.999
    # RGB-ordered frames with bgr=False give the same result
    m.submit_host(1, frames.flip(-1).contiguous().pin_memory(), logits_out=lg, bgr=False)
    m.wait_host(1)
    assert float((lg - ref).abs().max()) <= 2e-2
