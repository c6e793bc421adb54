"""Layer-local parity of the backward kernels (B200).

Whole-network gradient comparisons against the fp32 oracle are dominated by how a random-init ReLU/BatchNorm network
amplifies bf16 rounding (the bf16-EMULATED oracle itself is 10-80 % away from the fp32 oracle layer by layer, see
scripts/train_debug.py).  So every backward kernel is checked here on ITS OWN inputs: the library's saved tensors
(z, a, dz, dA, batch mean / invstd) are read back through unetb200_train_debug_* and each dz, dA, d_skip and parameter
gradient is re-derived from them with PyTorch fp32 ops (torch.nn.grad.conv2d_input / conv2d_weight, the BatchNorm
backward closed form, max_pool2d autograd).  Tolerance: relative L2 1e-2 for bf16-stored tensors (one rounding),
2e-3 for fp32 parameter gradients.
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F
from torch.nn.grad import conv2d_input, conv2d_weight

import vickers_hardness_unet_b200 as vb
from oracle import build_oracle

pytestmark = pytest.mark.gpu


def _debug_tensors(model, N):
    ctx = model._ctx
    lib = ctx.lib
    out = {}
    name = C.create_string_buffer(256)
    shape = (C.c_int * 4)()
    bf = C.c_int()
    st = torch.cuda.current_stream().cuda_stream
    n = lib.unetb200_train_debug_count(ctx.handle, N)
    assert n > 0
    for i in range(n):
        assert lib.unetb200_train_debug_info(ctx.handle, N, i, name, 256, shape, C.byref(bf)) == 0
        n_, h, w, c = (shape[j] for j in range(4))
        t = torch.empty((n_, h, w, c), dtype=torch.bfloat16 if bf.value else torch.float32, device="cuda")
        ctx.check(lib.unetb200_train_debug_copy(ctx.handle, N, i, t.data_ptr(), t.numel() * t.element_size(), st),
                  "debug_copy")
        key = name.value.decode()
        out[key] = t.float().permute(0, 3, 1, 2).contiguous() if (n_, h, w) != (1, 1, 1) else t.float().view(-1)
    torch.cuda.synchronize()
    return out


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


class _Checker:
    def __init__(self):
        self.rows = []

    def add(self, what, got, ref, tol):
        r = _rel(got, ref)
        self.rows.append((what, r, tol, float(ref.norm())))

    def report(self):
        bad = [r for r in self.rows if not r[1] <= r[2]]
        worst = sorted(self.rows, key=lambda r: -r[1] / r[2])[:10]
        print(f"\n[layer-local backward parity] {len(self.rows)} checks, {len(bad)} over tolerance; worst:")
        for what, r, tol, nrm in worst:
            print(f"    {what:60s} rel-L2 {r:.5f} (tol {tol}) |ref| {nrm:.3e}")
        return bad


# the last two are the BASELINE train configurations themselves (configs[2]: 16 x 512^2 per GPU, configs[3]: 8 x 1024^2)
@pytest.mark.parametrize("shape", [(2, 64, 96), (3, 128, 128), (1, 256, 256), (16, 512, 512), (8, 1024, 1024)])
def test_every_backward_kernel_on_its_own_inputs(shape):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    N, H, W = shape
    o = build_oracle(42)
    m = vb.Unet("resnet34")
    m.load_state_dict(o.state_dict(), strict=True)
    m = m.cuda().train()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, 3, H, W, generator=g).cuda()
    y = (torch.rand(N, 1, H, W, generator=g) < 0.2).float().cuda()
    logits = m(x)
    logits.retain_grad()
    vb.losses.BCEDiceLoss()(logits, y).backward()
    dlogits = logits.grad
    T = _debug_tensors(m, N)
    P = {k: v.detach() for k, v in m.named_parameters()}
    G = {k: v.grad.detach() for k, v in m.named_parameters()}
    ck = _Checker()
    TB, TP = 1e-2, 2e-3

    def bn_backward(conv, bn, dA, relu):
        """closed-form BatchNorm(+ReLU) backward from the library's own dA, z, a, mean, invstd"""
        z, a, mean, invstd = T[conv + "/z"], T[conv + "/a"], T[conv + "/mean"], T[conv + "/invstd"]
        gg = dA * (a > 0) if relu else dA
        zhat = (z - mean.view(1, -1, 1, 1)) * invstd.view(1, -1, 1, 1)
        cnt = z.numel() / z.shape[1]
        s1, s2 = gg.sum((0, 2, 3)), (gg * zhat).sum((0, 2, 3))
        k1 = (P[bn + ".weight"] * invstd).view(1, -1, 1, 1)
        dz = k1 * (gg - (s1 / cnt).view(1, -1, 1, 1) - zhat * (s2 / cnt).view(1, -1, 1, 1))
        ck.add(conv + " dz (BN backward)", T[conv + "/dz"], dz, TB)
        ck.add(bn + ".weight grad", G[bn + ".weight"], s2, TP)
        ck.add(bn + ".bias grad", G[bn + ".bias"], s1, TP)
        return gg

    def wgrad(conv, xin, stride, pad):
        w = P[conv]
        ref = conv2d_weight(xin, w.shape, T[conv + "/dz"], stride=stride, padding=pad)
        ck.add(conv + " grad (wgrad)", G[conv], ref, TP)

    def dgrad(conv, in_shape, stride, pad):
        return conv2d_input(in_shape, P[conv], T[conv + "/dz"], stride=stride, padding=pad)

    # ---- head
    ck.add("head d_in", T["head/din"], conv2d_input(T["head/in"].shape, P["segmentation_head.0.weight"], dlogits,
                                                    padding=1), TB)
    ck.add("segmentation_head.0.weight grad", G["segmentation_head.0.weight"],
           conv2d_weight(T["head/in"], (1, 16, 3, 3), dlogits, padding=1), TP)
    ck.add("segmentation_head.0.bias grad", G["segmentation_head.0.bias"], dlogits.sum().view(1), TP)

    # ---- decoder (last block first)
    enc_last = {1: "encoder.layer1.2", 2: "encoder.layer2.3", 3: "encoder.layer3.5", 4: "encoder.layer4.2"}
    skips = [enc_last[3] + ".conv2.weight/a", enc_last[2] + ".conv2.weight/a", enc_last[1] + ".conv2.weight/a",
             "encoder.conv1.weight/a", None]
    cups = [512, 256, 128, 64, 32]
    d_out = T["head/din"]
    for i in range(4, -1, -1):
        pre = f"decoder.blocks.{i}"
        c1, c2 = pre + ".conv1.0.weight", pre + ".conv2.0.weight"
        ck.add(pre + " conv2 dA (incoming)", T[c2 + "/dA"], d_out, 1e-6) if i < 4 else None
        bn_backward(c2, pre + ".conv2.1", T[c2 + "/dA"] if i < 4 else d_out, True)
        wgrad(c2, T[c1 + "/a"], 1, 1)
        ck.add(c2 + " dgrad -> conv1 dA", T[c1 + "/dA"], dgrad(c2, T[c1 + "/a"].shape, 1, 1), TB)
        bn_backward(c1, pre + ".conv1.1", T[c1 + "/dA"], True)
        low_name = (f"decoder.blocks.{i - 1}.conv2.0.weight" if i > 0 else "encoder.layer4.2.conv2.weight")
        low = T[low_name + "/a"]
        up = low.repeat_interleave(2, 2).repeat_interleave(2, 3)
        xin = torch.cat([up, T[skips[i]]], 1) if skips[i] else up
        wgrad(c1, xin, 1, 1)
        full = dgrad(c1, xin.shape, 1, 1)
        d_up = full[:, :cups[i]]
        d_low = d_up[:, :, 0::2, 0::2] + d_up[:, :, 0::2, 1::2] + d_up[:, :, 1::2, 0::2] + d_up[:, :, 1::2, 1::2]
        ck.add(c1 + " dgrad -> low-res input dA", T[low_name + "/dA"], d_low, TB)
        if skips[i]:
            ck.add(c1 + " dgrad -> d_skip", T[pre + "/d_skip"], full[:, cups[i]:], TB)
        d_out = T[low_name + "/dA"]

    # ---- encoder blocks (last first)
    nblocks = {1: 3, 2: 4, 3: 6, 4: 3}
    dskip_of_layer = {3: "decoder.blocks.0/d_skip", 2: "decoder.blocks.1/d_skip", 1: "decoder.blocks.2/d_skip"}
    for layer in (4, 3, 2, 1):
        for b in range(nblocks[layer] - 1, -1, -1):
            pre = f"encoder.layer{layer}.{b}"
            c1, c2 = pre + ".conv1.weight", pre + ".conv2.weight"
            has_ds = (b == 0 and layer > 1)
            if b > 0:
                xin_name = f"encoder.layer{layer}.{b - 1}.conv2.weight"
            elif layer > 1:
                xin_name = enc_last[layer - 1] + ".conv2.weight"
            else:
                xin_name = None
            xin = T[xin_name + "/a"] if xin_name else T["pool/out"]
            stride = 2 if has_ds else 1
            gg = bn_backward(c2, pre + ".bn2", T[c2 + "/dA"], True)
            ck.add(c2 + " g (masked block-output gradient)", T[c2 + "/g"], gg, TB)
            wgrad(c2, T[c1 + "/a"], 1, 1)
            ck.add(c2 + " dgrad -> conv1 dA", T[c1 + "/dA"], dgrad(c2, T[c1 + "/a"].shape, 1, 1), TB)
            bn_backward(c1, pre + ".bn1", T[c1 + "/dA"], True)
            wgrad(c1, xin, stride, 1)
            d_in = dgrad(c1, xin.shape, stride, 1)
            if has_ds:
                cd = pre + ".downsample.0.weight"
                bn_backward(cd, pre + ".downsample.1", T[c2 + "/g"], False)
                wgrad(cd, xin, 2, 0)
                d_in = d_in + dgrad(cd, xin.shape, 2, 0) + T[dskip_of_layer[layer - 1]]
            else:
                d_in = d_in + T[c2 + "/g"]
            got = T[xin_name + "/dA"] if xin_name else T["pool/dout"]
            ck.add(pre + " block input gradient", got, d_in, TB)

    # ---- stem: max-pool backward (+ decoder skip gradient) -> BN backward -> wgrad
    f1 = T["encoder.conv1.weight/a"].clone().requires_grad_()
    F.max_pool2d(f1, 3, 2, 1).backward(T["pool/dout"])
    ck.add("maxpool backward + d_skip -> stem dA", T["encoder.conv1.weight/dA"], f1.grad + T["decoder.blocks.3/d_skip"], TB)
    bn_backward("encoder.conv1.weight", "encoder.bn1", T["encoder.conv1.weight/dA"], True)
    xb = x.to(torch.bfloat16).float()
    wgrad("encoder.conv1.weight", xb, 2, 3)

    bad = ck.report()
    _persist(shape, ck.rows)
    assert not bad, bad[:5]
    assert m._ctx.device_error_flag() == 0


def _persist(shape, rows):
    """worst relative L2 per check kind -> gpurun_out/parity_r2_backward_local.json (copied to profiles/)"""
    import json
    import os

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out",
                        "parity_r2_backward_local.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    doc = json.load(open(path)) if os.path.exists(path) else {}
    worst = sorted(rows, key=lambda r: -r[1] / r[2])[:5]
    doc["x".join(map(str, shape))] = {
        "checks": len(rows), "over_tolerance": sum(1 for r in rows if not r[1] <= r[2]),
        "max_rel_l2_bf16_tensors": max(r[1] for r in rows if r[2] >= 1e-2),
        "max_rel_l2_fp32_param_grads": max(r[1] for r in rows if r[2] < 1e-2 and r[2] > 1e-5),
        "worst": [{"what": w, "rel_l2": r, "tol": t} for w, r, t, _ in worst]}
    json.dump(doc, open(path, "w"), indent=1)
