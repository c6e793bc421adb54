"""CPU tests of the oracle: structure, golden vectors, torchvision pin, DiceLoss closed form and edge cases."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import OracleDiceLoss, OracleUnet, build_oracle, oracle_train_step

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def oracle():
    torch.set_num_threads(4)
    return build_oracle(42)


def test_structure(oracle):
    sd = oracle.state_dict()
    assert len(sd) == 278
    assert sum(p.numel() for p in oracle.parameters()) == 24_436_369  # SURVEY.md section 8a census
    assert len(list(oracle.parameters())) == 140
    keys = json.load(open(os.path.join(GOLD, "state_dict_keys.json")))
    assert list(sd.keys()) == list(keys.keys())
    assert all(list(sd[k].shape) == v for k, v in keys.items())
    assert not any(k.startswith("encoder.fc") for k in sd)
    assert sd["decoder.blocks.0.conv1.0.weight"].shape == (256, 768, 3, 3)
    assert sd["decoder.blocks.3.conv1.0.weight"].shape == (32, 128, 3, 3)
    assert sd["decoder.blocks.4.conv1.0.weight"].shape == (16, 32, 3, 3)
    assert sd["segmentation_head.0.weight"].shape == (1, 16, 3, 3) and sd["segmentation_head.0.bias"].shape == (1,)


def test_golden_vectors(oracle):
    g = np.load(os.path.join(GOLD, "oracle_golden.npz"))
    h = hashlib.sha256()
    for k, v in oracle.state_dict().items():
        h.update(v.numpy().tobytes())
    assert np.array_equal(np.frombuffer(h.digest(), dtype=np.uint8), g["weights_sha256"]), "seeded init drifted"
    oracle.eval()
    with torch.no_grad():
        out = oracle(torch.from_numpy(g["x"]))
    assert np.allclose(out.numpy(), g["logits_eval"], rtol=1e-4, atol=1e-4)


def test_encoder_is_torchvision_resnet34(oracle):
    """Pin: the encoder must reproduce torchvision.models.resnet34 stage by stage (resnet.py:266-276)."""
    import torchvision

    tv = torchvision.models.resnet34(weights=None)
    sd = {k[len("encoder."):]: v for k, v in oracle.state_dict().items() if k.startswith("encoder.")}
    missing, unexpected = tv.load_state_dict(sd, strict=False)
    assert set(missing) == {"fc.weight", "fc.bias"} and not unexpected
    tv.eval()
    oracle.eval()
    x = torch.randn(1, 3, 64, 64, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        feats = oracle.encoder(x)
        t = tv.relu(tv.bn1(tv.conv1(x)))
        assert torch.equal(t, feats[1])
        t = tv.layer1(tv.maxpool(t))
        assert torch.equal(t, feats[2])
        t = tv.layer4(tv.layer3(tv.layer2(t)))
        assert torch.equal(t, feats[5])
    assert [f.shape[1] for f in feats] == [3, 64, 64, 128, 256, 512]


def test_decoder_matches_explicit_formula(oracle):
    """Decoder block == conv(cat(nearest2x(x), skip)) -> BN -> ReLU twice, written out by hand."""
    oracle.eval()
    blk = oracle.decoder.blocks[2]
    g = torch.Generator().manual_seed(5)
    x, skip = torch.randn(1, 128, 8, 8, generator=g), torch.randn(1, 64, 16, 16, generator=g)
    with torch.no_grad():
        up = x.repeat_interleave(2, 2).repeat_interleave(2, 3)
        t = torch.cat([up, skip], 1)
        for seq in (blk.conv1, blk.conv2):
            t = F.relu(F.batch_norm(F.conv2d(t, seq[0].weight, None, 1, 1), seq[1].running_mean, seq[1].running_var,
                                    seq[1].weight, seq[1].bias, False, 0.1, 1e-5))
        assert torch.allclose(blk(x, skip), t, atol=1e-6)


def test_dice_closed_form_and_edges():
    d = OracleDiceLoss()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 1, 8, 8, generator=g, dtype=torch.float64)
    y = (torch.rand(3, 1, 8, 8, generator=g) < 0.4).double()
    p = torch.sigmoid(x)
    ref = 1 - 2 * (p * y).sum() / (p.sum() + y.sum())
    assert abs(float(d(x, y)) - float(ref)) < 1e-12
    # all-background batch: loss and gradient are exactly zero
    xz = torch.randn(2, 1, 4, 4, requires_grad=True)
    lz = d(xz, torch.zeros(2, 1, 4, 4))
    lz.backward()
    assert float(lz) == 0.0 and float(xz.grad.abs().max()) == 0.0
    # one score for the whole batch, not a per-image mean
    per_image = torch.stack([1 - 2 * (p[i] * y[i]).sum() / (p[i].sum() + y[i].sum()) for i in range(3)]).mean()
    assert abs(float(per_image) - float(ref)) > 1e-6


def test_train_step_decreases_loss():
    torch.set_num_threads(4)
    m = build_oracle(42).train()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-4)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 3, 64, 64, generator=g)
    y = (torch.rand(2, 1, 64, 64, generator=g) < 0.2).float()
    losses = [oracle_train_step(m, opt, x, y) for _ in range(4)]
    assert losses[-1] < losses[0]
    assert int(m.encoder.bn1.num_batches_tracked) == 4


def test_round_mantissa_matches_torch_casts():
    """oracle.bf16_emulation.round_mantissa (the storage-precision sweep of tests/test_gpu_trained.py): 7 explicit mantissa
    bits == bfloat16, 10 == float16 (inside half's exponent range), 23 == identity; round-to-nearest-even."""
    from oracle.bf16_emulation import round_mantissa
    g = torch.Generator().manual_seed(0)
    t = torch.randn(4096, generator=g) * torch.logspace(-3, 3, 4096)
    assert torch.equal(round_mantissa(t, 7), t.to(torch.bfloat16).float())
    n = t[t.abs() > 1e-3]                                         # half is subnormal below 6.1e-5: mantissa bits only
    assert torch.equal(round_mantissa(n, 10), n.to(torch.float16).float())
    assert torch.equal(round_mantissa(t, 23), t)
    tie = torch.tensor([1.0 + 2.0 ** -8, 1.0 + 3 * 2.0 ** -8])   # exactly half way between two bf16 values
    assert torch.equal(round_mantissa(tie, 7), torch.tensor([1.0, 1.0 + 2.0 ** -6]))


def test_emulated_forward_with_identity_rounding_is_the_oracle_forward():
    """emulated_forward(rnd = identity) re-implements the oracle's forward with the CUDA path's structure (parity-folded
    decoder conv1, explicit block wiring): it must reproduce OracleUnet.forward, and its taps name every activation the
    CUDA path materialises (47 conv outputs + the max-pool)."""
    from oracle.bf16_emulation import emulated_forward
    o = build_oracle(42).eval()
    x = torch.randn(1, 3, 64, 96, generator=torch.Generator().manual_seed(2))
    taps = {}
    with torch.no_grad():
        a = o(x)
        b = emulated_forward(o, x, False, rnd=lambda t: t, taps=taps)
    assert float((a - b).abs().max()) <= 1e-4 * float(a.abs().max()) + 1e-5
    assert len(taps) == 47 and "encoder.maxpool/out" in taps and "decoder.blocks.4.conv2.0.weight/out" in taps


def test_vickers_fixture_is_the_reference_dataset_split():
    """tests/golden/vickers_512.npz (built from /root/reference/data by tests/golden/make_vickers_fixture.py): 182
    annotated micrographs, 10 % validation (train.py:560-565), 512 x 512 after centred padding, binary masks with the
    foreground fraction the survey measured (mean 4.4 %)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import vickers_data as vd
    d = vd.load_vickers()
    assert d["train_u8"].shape == (164, 512, 512, 3) and d["val_u8"].shape == (18, 512, 512, 3)
    assert d["train_y"].shape == (164, 1, 512, 512) and set(d["val_y"].unique().tolist()) <= {0.0, 1.0}
    fg = float(torch.cat([d["train_y"], d["val_y"]]).mean())
    assert 0.025 < fg < 0.05, fg            # 4.4 % of the 410 x 512 image area = 3.5 % of the padded square
    x = vd.normalise(d["val_u8"][:1])
    assert x.shape == (1, 3, 512, 512) and abs(float(x[:, :, 200:300, 200:300].mean())) < 3.0
    sched = vd.batches(164, 16, 12, 1234)
    assert len(sched) == 12 and all(len(set(i.tolist())) == 16 for i, _ in sched)
    assert abs(vd.lr_at(30, 2000, 3e-4) - 3e-4 * 0.5 * (1 + __import__("math").cos(__import__("math").pi * 30 / 2000))) < 1e-12
