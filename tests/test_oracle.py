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
