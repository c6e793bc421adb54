"""Generates tests/golden/oracle_golden.npz from the CPU oracle (run in the build container, output is committed).

There is no upstream golden data for this path (the reference has no tests and its checkpoints are missing blobs,
SURVEY.md section 8c), so these vectors pin the ORACLE against drift: same torch build => same seeded weights =>
same logits.  They also give the GPU tests a small fixed case that does not need the oracle to run.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import OracleDiceLoss, build_oracle  # noqa: E402


def main():
    torch.set_num_threads(4)
    m = build_oracle(42).eval()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 64, 96, generator=g)
    with torch.no_grad():
        logits = m(x)
    sd = m.state_dict()
    keys = list(sd.keys())
    h = hashlib.sha256()
    for k in keys:
        h.update(sd[k].numpy().tobytes())
    # train-mode forward + loss on a second input (BatchNorm batch statistics)
    m.train()
    y = (torch.rand(2, 1, 64, 96, generator=g) < 0.3).float()
    lt = m(x)
    bce = torch.nn.functional.binary_cross_entropy_with_logits(lt, y)
    dice = OracleDiceLoss()(lt, y)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.npz")
    np.savez_compressed(out, x=x.numpy(), logits_eval=logits.numpy(), y=y.numpy(),
                        logits_train=lt.detach().numpy(), bce=float(bce), dice=float(dice),
                        weights_sha256=np.frombuffer(h.digest(), dtype=np.uint8))
    with open(os.path.join(os.path.dirname(out), "state_dict_keys.json"), "w") as f:
        json.dump({k: list(sd[k].shape) for k in keys}, f, indent=0)
    print("wrote", out, "eval mean|logit|", float(logits.abs().mean()), "bce", float(bce), "dice", float(dice))


if __name__ == "__main__":
    main()
