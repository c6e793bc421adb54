"""N inference steps of BASELINE configs[1] (batch 32 @512x512) and nothing else: the command ncu wraps for the
inference launch list / kernel captures (scripts/gpu_ncu_r2.sh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vickers_hardness_unet_b200 as vb

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda", 0)
torch.manual_seed(42)
model = vb.Unet("resnet34", precision=os.environ.get("UNET_PRECISION", "bf16")).to(dev).eval()
xs = [torch.randn(32, 3, 512, 512, device=dev) for _ in range(2)]
with torch.no_grad():
    for i in range(steps):
        y = model(xs[i & 1])
torch.cuda.synchronize()
print("logits", float(y.abs().mean()), "launches per step", model._ctx.lib.unetb200_infer_launch_count(model._ctx.handle, 32))
