"""N train steps of BASELINE configs[2] (batch 16 @512x512) and nothing else: the command ncu wraps for the train-step
launch list / kernel captures (scripts/gpu_ncu_train.sh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vickers_hardness_unet_b200 as vb

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda", 0)
torch.manual_seed(42)
model = vb.Unet("resnet34").to(dev).train()
opt = vb.FusedAdamW(model, lr=5e-5, weight_decay=1e-4)
crit = vb.losses.BCEDiceLoss()
x = torch.randn(16, 3, 512, 512, device=dev)
y = (torch.rand(16, 1, 512, 512, device=dev) < 0.05).float()
for i in range(steps):
    opt.zero_grad(set_to_none=True)
    loss = crit(model(x), y)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("loss", float(loss))
