"""Per-tensor gradient comparison against the fp32 CPU oracle, printed in backward order (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import vickers_hardness_unet_b200 as vb
from oracle import OracleDiceLoss, build_oracle

N, H, W = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (2, 64, 64)))
o = build_oracle(42).train()
m = vb.Unet("resnet34")
m.load_state_dict(o.state_dict(), strict=True)
m = m.cuda().train()
g = torch.Generator().manual_seed(7)
x = torch.randn(N, 3, H, W, generator=g)
y = (torch.rand(N, 1, H, W, generator=g) < 0.2).float()
lo = o(x)
loss_o = F.binary_cross_entropy_with_logits(lo, y) + OracleDiceLoss()(lo, y)
loss_o.backward()
from oracle.bf16_emulation import Bf16EmulatedTrainUnet
emu = Bf16EmulatedTrainUnet(o)
le = emu(x)
loss_e = F.binary_cross_entropy_with_logits(le, y) + OracleDiceLoss()(le, y)
loss_e.backward()
lg = m(x.cuda())
loss_g = vb.losses.BCEDiceLoss()(lg, y.cuda())
loss_g.backward()
torch.cuda.synchronize()
print("device error flag", m._ctx.device_error_flag())
print("logits: cuda-vs-emu mean", float((lg.detach().cpu() - le.detach()).abs().mean()), "cuda-vs-fp32",
      float((lg.detach().cpu() - lo.detach()).abs().mean()), "loss cuda/emu/fp32", float(loss_g), float(loss_e), float(loss_o))
rows = []
for (n, p1), (_, pe), p2 in zip(o.named_parameters(), emu.o.named_parameters(), m.parameters()):
    a, b, e = p1.grad, p2.grad.detach().cpu(), pe.grad
    rel = float((e - b).norm() / (e.norm() + 1e-20))
    cos = float((e * b).sum() / (e.norm() * b.norm() + 1e-20))
    rel32 = float((a - e).norm() / (a.norm() + 1e-20))
    rows.append((n, rel, cos, rel32, float(e.norm()), float(b.norm())))
for n, rel, cos, rel32, na, nb in reversed(rows):
    print(f"{n:48s} cuda-vs-emu rel {rel:8.4f} cos {cos:8.5f} | emu-vs-fp32 rel {rel32:8.4f} | |emu| {na:.3e} |got| {nb:.3e}")
