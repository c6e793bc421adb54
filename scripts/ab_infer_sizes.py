import sys, os, torch
sys.path.insert(0, os.getcwd())
import vickers_hardness_unet_b200 as vb
dev = torch.device("cuda", 0)
torch.manual_seed(42)
model = vb.Unet("resnet34").to(dev).eval()
for (B, S) in ((8, 1024), (32, 512)):
    x = torch.randn(B, 3, S, S, device=dev)
    with torch.no_grad():
        for _ in range(5): model(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30): model(x)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    print(f"B={B} S={S}: {ms:.4f} ms  {B/ms*1e3:.0f} img/s")
