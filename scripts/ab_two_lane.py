"""Experiment (GPU): does running the batch as two half-batch pipelines on two streams (tails of one launch overlapping
the heads of the other stream's launch) beat one batch-32 pipeline?  Two model instances = two contexts = two arenas."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import vickers_hardness_unet_b200 as vb

torch.manual_seed(0)
dev = torch.device("cuda")
B, S = int(os.environ.get("B", "32")), 512
ms = [vb.Unet("resnet34").to(dev).eval() for _ in range(2)]
ms[1].load_state_dict(ms[0].state_dict())
g = torch.Generator(device=dev).manual_seed(1)
xs = [torch.randn(B, 3, S, S, device=dev, generator=g) for _ in range(2)]
st = [torch.cuda.Stream(), torch.cuda.Stream()]


def one(n):
    for i in range(n):
        with torch.no_grad():
            ms[0](xs[i & 1])


def two(n, h):
    for i in range(n):
        x = xs[i & 1]
        for k in range(2):
            with torch.cuda.stream(st[k]), torch.no_grad():
                ms[k](x[k * h:(k + 1) * h])


def timeit(fn, n):
    fn(5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    fn(n)
    for s in st:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return B * n / (e0.elapsed_time(e1) * 1e-3), B * n / (time.perf_counter() - t0)


for rep in range(2):
    print("one pipeline  batch %d: %.0f img/s (wall %.0f)" % ((B,) + timeit(one, 40)))
    print("two pipelines batch %d: %.0f img/s (wall %.0f)" % ((B // 2,) + timeit(lambda n: two(n, B // 2), 40)))
