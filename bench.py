#!/usr/bin/env python
"""Benchmark of the U-Net(ResNet-34) hot path (BASELINE.json: 512x512 images/sec, infer + train step).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--size S]

One JSON line on stdout (rank 0).  N=1 workload = BASELINE.json configs[1]: batch-32 bf16 inference at 512x512.
A "step" is one forward pass of the whole network over one batch of synthetic inputs.
  value     images/s with the fp32 NCHW input batch already resident in HBM (CUDA events, max over ranks)
  e2e       same metric through the C-ABI host-buffer call (pinned uint8 frame -> H2D -> normalise + forward -> D2H of the
            uint8 mask); the fp32-frame form of the call is reported beside it
  roofline  the conv stack (all wconv / tconv / wpconv / igemm launches of one step) against the measured dense bf16 peak
            (MEASURED_PEAKS.json); roofline.per_layer bounds every launch by max(FLOPs / peak, min bytes / HBM copy bandwidth)
  cpu_baseline  the fp32 CPU oracle timed on this box's host cores on a bounded sample (rank 0, N=1)
`--impl reference` times the reference's CPU implementation of the path (the oracle port: smp is not installable
offline) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_FWD_512 = 62.512  # algorithmic forward GFLOP / image @512x512 (scales with H*W)


def workload_config(B, S, world):
    """`config` of the JSON line: the SAME dict for both arms (`--impl ours` and `--impl reference`)."""
    return {"workload": f"Unet(resnet34) {S}x{S} batch-{B}/GPU inference, random-init seed 42 (BASELINE configs[1])",
            "weights": "random-init seed 42", "input": "fp32 NCHW randn, 2 batches alternated",
            "l2": "per-step working set >> 126 MB L2 (inputs larger than L2)",
            "parallelism": f"batch-sharded replicas x{world}, no collective"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--profile-out", default="", help="write the per-launch timing table here (rank 0)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true", help="skip the stock PyTorch/cuDNN timing on this GPU")
    ap.add_argument("--train-batch", type=int, default=16, help="images per GPU per train step (BASELINE configs[2])")
    ap.add_argument("--no-train", action="store_true", help="skip the train-step measurement")
    ap.add_argument("--train-profile-out", default="", help="write the per-launch timing table of one train step here")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])), mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_rate(size: int, batch: int, budget_s: float, threads: int):
    """images/s of the fp32 oracle forward (eval, no_grad) on the host: bounded sample."""
    import torch
    from oracle import build_oracle
    torch.set_num_threads(threads)
    m = build_oracle(42).eval()
    x = torch.randn(batch, 3, size, size, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        m(x)
        t0 = time.perf_counter()
        m(x)
        one = time.perf_counter() - t0
        n = max(3, min(1000, int(budget_s / max(one, 1e-3))))  # ~budget_s seconds of CPU work
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            m(x)
            ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    return batch / med, n, med


def library_baseline(a, dev):
    """The "library bar" (BASELINE.md section 3): what the reference's user gets today by typing
    `.to("cuda")` (train.py:592, infer_pth_gui.py:92) — stock PyTorch eager on cuDNN, same B200, same run, same shapes.
    The smp model is stood in for by the oracle restatement (smp is not installable offline).  Variants:
      fp32            torch defaults (cuDNN may use TF32 for convs), NCHW            — infer_pth_gui.py:50 as written
      amp_fp16        autocast(float16) NCHW (+ GradScaler for the train step)      — train.py:431-449 as written
      bf16_cl         autocast(bfloat16) + channels_last                            — the strongest stock configuration
    CUDA events, 3 warm-up + 10 timed steps each.  Returns the `library_baseline` object of the JSON line."""
    import torch
    import torch.nn.functional as F
    from oracle import OracleDiceLoss, build_oracle

    B, TB, S = a.batch, a.train_batch, a.size
    g = torch.Generator(device=dev).manual_seed(7)

    def timeit(fn, n=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n

    out = {"what": "stock PyTorch eager (cuDNN %s) on the oracle restatement of smp.Unet, this GPU, this run" %
                   str(torch.backends.cudnn.version()), "infer": {}, "train": {}}
    variants = [("fp32", None, False), ("amp_fp16", torch.float16, False), ("bf16_cl", torch.bfloat16, True)]
    for tag, dt, cl in variants:
        m = build_oracle(42).to(dev).eval()
        x = torch.randn(B, 3, S, S, device=dev, generator=g)
        if cl:
            m = m.to(memory_format=torch.channels_last)
            x = x.contiguous(memory_format=torch.channels_last)

        def infer():
            with torch.no_grad(), torch.autocast("cuda", dtype=dt or torch.float16, enabled=dt is not None):
                return m(x)
        ms = timeit(infer)
        out["infer"][tag] = {"images_per_s": B / (ms * 1e-3), "ms_per_step": ms, "batch": B}
        del x
        m.train()
        opt = torch.optim.AdamW(m.parameters(), lr=5e-5, weight_decay=1e-4)
        scaler = torch.amp.GradScaler("cuda", enabled=dt is torch.float16)
        dice = OracleDiceLoss()
        xt = torch.randn(TB, 3, S, S, device=dev, generator=g)
        yt = (torch.rand(TB, 1, S, S, device=dev, generator=g) < 0.05).float()
        if cl:
            xt = xt.contiguous(memory_format=torch.channels_last)

        def train_step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=dt or torch.float16, enabled=dt is not None):
                lg = m(xt)
                loss = F.binary_cross_entropy_with_logits(lg.float(), yt) + dice(lg.float(), yt)
            scaler.scale(loss).backward()
            scaler.step(opt)
            scaler.update()
        ms = timeit(train_step)
        out["train"][tag] = {"images_per_s": TB / (ms * 1e-3), "ms_per_step": ms, "batch": TB}
        del m, opt, xt, yt
        torch.cuda.empty_cache()
    return out


def bench_train(a, dev, rank, world, barrier):
    """Train step of BASELINE configs[2]: batch 16/GPU @512x512, BCE+Dice, fused AdamW, bucketed NCCL all-reduce of the
    gradients overlapped with the backward when world > 1.  Returns the `train` sub-object of the JSON line."""
    import ctypes

    import torch
    import torch.distributed as dist
    import vickers_hardness_unet_b200 as vb

    B, S = a.train_batch, a.size
    torch.manual_seed(42)
    model = vb.Unet("resnet34", encoder_weights=None, in_channels=3, classes=1, activation=None).to(dev).train()
    if world > 1:
        vb.distributed.enable_data_parallel(model)
    # /root/reference/train.py:606, RECOMMENDED_CFG lr; with N > 1 the update runs bucket by bucket as the all-reduces land
    opt = vb.FusedAdamW(model, lr=5e-5, weight_decay=1e-4, overlap_allreduce=world > 1)
    crit = vb.losses.BCEDiceLoss()
    g = torch.Generator(device=dev).manual_seed(4321 + rank)
    xs = [torch.randn(B, 3, S, S, device=dev, generator=g) for _ in range(2)]
    ys = [(torch.rand(B, 1, S, S, device=dev, generator=g) < 0.05).float() for _ in range(2)]
    last = {}

    def step(x, y):
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x), y)
        loss.backward()
        opt.step()
        last["loss"] = loss
    steps, warm = max(3, a.steps // 2), max(3, min(a.warmup, 5))
    for i in range(warm):
        step(xs[i & 1], ys[i & 1])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(xs[i & 1], ys[i & 1])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # phase split of one step (CUDA events on the launch stream)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    opt.zero_grad(set_to_none=True)
    logits = model(xs[0])
    ev[1].record()
    loss = crit(logits, ys[0])
    ev[2].record()
    loss.backward()
    ev[3].record()
    opt.step()
    ev[4].record()
    torch.cuda.synchronize(dev)
    phases = {k: ev[i].elapsed_time(ev[i + 1]) for i, k in enumerate(["forward_ms", "loss_ms", "backward_ms", "adamw_ms"])}
    if a.train_profile_out and rank == 0:
        ctx = model._ctx
        ctx.lib.unetb200_profile_enable(ctx.handle, 1)
        step(xs[1], ys[1])
        ctx.check(ctx.lib.unetb200_profile_dump(ctx.handle, a.train_profile_out.encode()), "profile_dump")
        ctx.lib.unetb200_profile_enable(ctx.handle, 0)
    # end to end: pinned host batch -> H2D -> step -> loss.item() (the D2H sync of train.py:452).
    # vb.data.DevicePrefetcher: batch k+1 uploads on a side stream while batch k trains (every batch is still copied
    # host -> device once per step, inside the timed region).  Headline: uint8 frames + uint8 masks as a data loader
    # holds them after cv2.imread / letterbox (normalisation of train.py:108-112 on the device); beside it the
    # reference-style fp32 host tensors (4x the bytes).
    def e2e_run(xh, yh):
        class _Loader:   # a re-iterable loader, like the DataLoader of train.py:585 (one pass = one "epoch")
            n = 2

            def __iter__(self):
                for i in range(self.n):
                    yield xh[i & 1], yh[i & 1]

            def __len__(self):
                return self.n
        loader = _Loader()
        pf = vb.data.DevicePrefetcher(loader, dev)   # ONE prefetcher (its device slots are reused across epochs)
        for xd, yd in pf:
            step(xd, yd)
            last["loss"].item()
        loader.n = steps
        barrier()
        t0 = time.perf_counter()
        for xd, yd in pf:
            step(xd, yd)
            lv = last["loss"].item()
        torch.cuda.synchronize(dev)
        return time.perf_counter() - t0, lv
    e2e_f32_s, lv = e2e_run([x.cpu().pin_memory() for x in xs], [y.cpu().pin_memory() for y in ys])
    gh = torch.Generator().manual_seed(99 + rank)
    x8h = [torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8, generator=gh).pin_memory() for _ in range(2)]
    y8h = [y.cpu().to(torch.uint8).pin_memory() for y in ys]
    e2e_s, _ = e2e_run(x8h, y8h)
    if world > 1:
        t = torch.tensor([ms, e2e_s, e2e_f32_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s, e2e_f32_s = float(t[0]), float(t[1]), float(t[2])
    ctx = model._ctx
    nf, nb = ctypes.c_int(), ctypes.c_int()
    ctx.lib.unetb200_train_launch_count(ctx.handle, B, ctypes.byref(nf), ctypes.byref(nb))
    dp_par = None
    if world > 1:
        # multi-GPU correctness inside the bench line: reduced gradients vs the sliced-average emulation on rank 0
        rel, mx, same = vb.distributed.dp_gradient_parity(dev, rank, world)
        dp_par = {"dp_parity_rel_l2": rel, "dp_parity_max_abs": mx, "dp_grads_identical_across_ranks": same}
    pk, pk_kind = peaks()
    gflop = 186.3 * (S * S) / (512 * 512)  # algorithmic train-step GFLOP / image
    val = world * B * steps / (ms * 1e-3)
    tf = gflop * B * steps / (ms * 1e-3) / 1e3
    return {"metric": "images_per_sec_train_512", "value": val, "unit": "images/s", "ms_per_step": ms / steps,
            "steps": steps, "batch_per_gpu": B, "loss": lv, **(dp_par or {}),
            "workload": f"Unet(resnet34) {S}x{S} train step BCE+Dice + fused AdamW, batch {B}/GPU (BASELINE configs[2])",
            "parallelism": f"dp{world}: bucketed NCCL all-reduce (4 buckets) overlapped with backward" if world > 1
            else "single GPU, no collective",
            "e2e": {"value": world * B * steps / e2e_s, "unit": "images/s",
                    "h2d_bytes_per_step": B * 4 * S * S, "d2h_bytes_per_step": 4,
                    "call": "DevicePrefetcher(pinned uint8 HWC frames + uint8 masks) -> model(frames) -> BCEDiceLoss -> "
                            "backward -> FusedAdamW.step -> loss.item()",
                    "fp32_frames": {"value": world * B * steps / e2e_f32_s, "h2d_bytes_per_step": B * 4 * S * S * 4,
                                    "call": "same with host-normalised fp32 NCHW images + fp32 masks (train.py:428-434)"}},
            "phases": phases, "gpu_launches_per_step": nf.value + nb.value + 4,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": tf / pk["bf16_tflops_sustained"], "peak_kind": pk_kind,
                         "kernel": "whole train step (fwd + dgrad + wgrad algorithmic FLOPs)"}}


def run_reference(a, rank):
    """Reference arm: the reference's own CPU path for this metric (oracle port), all host threads, rank 0 only."""
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    sb = a.batch  # the full batch of the workload: one reference step == one step of our arm
    torch.set_num_threads(threads)
    from oracle import build_oracle
    m = build_oracle(42).eval()
    x = torch.randn(sb, 3, a.size, a.size, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        for _ in range(max(1, min(a.warmup, 3))):
            m(x)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            m(x)
        dt = time.perf_counter() - t0
    val = sb * a.steps / dt
    sample = (f"all {sb} images of every step, {a.steps} steps after {max(1, min(a.warmup, 3))} warm-up, fp32 oracle "
              f"forward (eval, no_grad) on the host, {threads} torch threads")
    print(json.dumps({
        "impl": "reference", "metric": "images_per_sec_infer_512", "value": val, "unit": "images/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a.batch, a.size, a.gpus),
        "arm": "reference CPU path: fp32 PyTorch eager on the host cores (infer_pth_gui.py:50-51 on the oracle port)",
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        return run_reference(a, rank)

    import torch
    import torch.distributed as dist
    import vickers_hardness_unet_b200 as vb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B, S = a.batch, a.size
    model = vb.Unet("resnet34", encoder_weights=None, in_channels=3, classes=1, activation=None)
    torch.manual_seed(42)
    model = model.to(dev).eval()
    # two input batches, alternated (each step's working set, ~0.18 GB/image of activations, is far beyond L2)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    xs = [torch.randn(B, 3, S, S, device=dev, generator=g) for _ in range(2)]
    logits = None

    def step(i):
        nonlocal logits
        with torch.no_grad():
            logits = model(xs[i & 1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(a.warmup):
        step(i)
    barrier()
    ctx = model._ctx
    launches = ctx.lib.unetb200_infer_launch_count(ctx.handle, B)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(a.steps):
        step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * a.steps / (ms * 1e-3)

    # ---- sustained: the same step back to back for >= 3 s (clocks sampled by the sampler that is still running), so that
    # a fraction of the SUSTAINED peak in MEASURED_PEAKS.json is like for like
    n_sus = max(a.steps, int(3000.0 / (ms / a.steps)) + 1)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s0.record()
    for i in range(n_sus):
        step(i)
    s1.record()
    barrier()
    sus_ms = s0.elapsed_time(s1)
    if world > 1:
        t = torch.tensor([sus_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sus_ms = float(t.item())
    sustained = {"value": world * B * n_sus / (sus_ms * 1e-3), "unit": "images/s", "steps": n_sus,
                 "seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / n_sus}

    # ---- end to end through the C-ABI host-buffer entry points (pinned host buffers; every step's H2D copy of its
    # input and D2H copy of its result are inside the timed region).  Headline: the two-slot submit / wait form a caller
    # streaming frames uses (request k+1 uploads while request k computes); also the single blocking call and the
    # uint8-frame form (pre-processing fused into the input pack).
    xh = [x.cpu().pin_memory() for x in xs]
    mask_h = [torch.empty(B, 1, S, S, dtype=torch.uint8).pin_memory() for _ in range(2)]
    x8h = [torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]

    def timed(fn_submit, fn_drain):
        for i in range(max(2, min(a.warmup, 3))):
            fn_submit(i)
        fn_drain()
        barrier()
        t0 = time.perf_counter()
        for i in range(a.steps):
            fn_submit(i)
        fn_drain()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return world * B * a.steps / dt

    def pipe_submit(src):
        def f(i):
            model.wait_host(i & 1)          # the slot's previous request (step i-2) has fully landed in mask_h[i & 1]
            model.submit_host(i & 1, src[i & 1], mask_out=mask_h[i & 1])
        return f

    def drain():
        model.wait_host(0)
        model.wait_host(1)

    def sync_call(i):
        model.submit_host(0, xh[i & 1], mask_out=mask_h[0])
        model.wait_host(0)

    model.submit_host(0, xh[0], mask_out=mask_h[0])
    model.wait_host(0)
    e2e_val = timed(pipe_submit(xh), drain)
    e2e_sync = timed(sync_call, lambda: None)
    e2e_u8 = timed(pipe_submit(x8h), drain)
    # ---- the fp16-operand build of the same kernels (inference only): same tensor-core rate, 10 mantissa bits
    f16 = None
    if rank == 0 and world == 1:
        try:
            m16 = vb.Unet("resnet34", encoder_weights=None, in_channels=3, classes=1, activation=None, precision="fp16")
            m16.load_state_dict(model.state_dict())
            m16 = m16.to(dev).eval()
            with torch.no_grad():
                for i in range(3):
                    m16(xs[i & 1])
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(dev)
                f0.record()
                for i in range(a.steps):
                    m16(xs[i & 1])
                f1.record()
                torch.cuda.synchronize(dev)
            f16 = {"value": B * a.steps / (f0.elapsed_time(f1) * 1e-3), "unit": "images/s",
                   "what": "Unet(precision='fp16'): libunetb200_f16.so, IEEE-half activations / operands"}
            del m16
            torch.cuda.empty_cache()
        except Exception as e:  # the variant is optional evidence, never the headline
            f16 = {"error": str(e)[:200]}
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-launch timing of one step (CUDA events on the launch stream) -> conv-stack roofline
    pk, pk_kind = peaks()
    prof = None
    if hasattr(ctx.lib, "unetb200_profile_infer"):
        from vickers_hardness_unet_b200.profile import profile_infer
        prof = profile_infer(model, xs[0], reps=3)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        thr = os.cpu_count() or 1
        rate, n, med = cpu_oracle_rate(S, 1, 12.0, thr)
        cpu = {"value": rate, "unit": "images/s", "cores": thr, "kind": "port",
               "sample": f"fp32 oracle forward, batch 1 @ {S}x{S}, median of {n} runs ({med * 1e3:.0f} ms each), "
                         f"{thr} torch threads on {os.cpu_count()} host cores"}
    train = None
    lib_bar = None
    if not a.no_train:
        del xs, xh, x8h
        torch.cuda.empty_cache()
        train = bench_train(a, dev, rank, world, barrier)
    if rank == 0 and world == 1 and not a.no_library_baseline:
        lib_bar = library_baseline(a, dev)
    if rank == 0:
        gflop = GFLOP_FWD_512 * (S * S) / (512 * 512)
        whole_tf = gflop * B * a.steps / (ms * 1e-3) / 1e3  # whole step, per GPU
        roof = {"bound": "tensor", "achieved": whole_tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": whole_tf / pk["bf16_tflops_sustained"], "traffic": None, "peak_kind": pk_kind,
                "kernel": "whole forward step (no per-launch breakdown available)"}
        if prof is not None:
            conv_ms = prof["igemm_ms"]
            conv_tf = gflop * B / (conv_ms * 1e-3) / 1e3
            roof.update({"achieved": conv_tf, "frac": conv_tf / pk["bf16_tflops_sustained"],
                         "kernel": "conv stack = all ub::wconv_kernel / tconv_kernel / igemm_kernel launches of one step (algorithmic FLOPs)",
                         "kernel_ms_per_step": conv_ms, "kernel_share_of_step": conv_ms / prof["total_ms"],
                         "launches_per_step": prof["n_igemm"]})
            from vickers_hardness_unet_b200.profile import infer_layer_work, layer_roofline
            roof["per_layer"] = layer_roofline(prof["rows"], infer_layer_work(model, B, S, S),
                                               pk["bf16_tflops_sustained"], pk["hbm_gbs"])
            roof["per_layer"]["note"] = ("each conv launch bounded by max(FLOPs / measured bf16 peak, min HBM bytes / "
                                         "measured copy bandwidth); CUDA-event time per launch")
            if a.profile_out:
                with open(a.profile_out, "w") as f:
                    f.write(prof["table"])
        # DRAM traffic of the conv stack per step from the committed ncu capture (same workload only)
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "infer_dram_traffic_b32_512.json")
        if B == 32 and S == 512 and os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            roof["traffic"] = tj["conv_stack_dram_bytes_per_step"]
            roof["traffic_note"] = "bytes per step (sum over the conv launches), ncu dram__bytes_read+write: " + tj["source"]
        out = {
            "metric": "images_per_sec_infer_512", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(B, S, world),
            "arm": "B200 path: bf16 operands, fp32 accumulate (libunetb200.so)",
            # headline end-to-end call: the uint8 camera frame the reference's callers hold (cv2.imread / letterbox,
            # infer_pth_gui.py:45-46) goes in, the uint8 mask comes out; BGR->RGB, /255, (x-mean)/std run on the device
            "e2e": {"value": e2e_u8, "unit": "images/s", "h2d_bytes_per_step": B * 3 * S * S,
                    "d2h_bytes_per_step": B * S * S,
                    "call": "unetb200_infer_host_u8_submit/_wait, 2 slots (pinned uint8 HWC frames in, normalise on "
                            "device, uint8 mask out)",
                    "fp32_frames": {"value": e2e_val, "h2d_bytes_per_step": B * 3 * S * S * 4,
                                    "call": "unetb200_infer_host_submit/_wait (pinned fp32 NCHW in, host-normalised as "
                                            "infer_pth_gui.py:46-49 does, uint8 mask out)"},
                    "blocking_call": {"value": e2e_sync, "call": "unetb200_infer_host (fp32 frames, one request at a time)"}},
            "gpu_launches": launches * a.steps, "roofline": roof, "clocks": clocks, "sustained": sustained,
        }
        sus_tf = gflop * sustained["value"] / world / 1e3
        out["sustained"]["whole_step_tflops"] = sus_tf
        out["sustained"]["frac_of_sustained_peak"] = sus_tf / pk["bf16_tflops_sustained"]
        if f16:
            out["fp16_operands"] = f16
        if lib_bar:
            lib_bar["speedup_vs_best_library"] = {
                "infer": value / max(v["images_per_s"] for v in lib_bar["infer"].values()),
                "train": (train["value"] / max(v["images_per_s"] for v in lib_bar["train"].values())) if train else None}
            out["library_baseline"] = lib_bar
        if cpu:
            out["cpu_baseline"] = cpu
        if train:
            out["train"] = train
            out["gpu_launches"] += train["gpu_launches_per_step"] * train["steps"]
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
