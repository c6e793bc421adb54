"""Per-launch timing of one inference step (CUDA events between launches on the launch stream)."""
from __future__ import annotations

import ctypes as C

import torch


def profile_infer(model, x: torch.Tensor, reps: int = 3):
    ctx = model._ctx
    N = x.shape[0]
    stream = torch.cuda.current_stream(x.device).cuda_stream
    logits = torch.empty((N, 1, x.shape[2], x.shape[3]), dtype=torch.float32, device=x.device)
    cap = 256
    ms = (C.c_float * cap)()
    isg = (C.c_int * cap)()
    n = C.c_int()
    acc = None
    for _ in range(reps):
        ctx.check(ctx.lib.unetb200_profile_infer(ctx.handle, x.data_ptr(), logits.data_ptr(), N, stream, ms, isg, cap,
                                                 C.byref(n)), "profile_infer")
        cur = [ms[i] for i in range(n.value)]
        acc = cur if acc is None else [min(a, b) for a, b in zip(acc, cur)]
    name = C.create_string_buffer(128)
    rows, ig_ms, n_ig = [], 0.0, 0
    for i in range(n.value):
        ctx.lib.unetb200_profile_name(ctx.handle, N, i, name, 128)
        rows.append((name.value.decode(), acc[i], isg[i]))
        if isg[i]:
            ig_ms += acc[i]
            n_ig += 1
    total = sum(acc)
    table = "launch,kernel,ms,share\n" + "".join(
        f"{nm},{'igemm' if g else 'elementwise'},{t:.4f},{t / total:.4f}\n" for nm, t, g in rows)
    return {"igemm_ms": ig_ms, "total_ms": total, "n_igemm": n_ig, "rows": rows, "table": table}


def infer_layer_work(model, N: int, H: int, W: int):
    """Algorithmic FLOPs and minimum HBM bytes of every conv launch of the inference step, keyed by the launch names of
    `profile_infer` (layer name, with [up] / [skip] for a decoder conv1 split into its two sources).  bf16 activations:
    bytes = input read once + output written once (+ the bf16 partial a [skip] launch reads and an [up] launch writes,
    + the residual an encoder conv2 reads); weights are negligible.  FLOPs are those of the reference's dense conv on
    the up-sampled + concatenated tensor (2 * cout * cin * k*k * Hout * Wout), whatever the kernel actually issues."""
    sd = model.state_dict()
    enc_ch = {"layer1": 64, "layer2": 128, "layer3": 256, "layer4": 512}
    enc_div = {"layer1": 4, "layer2": 8, "layer3": 16, "layer4": 32}
    dec_up = [512, 256, 128, 64, 32]
    dec_div = [16, 8, 4, 2, 1]
    work = {}

    def put(name, cin, cout, k, ho, wo, in_px, extra_out_ch=0):
        flops = 2.0 * cout * cin * k * k * ho * wo * N
        byts = 2.0 * N * (cin * in_px + cout * ho * wo + extra_out_ch * ho * wo)
        work[name] = (flops, byts)

    for key, w in sd.items():
        if w.dim() != 4:
            continue
        cout, cin, k, _ = w.shape
        if key == "encoder.conv1.weight":
            put("encoder.conv1", 3, 64, 7, H // 2, W // 2, H * W * 4 / 3)      # packed 4-channel input
        elif key.startswith("encoder.layer"):
            lay = key.split(".")[1]
            d = enc_div[lay]
            ho, wo = H // d, W // d
            stride2 = (".0.conv1." in key or "downsample" in key) and lay != "layer1"
            in_px = ho * wo * (4 if stride2 else 1)
            res = cout if ".conv2." in key else 0                                # identity / downsample branch read
            put(key, cin, cout, k, ho, wo, in_px, res)
        elif key.startswith("decoder.blocks."):
            b = int(key.split(".")[2])
            ho, wo = H // dec_div[b], W // dec_div[b]
            if ".conv1." in key:
                cup = dec_up[b]
                cskip = cin - cup
                if cskip:
                    put(key + "[up]", cup, cout, 3, ho, wo, ho * wo / 4)
                    put(key + "[skip]", cskip, cout, 3, ho, wo, ho * wo, cout)   # reads the bf16 partial
                put(key, cin, cout, 3, ho, wo, ho * wo * (cup / 4 + cskip) / cin)
            else:
                put(key, cin, cout, 3, ho, wo, ho * wo)
        elif key.startswith("segmentation_head"):
            flops = 2.0 * 1 * 16 * 9 * H * W * N
            work["segmentation_head"] = (flops, 2.0 * N * 16 * H * W + 4.0 * N * H * W)
    return work


def layer_roofline(rows, work, tflops_peak: float, hbm_gbs: float):
    """rows: (launch name, ms, is_conv) of profile_infer.  Per conv launch the attainable time is
    max(FLOPs / tensor peak, bytes / HBM peak); returns sum(attainable) / sum(measured) over the conv launches and the
    split of the measured time between launches whose bound is the tensor pipe and those whose bound is HBM."""
    att = meas = t_tensor = t_hbm = 0.0
    missing = []
    for name, ms, is_conv in rows:
        if not is_conv:
            continue
        if name not in work:
            missing.append(name)
            continue
        f, b = work[name]
        tf, tb = f / (tflops_peak * 1e12) * 1e3, b / (hbm_gbs * 1e9) * 1e3
        att += max(tf, tb)
        meas += ms
        if tf >= tb:
            t_tensor += ms
        else:
            t_hbm += ms
    return {"frac_of_per_layer_roofline": att / meas if meas else None, "attainable_ms": att, "measured_ms": meas,
            "ms_in_tensor_bound_layers": t_tensor, "ms_in_hbm_bound_layers": t_hbm, "unmatched_launches": missing}
