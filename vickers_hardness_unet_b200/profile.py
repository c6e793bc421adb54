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
