"""ctypes binding of libunetb200.so (C ABI declared in include/unetb200.h).

The product path has no fallback: if the shared library is missing this module raises at first use.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libunetb200.so"
LIB_PATHS = {"bf16": LIB_PATH, "fp16": _HERE / "libunetb200_f16.so"}   # fp16: same kernels, IEEE-half operands (inference)
_libs = {}


class UnetB200Error(RuntimeError):
    pass


def _proto(lib):
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_longlong, C.c_float
    P = C.POINTER
    sigs = {
        "unetb200_create": (i32, [P(vp), i32, i32, i32, i32]),
        "unetb200_destroy": (None, [vp]),
        "unetb200_last_error": (C.c_char_p, [vp]),
        "unetb200_check_device_error": (i32, [vp, P(i32)]),
        "unetb200_num_tensors": (i32, []),
        "unetb200_tensor_info": (i32, [i32, C.c_char_p, i32, P(i32), P(i32), P(i64), P(i32)]),
        "unetb200_num_params": (i64, []),
        "unetb200_num_buffers": (i64, []),
        "unetb200_num_counters": (i32, []),
        "unetb200_load_weights": (i32, [vp, vp, vp, vp]),
        "unetb200_load_weights_ex": (i32, [vp, vp, vp, i32, vp]),
        "unetb200_forward_infer": (i32, [vp, vp, vp, vp, vp, f32, i32, vp]),
        "unetb200_infer_host": (i32, [vp, vp, vp, vp, vp, f32, i32]),
        "unetb200_infer_host_submit": (i32, [vp, i32, vp, vp, vp, vp, f32, i32]),
        "unetb200_infer_host_wait": (i32, [vp, i32]),
        "unetb200_infer_host_u8": (i32, [vp, vp, i32, P(f32), P(f32), vp, vp, vp, f32, i32]),
        "unetb200_infer_host_u8_submit": (i32, [vp, i32, vp, i32, P(f32), P(f32), vp, vp, vp, f32, i32]),
        "unetb200_forward_infer_u8": (i32, [vp, vp, i32, P(f32), P(f32), vp, vp, vp, f32, i32, vp]),
        "unetb200_infer_launch_count": (i32, [vp, i32]),
        "unetb200_infer_debug_count": (i32, [vp, i32]),
        "unetb200_infer_debug_info": (i32, [vp, i32, i32, C.c_char_p, i32, P(i32)]),
        "unetb200_infer_debug_copy": (i32, [vp, i32, i32, vp, i64, vp]),
        "unetb200_profile_infer": (i32, [vp, vp, vp, i32, vp, P(f32), P(i32), i32, P(i32)]),
        "unetb200_profile_name": (i32, [vp, i32, i32, C.c_char_p, i32]),
        "unetb200_conv_nhwc": (i32, [vp, vp, vp, vp, vp, vp, i32, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
        "unetb200_train_forward": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, vp]),
        "unetb200_train_forward_u8": (i32, [vp, vp, i32, P(f32), P(f32), vp, vp, vp, vp, vp, i32, vp]),
        "unetb200_train_backward": (i32, [vp, vp, i32, i32, i32, vp]),
        "unetb200_grad_bucket_range": (i32, [i32, P(i64), P(i64)]),
        "unetb200_train_launch_count": (i32, [vp, i32, P(i32), P(i32)]),
        "unetb200_train_debug_count": (i32, [vp, i32]),
        "unetb200_train_debug_info": (i32, [vp, i32, i32, C.c_char_p, i32, P(i32), P(i32)]),
        "unetb200_train_debug_copy": (i32, [vp, i32, i32, vp, i64, vp]),
        "unetb200_profile_enable": (i32, [vp, i32]),
        "unetb200_profile_dump": (i32, [vp, C.c_char_p]),
        "unetb200_loss_scratch_floats": (i32, []),
        "unetb200_loss_bce_dice_forward": (i32, [vp, vp, i64, f32, vp, vp, vp]),
        "unetb200_loss_bce_dice_backward": (i32, [vp, vp, vp, vp, vp, f32, f32, vp, i64, vp]),
        "unetb200_seg_metrics_scratch_floats": (i32, [i32]),
        "unetb200_seg_metrics": (i32, [vp, vp, i32, i64, f32, f32, vp, vp, vp]),
        "unetb200_adamw_step": (i32, [vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i64, f32, i32, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return sigs


def exported_symbols():
    """Names include/unetb200.h declares (used by the CPU-side export test)."""
    hdr = (_HERE.parent / "include" / "unetb200.h").read_text()
    import re

    return sorted(set(re.findall(r"\b(unetb200_[a-z_0-9]+)\s*\(", hdr)))


def load(precision: str = "bf16"):
    """The C-ABI library for a storage precision: "bf16" (default; training and inference) or "fp16" (inference only).
    Both export the same symbols (built with -Bsymbolic) and are loaded RTLD_LOCAL, side by side if needed."""
    lib = _libs.get(precision)
    if lib is None:
        path = LIB_PATHS.get(precision)
        if path is None:
            raise ValueError(f"precision must be one of {sorted(LIB_PATHS)} (got {precision!r})")
        if not path.exists():
            raise UnetB200Error(
                f"{path} is missing: build it with `make` (nvcc, sm_100a). There is no CPU / cuDNN fallback."
            )
        lib = C.CDLL(str(path), mode=getattr(os, "RTLD_LOCAL", 0))
        _proto(lib)
        _libs[precision] = lib
    return lib


def grad_bucket_ranges():
    """[(begin, end)] element ranges of the flat gradient array, in backward-completion order (stage 0..3)."""
    lib = load()
    out = []
    for stage in range(4):
        b, e = C.c_longlong(), C.c_longlong()
        if lib.unetb200_grad_bucket_range(stage, C.byref(b), C.byref(e)):
            raise UnetB200Error("grad_bucket_range failed")
        out.append((b.value, e.value))
    return out


def check_global(rc: int, what: str):
    """Error check for the ctx-free entry points (loss): message is the thread's last global error."""
    if rc:
        raise UnetB200Error(f"{what}: " + load().unetb200_last_error(None).decode())


def tensor_table():
    """[(name, shape tuple, offset, kind)] in state_dict order; kind 0 param, 1 fp32 buffer, 2 int64 counter."""
    lib = load()
    out = []
    name = C.create_string_buffer(256)
    nd, kind, off = C.c_int(), C.c_int(), C.c_longlong()
    shape = (C.c_int * 4)()
    for i in range(lib.unetb200_num_tensors()):
        rc = lib.unetb200_tensor_info(i, name, 256, C.byref(nd), shape, C.byref(off), C.byref(kind))
        if rc:
            raise UnetB200Error("tensor_info failed")
        out.append((name.value.decode(), tuple(shape[j] for j in range(nd.value)), off.value, kind.value))
    return out


class Context:
    """Owns one unetb200_ctx (device, max_batch, H, W)."""

    def __init__(self, device: int, max_batch: int, H: int, W: int, precision: str = "bf16"):
        self.lib = load(precision)
        self.precision = precision
        self.handle = C.c_void_p()
        rc = self.lib.unetb200_create(C.byref(self.handle), device, max_batch, H, W)
        if rc:
            raise UnetB200Error("unetb200_create: " + self.lib.unetb200_last_error(None).decode())
        self.device, self.max_batch, self.H, self.W = device, max_batch, H, W

    def check(self, rc: int, what: str):
        if rc:
            raise UnetB200Error(f"{what}: " + self.lib.unetb200_last_error(self.handle).decode())

    def device_error_flag(self) -> int:
        f = C.c_int()
        self.check(self.lib.unetb200_check_device_error(self.handle, C.byref(f)), "check_device_error")
        return f.value

    def close(self):
        if self.handle:
            self.lib.unetb200_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
