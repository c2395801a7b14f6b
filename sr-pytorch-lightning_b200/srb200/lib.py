"""ctypes binding of libsrb200.so (C ABI: include/srb200.h).

There is NO CPU fallback: if the shared library is missing or a kernel reports an error the
product path raises (`RuntimeError`, the exception class the reference's training loop catches,
/root/reference/train.py:237-253).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SRB200_LIB", os.path.join(os.path.dirname(_HERE), "csrc", "libsrb200.so"))

F32, BF16 = 0, 1
RELU, RESIDUAL, MASK, COLSUM, OUT2 = 1, 2, 4, 8, 16
PACK_SIMT, PACK_UMMA = 0, 1
PACK_FWD, PACK_DGRAD = 0, 1
BACKEND_AUTO, BACKEND_SIMT, BACKEND_UMMA = 0, 1, 2

c_i32, c_i64, c_f32, c_vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class ConvDesc(C.Structure):
    _fields_ = [(n, c_i32) for n in ("N", "H", "W", "Cin", "Cout", "ksize", "dtype", "flags")] + \
               [("scale", c_f32)] + \
               [(n, c_i32) for n in ("shuffle", "colsum_groups", "backend", "x_cs", "x_co", "y_cs", "y_co",
                                     "r_cs", "r_co", "m_cs", "m_co", "y2_cs", "y2_co")]


class WgradDesc(C.Structure):
    _fields_ = [(n, c_i32) for n in ("N", "H", "W", "Cin", "Cout", "ksize", "dtype", "accumulate", "shuffle",
                                     "backend", "x_cs", "x_co", "g_cs", "g_co")] + [("alpha", c_f32)]


class PatchItem(C.Structure):
    _fields_ = [("lr_img", c_vp), ("hr_img", c_vp)] + [(n, c_i32) for n in ("lr_h", "lr_w", "hr_h", "hr_w", "lr_top", "lr_left",
                                                                           "angle", "hflip", "vflip", "reserved")]


class PackItem(C.Structure):
    _fields_ = [("src", c_vp), ("dst", c_vp)] + [(n, c_i32) for n in ("Cout", "Cin", "ksize", "packing", "mode", "shuffle")]


CHAIN_CONV, CHAIN_CA_BWD = 0, 1
CHAIN_CA = 32
CHAIN_CA_BWD_FUSED = 64
CHAIN_Y_SCRATCH = 128
CHAIN_NONE = 0xFFFF
CHAIN_MAX_OPS = 64


class ChainOp(C.Structure):
    _fields_ = [("kind", c_i32), ("flags", C.c_uint32), ("x", C.c_uint16), ("y", C.c_uint16), ("e", C.c_uint16),
                ("y2", C.c_uint16), ("e2", C.c_uint16), ("reserved0", C.c_uint16), ("w_layer", c_i32), ("scale", c_f32),
                ("colsum_groups", c_i32), ("ca_cr", c_i32), ("colsum_scale", c_f32),
                ("bias", c_vp), ("colsum", c_vp), ("colsum2", c_vp), ("ca_w1", c_vp), ("ca_b1", c_vp), ("ca_w2", c_vp), ("ca_b2", c_vp),
                ("ca_s", c_vp), ("ca_y", c_vp), ("ca_dw1", c_vp), ("ca_db1", c_vp), ("ca_dw2", c_vp), ("ca_db2", c_vp),
                ("ca_scratch", c_vp)]


class ChainDesc(C.Structure):
    _fields_ = [("N", c_i32), ("H", c_i32), ("W", c_i32), ("n_ops", c_i32), ("ops", C.POINTER(ChainOp)),
                ("space_base", c_vp * 4), ("space_slots", c_i32 * 4), ("weights", c_vp), ("n_layers", c_i32),
                ("counters", c_vp), ("trace", c_vp), ("tile_flags", c_vp), ("kernel_hint", c_i32)]


class HaloDesc(C.Structure):
    _fields_ = [("src_top", c_vp), ("src_bot", c_vp), ("dst_up", c_vp), ("dst_dn", c_vp), ("slab_bytes", c_i64),
                ("flag_up", c_vp), ("flag_dn", c_vp), ("wait_up", c_vp), ("wait_dn", c_vp), ("frame", c_vp), ("done", c_vp)]


class WgradItem(C.Structure):
    _fields_ = [("d", WgradDesc), ("x", c_vp), ("gy", c_vp), ("dw", c_vp), ("dbias", c_vp)]


# name -> (restype, argtypes); every symbol include/srb200.h declares
PROTOTYPES = {
    "srb_abi_version": (c_i32, []),
    "srb_last_error": (C.c_char_p, []),
    "srb_create": (c_i32, [c_i32, C.POINTER(c_vp)]),
    "srb_destroy": (c_i32, [c_vp]),
    "srb_num_sms": (c_i32, [c_vp]),
    "srb_launch_count": (C.c_ulonglong, []),
    "srb_packed_weight_bytes": (C.c_size_t, [c_i32, c_i32, c_i32, c_i32, c_i32]),
    "srb_pack_weight": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "srb_pack_bias": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_vp, c_vp]),
    "srb_conv": (c_i32, [c_vp, C.POINTER(ConvDesc)] + [c_vp] * 9),
    "srb_conv_wgrad": (c_i32, [c_vp, C.POINTER(WgradDesc)] + [c_vp] * 5),
    "srb_conv_wgrad_batched": (c_i32, [c_vp, C.POINTER(WgradItem), c_i32, c_vp]),
    "srb_set_wgrad_sm_budget": (c_i32, [c_vp, c_i32]),
    "srb_delay": (c_i32, [c_vp, c_i64, c_vp]),
    "srb_patch_batch": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "srb_bn_stats": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i64, c_i32, c_f32, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "srb_bn_act_fwd": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32,
                               c_vp, c_i32, c_i32, c_vp]),
    "srb_bn_act_bwd": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_vp, c_i32, c_i32, c_i32, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                               c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp]),
    "srb_conv_uses_umma": (c_i32, [C.POINTER(ConvDesc)]),
    "srb_wgrad_uses_umma": (c_i32, [C.POINTER(WgradDesc)]),
    "srb_wgrad_plan": (c_i32, [c_i32, c_i32] + [C.POINTER(c_i32)] * 5),
    "srb_conv_chain": (c_i32, [c_vp, C.POINTER(ChainDesc), c_vp]),
    "srb_conv_chain_grid": (c_i32, [c_vp, c_i32, c_i32, c_i32]),
    "srb_conv_chain_uses_cluster": (c_i32, [C.POINTER(ChainDesc)]),
    "srb_ca_fwd": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_i32] + [c_vp] * 8),
    "srb_ca_bwd": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32] + [c_vp] * 15 + [c_i32, c_i32, c_vp]),
    "srb_pack_table": (c_i32, [c_vp, c_vp, c_i32, c_i64, c_vp]),
    "srb_nchw_to_nhwc": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_vp, c_i32, c_i32, c_vp]),
    "srb_nhwc_to_nchw": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "srb_copy_channels": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_vp, c_i32, c_i32, c_i32, c_i64, c_i32, c_vp]),
    "srb_add_channels": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_vp, c_i32, c_i32, c_vp, c_i32, c_i32, c_i32, c_i64, c_i32, c_vp]),
    "srb_relu_bwd": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_vp, c_i32, c_i32, c_vp, c_i32, c_i32, c_i32, c_i64, c_i32, c_vp]),
    "srb_pixel_unshuffle": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "srb_colsum": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i64, c_i32, c_vp, c_i32, c_vp]),
    "srb_l1_loss": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "srb_adam_step": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_f32, c_f32, c_f32, c_i32, c_vp, c_f32, c_vp]),
    "srb_inc_counter": (c_i32, [c_vp, c_vp, c_vp]),
    "srb_inc_counter64": (c_i32, [c_vp, c_vp, c_vp]),
    "srb_halo_exchange": (c_i32, [c_vp, C.POINTER(HaloDesc), c_vp]),
    "srb_ipc_alloc": (c_i32, [c_vp, C.c_size_t, C.POINTER(c_vp), C.c_char_p]),
    "srb_ipc_open": (c_i32, [c_vp, C.c_char_p, C.POINTER(c_vp)]),
    "srb_ipc_close": (c_i32, [c_vp, c_vp]),
    "srb_ipc_free": (c_i32, [c_vp, c_vp]),
    "srb_debug_set_trace": (c_i32, [c_vp, c_vp]),
}

_lib = None
_lock = threading.Lock()
_ctx: dict[int, int] = {}


def load():
    """dlopen libsrb200.so and bind every prototype; raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"libsrb200.so not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C sr-pytorch-lightning_b200/csrc`). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        if lib.srb_abi_version() != 1:
            raise RuntimeError(f"libsrb200.so ABI version {lib.srb_abi_version()} != 1")
        _lib = lib
    return _lib


def last_error() -> str:
    return load().srb_last_error().decode(errors="replace")


def check(rc: int, what: str = ""):
    if rc != 0:
        raise RuntimeError(f"libsrb200 {what} failed (code {rc}): {last_error()}")


def launch_count() -> int:
    return int(load().srb_launch_count())


def ctx(device_index: int) -> int:
    """Per-device context handle (created on first use)."""
    h = _ctx.get(device_index)
    if h is None:
        lib = load()
        out = c_vp()
        check(lib.srb_create(int(device_index), C.byref(out)), "srb_create")
        h = out.value
        _ctx[device_index] = h
    return h
