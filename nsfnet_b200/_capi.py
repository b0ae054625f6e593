"""ctypes binding of libnsf_b200.so (include/nsf_b200.h).

The library is the product: there is no Python / PyTorch / CPU fallback.  ``load()`` raises if the
shared object has not been built (``python -c "import __graft_entry__ as g; g.build()"``) and
``Context`` raises if the device is not an sm_100 GPU (``NSF_E_ARCH``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnsf_b200.so")

NSF_OK = 0
NSF_E_ARG, NSF_E_ARCH, NSF_E_CUDA, NSF_E_SHAPE = -1, -2, -3, -4
NSF_HAS_EVM = 1
NSF_EVM_TRAINABLE = 2
NSF_VTM_FROM_E = 4
NSF_MAX_BLOCKS = 2
NSF_LOSS_SLOTS = 16

EXPORTS = ["nsf_abi_version", "nsf_last_error", "nsf_create", "nsf_destroy", "nsf_set_path", "nsf_get_info",
           "nsf_set_timing", "nsf_last_kernel_ms", "nsf_get_stage_cycles",
           "nsf_step", "nsf_residuals", "nsf_forward", "nsf_adam", "nsf_selftest_umma",
           "nsf_adam_dev", "nsf_adam_tick", "nsf_lhs_points", "nsf_wall_distance", "nsf_sdf_weights", "nsf_error_norms"]


class NsfNetDesc(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("n_out", C.c_int32), ("n_hidden_layers", C.c_int32), ("hidden", C.c_int32)]


class NsfPhysics(C.Structure):
    _fields_ = [("inv_Re", C.c_float), ("vis_t0", C.c_float), ("alpha_evm", C.c_float), ("alpha_e", C.c_float),
                ("coord_scale", C.c_float), ("eq4_weight", C.c_float), ("flags", C.c_uint32), ("alpha_evm_init", C.c_float),
                ("n_f_norm", C.c_double)]


class NsfDataBlock(C.Structure):
    _fields_ = [("x", C.c_void_p), ("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("p", C.c_void_p),
                ("n", C.c_int64), ("cu", C.c_float), ("cv", C.c_float), ("cp", C.c_float), ("reserved", C.c_int32)]


class NsfAdamDev(C.Structure):
    """Device-resident Adam state (32 bytes); host mirror used to initialise the device copy."""
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("grad_scale", C.c_float), ("step", C.c_int32), ("reserved", C.c_int32 * 2)]


class NsfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libnsf_b200 error {code}: {msg}")
        self.code = code


def bind(lib: C.CDLL) -> C.CDLL:
    """Declare the prototypes of include/nsf_b200.h on an already opened library."""
    vp, f, i32, i64 = C.c_void_p, C.c_float, C.c_int32, C.c_int64
    lib.nsf_abi_version.restype = C.c_int
    lib.nsf_abi_version.argtypes = []
    lib.nsf_last_error.restype = C.c_char_p
    lib.nsf_last_error.argtypes = []
    lib.nsf_create.restype = C.c_int
    lib.nsf_create.argtypes = [C.c_int, C.POINTER(NsfNetDesc), C.POINTER(NsfNetDesc), C.POINTER(vp)]
    lib.nsf_destroy.restype = C.c_int
    lib.nsf_destroy.argtypes = [vp]
    lib.nsf_set_path.restype = C.c_int
    lib.nsf_set_path.argtypes = [vp, C.c_int]
    lib.nsf_get_info.restype = C.c_int
    lib.nsf_get_info.argtypes = [vp, C.POINTER(i64 * 4)]
    lib.nsf_set_timing.restype = C.c_int
    lib.nsf_set_timing.argtypes = [vp, C.c_int]
    lib.nsf_last_kernel_ms.restype = C.c_int
    lib.nsf_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.nsf_get_stage_cycles.restype = C.c_int
    lib.nsf_get_stage_cycles.argtypes = [vp, C.POINTER(C.c_double)]
    lib.nsf_step.restype = C.c_int
    lib.nsf_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i64, C.POINTER(NsfDataBlock), i32, C.POINTER(NsfPhysics),
                             vp, vp, vp, vp, vp, vp, vp]
    lib.nsf_residuals.restype = C.c_int
    lib.nsf_residuals.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, C.POINTER(NsfPhysics), vp, vp, vp, vp]
    lib.nsf_forward.restype = C.c_int
    lib.nsf_forward.argtypes = [vp, i32, vp, vp, vp, i64, vp, vp]
    lib.nsf_adam.restype = C.c_int
    lib.nsf_adam.argtypes = [vp, vp, vp, vp, i64, f, f, f, f, i64, f, vp]
    lib.nsf_adam_dev.restype = C.c_int
    lib.nsf_adam_dev.argtypes = [vp, vp, vp, vp, i64, vp, vp]
    lib.nsf_adam_tick.restype = C.c_int
    lib.nsf_adam_tick.argtypes = [vp, vp]
    lib.nsf_lhs_points.restype = C.c_int
    lib.nsf_lhs_points.argtypes = [i64, i64, i64, C.c_uint32, f, f, f, f, vp, vp, vp]
    lib.nsf_wall_distance.restype = C.c_int
    lib.nsf_wall_distance.argtypes = [vp, vp, i64, vp, vp, i32, vp, vp]
    lib.nsf_sdf_weights.restype = C.c_int
    lib.nsf_sdf_weights.argtypes = [vp, vp, i64, vp, vp, i32, f, f, vp, vp, vp]
    lib.nsf_error_norms.restype = C.c_int
    lib.nsf_error_norms.argtypes = [vp, vp, vp, vp, i64, vp, vp]
    lib.nsf_selftest_umma.restype = C.c_int
    lib.nsf_selftest_umma.argtypes = [C.c_int, i32, vp, vp, vp, i32, i32, vp]
    return lib


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Open the CUDA library.  Fails loudly when it is missing -- there is nothing to fall back to."""
    global _lib
    if _lib is None:
        path = os.environ.get("NSF_B200_LIB", LIB_PATH)      # kernel experiments: another build of the same library
        if not os.path.exists(path):
            raise ImportError(f"{path} is not built; run __graft_entry__.build() (nvcc, sm_100a). "
                              "nsfnet_b200 has no CPU or PyTorch fallback for the hot path.")
        _lib = bind(C.CDLL(path))
        if _lib.nsf_abi_version() != 1:
            raise ImportError("libnsf_b200.so ABI version mismatch")
    return _lib


def check(lib: C.CDLL, rc: int) -> None:
    if rc != NSF_OK:
        raise NsfError(rc, lib.nsf_last_error().decode("utf-8", "replace"))


def physics(Re: float, *, vis_t0: Optional[float] = None, alpha_evm: float = 0.0, alpha_e: float = 1.0,
            coord_scale: float = 1.0, eq4_weight: float = 0.1, has_evm: bool = False, evm_trainable: bool = False,
            n_f_norm: float = 0.0, vtm_from_e_alpha: Optional[float] = None) -> NsfPhysics:
    flags = (NSF_HAS_EVM if has_evm else 0) | (NSF_EVM_TRAINABLE if (has_evm and evm_trainable) else 0)
    if has_evm and vtm_from_e_alpha is not None:
        flags |= NSF_VTM_FROM_E
    return NsfPhysics(1.0 / Re, (20.0 / Re) if vis_t0 is None else vis_t0, alpha_evm, alpha_e, coord_scale, eq4_weight,
                      flags, 0.0 if vtm_from_e_alpha is None else float(vtm_from_e_alpha), float(n_f_norm))


class Context:
    """Owner of one NsfCtx.  Arguments of the methods are raw device addresses (``tensor.data_ptr()``)."""

    def __init__(self, lib: C.CDLL, device: int, main: Sequence[int], evm: Optional[Sequence[int]] = None):
        self.lib = lib
        self.h = C.c_void_p()
        md = NsfNetDesc(*main)
        ed = NsfNetDesc(*evm) if evm is not None else None
        check(lib, lib.nsf_create(device, C.byref(md), C.byref(ed) if ed is not None else None, C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.nsf_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_path(self, path: int):
        check(self.lib, self.lib.nsf_set_path(self.h, path))

    def info(self):
        arr = (C.c_int64 * 4)()
        check(self.lib, self.lib.nsf_get_info(self.h, C.byref(arr)))
        return dict(sms=arr[0], path=arr[1], launches=arr[2], workspace_bytes=arr[3])

    def set_timing(self, enable: bool):
        check(self.lib, self.lib.nsf_set_timing(self.h, 1 if enable else 0))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        check(self.lib, self.lib.nsf_last_kernel_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def stage_cycles(self, read: bool = True):
        if not read:
            check(self.lib, self.lib.nsf_get_stage_cycles(self.h, None))
            return None
        arr = (C.c_double * 256)()
        check(self.lib, self.lib.nsf_get_stage_cycles(self.h, arr))
        return [list(arr[w * 16:(w + 1) * 16]) for w in range(16)]

    def step(self, params_main, params_evm, x, y, w, vtm_in, vtm_out, n_f, blocks, phys, grad_main, grad_evm,
             loss_parts, residuals_out=None, e_out=None, vis_t_out=None, stream=None):
        nb = len(blocks)
        arr = (NsfDataBlock * max(nb, 1))(*blocks)
        check(self.lib, self.lib.nsf_step(self.h, params_main, params_evm, x, y, w, vtm_in, vtm_out, n_f, arr, nb,
                                          C.byref(phys), grad_main, grad_evm, loss_parts, residuals_out, e_out, vis_t_out,
                                          stream))

    def residuals(self, params_main, params_evm, x, y, vtm_in, vtm_out, n, phys, residuals_out, e_out=None,
                  vis_t_out=None, stream=None):
        check(self.lib, self.lib.nsf_residuals(self.h, params_main, params_evm, x, y, vtm_in, vtm_out, n, C.byref(phys),
                                               residuals_out, e_out, vis_t_out, stream))

    def forward(self, which, params, x, y, n, out, stream=None):
        check(self.lib, self.lib.nsf_forward(self.h, which, params, x, y, n, out, stream))
