"""Host SIMT emulation of the FFMA kernels -- TEST INFRASTRUCTURE ONLY.

Builds nsfnet_b200/csrc/{nsf_ffma,nsf_capi}.cu with g++ and -DNSF_EMU (CTAs and threads become host
loops, "device" memory is malloc) into tests/emu/_build/libnsf_emu.so and drives it through the very
same ctypes prototypes as the CUDA library.  Purpose: check the kernels' index logic, the C-ABI
orchestration and the host-side layer in a container without a GPU.  nsfnet_b200 never loads it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from nsfnet_b200 import _capi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build", "libnsf_emu.so")
SRCS = [os.path.join(ROOT, "nsfnet_b200", "csrc", f) for f in ("nsf_ffma.cu", "nsf_capi.cu", "nsf_aux.cu")]
DEPS = SRCS + [os.path.join(ROOT, "nsfnet_b200", "csrc", f) for f in ("nsf_ffma_body.h", "nsf_geom.h", "nsf_internal.h")] + \
    [os.path.join(ROOT, "include", "nsf_b200.h")]


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-DNSF_EMU", "-shared", "-fPIC", "-ffp-contract=off",
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "nsfnet_b200", "csrc"), "-x", "c++"] + SRCS + ["-o", OUT]
    subprocess.run(cmd, check=True)
    return OUT


def load():
    return _capi.bind(C.CDLL(build()))


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, np.float32).reshape(-1))


def run_step(lib, main_desc, params_main, phys, x, y, blocks=(), evm_desc=None, params_evm=None, w=None, vtm_in=None,
             want_resid=True):
    """One nsf_step on host arrays.  blocks: [(x,y,u,v,p_or_None,cu,cv,cp)].  Returns a dict of numpy outputs."""
    ctx = _capi.Context(lib, 0, main_desc, evm_desc)
    x, y, w, vtm_in = f32(x), f32(y), f32(w), f32(vtm_in)
    pm, pe = f32(params_main), f32(params_evm)
    n = x.size
    keep = []
    blks = []
    for (bx, by, bu, bv, bp, cu, cv, cp) in blocks:
        arrs = [f32(bx), f32(by), f32(bu), f32(bv), f32(bp)]
        keep.append(arrs)
        blks.append(_capi.NsfDataBlock(ptr(arrs[0]), ptr(arrs[1]), ptr(arrs[2]), ptr(arrs[3]), ptr(arrs[4]), arrs[0].size, cu, cv, cp, 0))
    gm = np.full(pm.size, np.nan, np.float32)
    ge = np.full(pe.size, np.nan, np.float32) if pe is not None else None
    lp = np.full(16, np.nan, np.float32)
    res = np.full(4 * n, np.nan, np.float32) if want_resid else None
    e = np.full(n, np.nan, np.float32) if pe is not None else None
    vis = np.full(n, np.nan, np.float32)
    vtm_out = np.full(n, np.nan, np.float32) if pe is not None else None
    ctx.step(ptr(pm), ptr(pe), ptr(x), ptr(y), ptr(w), ptr(vtm_in), ptr(vtm_out), n, blks, phys, ptr(gm), ptr(ge), ptr(lp),
             ptr(res), ptr(e), ptr(vis))
    info = ctx.info()
    ctx.close()
    return dict(grad_main=gm, grad_evm=ge, loss_parts=lp, resid=None if res is None else res.reshape(4, n), e=e, vis_t=vis,
                vtm_out=vtm_out, info=info)
