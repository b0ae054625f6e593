"""Helpers of the `-m gpu` tests: drive libnsf_b200.so through its C ABI with torch CUDA tensors."""
import numpy as np
import torch

from nsfnet_b200 import _capi


def rel(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def dev(a):
    return None if a is None else torch.as_tensor(np.ascontiguousarray(np.asarray(a, np.float32).reshape(-1))).cuda()


def ptr(t):
    return None if t is None else t.data_ptr()


class Abi:
    """One NsfCtx + device buffers; `step` mirrors tests/emu/emu.py:run_step on the real library."""

    def __init__(self, main_desc, evm_desc=None, path=0):
        self.lib = _capi.load()
        self.ctx = _capi.Context(self.lib, torch.cuda.current_device(), main_desc, evm_desc)
        if path:
            self.ctx.set_path(path)

    def step(self, params_main, phys, x, y, blocks=(), params_evm=None, w=None, vtm_in=None, want_resid=True):
        pm, pe = dev(params_main), dev(params_evm)
        x, y, w, vtm_in = (a if isinstance(a, torch.Tensor) else dev(a) for a in (x, y, w, vtm_in))
        n = x.numel()
        keep, blks = [], []
        for (bx, by, bu, bv, bp, cu, cv, cp) in blocks:
            arrs = [dev(bx), dev(by), dev(bu), dev(bv), dev(bp)]
            keep.append(arrs)
            blks.append(_capi.NsfDataBlock(ptr(arrs[0]), ptr(arrs[1]), ptr(arrs[2]), ptr(arrs[3]), ptr(arrs[4]), arrs[0].numel(), cu, cv, cp, 0))
        nan = float("nan")
        gm = torch.full((pm.numel(),), nan, device="cuda")
        ge = torch.full((pe.numel(),), nan, device="cuda") if pe is not None else None
        lp = torch.full((16,), nan, device="cuda")
        res = torch.full((4 * n,), nan, device="cuda") if want_resid else None
        e = torch.full((n,), nan, device="cuda") if pe is not None else None
        vis = torch.full((n,), nan, device="cuda")
        vtm_out = torch.full((n,), nan, device="cuda") if pe is not None else None
        self.ctx.step(ptr(pm), ptr(pe), ptr(x), ptr(y), ptr(w), ptr(vtm_in), ptr(vtm_out), n, blks, phys, ptr(gm), ptr(ge), ptr(lp),
                      ptr(res), ptr(e), ptr(vis), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        c = lambda t: None if t is None else t.cpu().numpy()
        return dict(grad_main=c(gm), grad_evm=c(ge), loss_parts=c(lp), resid=None if res is None else c(res).reshape(4, n),
                    e=c(e), vis_t=c(vis), vtm_out=c(vtm_out), info=self.ctx.info())
