"""A/B of the boundary block's join position (before the jet kernel / before the gradient-row reduction), alternating in one process."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from nsfnet_b200 import _capi
from nsfnet_b200.cavity_data import cavity_boundary
from oracle import jet_numpy as J
from tests import gpu_util as gu

L, H = 6, 80
abi = gu.Abi((2, 3, L, H), (2, 1, 4, 40), path=3)
pm = gu.dev(J.init_params(J.NetDesc(2, 3, L, H), 1)); pe = gu.dev(J.init_params(J.NetDesc(2, 1, 4, 40), 2))
cp = _capi.physics(2000., alpha_evm=0.05, has_evm=True)
b = cavity_boundary(513)
bx, by, bu, bv = (gu.dev(np.asarray(b[k], np.float32)) for k in range(4))
blk = [_capi.NsfDataBlock(gu.ptr(bx), gu.ptr(by), gu.ptr(bu), gu.ptr(bv), None, bx.numel(), 10.0, 10.0, 0.0, 0)]
st = torch.cuda.current_stream().cuda_stream
for n in [int(v) for v in sys.argv[1:]] or [1_000_000]:
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(n, device="cuda", generator=g); y = torch.rand(n, device="cuda", generator=g)
    gm = torch.empty(pm.numel(), device="cuda"); lp = torch.empty(16, device="cuda")
    e = torch.empty(n, device="cuda"); vis = torch.empty(n, device="cuda"); vtm = torch.empty(n, device="cuda")
    call = lambda: abi.ctx.step(gu.ptr(pm), gu.ptr(pe), gu.ptr(x), gu.ptr(y), None, None, gu.ptr(vtm), n, blk, cp, gu.ptr(gm), None, gu.ptr(lp),
                                None, gu.ptr(e), gu.ptr(vis), st)
    res = {"0": [], "1": []}
    for rnd in range(6):
        for mode in ("0", "1"):
            os.environ["NSF_LATE_JOIN"] = mode
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(20):
                call()
            t1.record(); torch.cuda.synchronize()
            res[mode].append(t0.elapsed_time(t1) / 20 * 1e3)
    print(f"n={n}: join before the jet kernel {np.median(res['0']):.0f} us (min {min(res['0']):.0f}), late join {np.median(res['1']):.0f} us (min {min(res['1']):.0f})", flush=True)
