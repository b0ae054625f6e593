"""Where the time of one nsf_step goes at a given collocation count: the whole call, the same without the boundary block, and the
jet kernel alone (CUDA events; L2 warm).  python scripts/step_breakdown.py [n ...]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from nsfnet_b200 import _capi
from nsfnet_b200.cavity_data import cavity_boundary
from oracle import jet_numpy as J
from tests import gpu_util as gu

ns = [int(v) for v in sys.argv[1:]] or [120_000, 1_000_000]
L, H = 6, 80
abi = gu.Abi((2, 3, L, H), (2, 1, 4, 40), path=3)
pm = gu.dev(J.init_params(J.NetDesc(2, 3, L, H), 1)); pe = gu.dev(J.init_params(J.NetDesc(2, 1, 4, 40), 2))
cp = _capi.physics(2000., alpha_evm=0.05, has_evm=True)
b = cavity_boundary(513)
bx, by, bu, bv = (gu.dev(np.asarray(b[k], np.float32)) for k in range(4))
blk = [_capi.NsfDataBlock(gu.ptr(bx), gu.ptr(by), gu.ptr(bu), gu.ptr(bv), None, bx.numel(), 10.0, 10.0, 0.0, 0)]
st = torch.cuda.current_stream().cuda_stream
for n in ns:
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(n, device="cuda", generator=g); y = torch.rand(n, device="cuda", generator=g)
    gm = torch.empty(pm.numel(), device="cuda"); lp = torch.empty(16, device="cuda")
    e = torch.empty(n, device="cuda"); vis = torch.empty(n, device="cuda"); vtm = torch.empty(n, device="cuda")

    def run(blocks, reps=30, n=n):
        call = lambda: abi.ctx.step(gu.ptr(pm), gu.ptr(pe), gu.ptr(x), gu.ptr(y), None, None, gu.ptr(vtm), n, blocks, cp, gu.ptr(gm), None, gu.ptr(lp),
                                    None, gu.ptr(e), gu.ptr(vis), st)
        for _ in range(5):
            call()
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            call()
        t1.record(); torch.cuda.synchronize()
        return t0.elapsed_time(t1) / reps

    abi.ctx.set_timing(False)
    full, nob = run(blk), run([])
    abi.ctx.set_timing(True)
    run(blk, 3)
    k = abi.ctx.last_kernel_ms()
    abi.ctx.set_timing(False)
    alone = run(blk, n=0)
    print(f"n={n}: step {full*1e3:.0f} us, without the boundary block {nob*1e3:.0f} us, jet kernel {k*1e3:.0f} us; the boundary block alone (pack + block + finalize) {alone*1e3:.0f} us", flush=True)
