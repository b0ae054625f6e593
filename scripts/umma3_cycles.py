"""Per-warp cycle breakdown of the tcgen05 jet kernel with the weights in tensor memory (path 4; diagnostic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import gpu_util as gu
from nsfnet_b200 import _capi
from oracle import jet_numpy as J
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
md, ed = J.NetDesc(2, 3, 6, 80), J.NetDesc(2, 1, 4, 40)
pm, pe = J.init_params(md, 1), J.init_params(ed, 2)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(n, device="cuda", generator=g); y = torch.rand(n, device="cuda", generator=g)
abi = gu.Abi((2, 3, 6, 80), (2, 1, 4, 40), path=4)
cp = _capi.physics(2000., alpha_evm=0.05, has_evm=True)
abi.step(pm, cp, x, y, params_evm=pe, want_resid=False)
abi.ctx.stage_cycles(read=False)
abi.step(pm, cp, x, y, params_evm=pe, want_resid=False)
c = abi.ctx.stage_cycles()
for w in range(15):
    r = c[w]
    if w == 3:
        print(f"warp {w:2d} issuer : per stage-slot: weights_wait={r[0] / max(r[4], 1):8.1f} issue={r[1] / max(r[4], 1):8.1f} operand_wait={r[2] / max(r[4], 1):8.1f} stage-slots={r[4]:.0f}")
    elif (w & 3) != 3:
        f, v = max(r[4], 1), max(r[9], 1)
        print(f"warp {w:2d} q{w & 3} sub{w >> 2}: fwd mma_wait={r[0] / f:7.1f} work={r[1] / f:7.1f} fence={r[2] / f:6.1f} wload={r[3] / f:6.1f} | "
              f"rev mma_wait={r[5] / v:7.1f} work={r[6] / v:7.1f} fence={r[7] / v:6.1f} flush={r[8] / v:6.1f}  steps={r[4]:.0f}+{r[9]:.0f}")
