"""Per-warp cycle breakdown of the layer-major tcgen05 kernel (diagnostic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import gpu_util as gu
from nsfnet_b200 import _capi
from oracle import jet_numpy as J
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 4
md, ed = J.NetDesc(2, 3, 6, 80), J.NetDesc(2, 1, 4, 40)
pm, pe = J.init_params(md, 1), J.init_params(ed, 2)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(n, device="cuda", generator=g); y = torch.rand(n, device="cuda", generator=g)
abi = gu.Abi((2, 3, 6, 80), (2, 1, 4, 40), path=3)
abi.ctx.set_tiles_per_batch(nt)
cp = _capi.physics(2000., alpha_evm=0.05, has_evm=True)
abi.step(pm, cp, x, y, params_evm=pe, want_resid=False)
abi.ctx.stage_cycles(read=False)
abi.step(pm, cp, x, y, params_evm=pe, want_resid=False)
c = abi.ctx.stage_cycles()
for w in (0, 1, 2, 3, 6, 13):
    r = c[w]
    if w == 3:
        print(f"issuer: per item ready_wait={r[0]/max(r[3],1):.0f} issue={r[1]/max(r[3],1):.0f} dw_wait={r[2]/max(r[3],1):.0f} items={r[3]:.0f}")
    else:
        it = max(r[4], 1)
        print(f"warp {w:2d}: per item pre={r[0]/it:.0f} fence+arrive={r[1]/it:.0f} post={r[2]/it:.0f} (of which MMA wait {r[5]/it:.0f}) stage_end={r[3]/it:.0f} items={r[4]:.0f}")
