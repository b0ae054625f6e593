"""Per-warp cycle breakdown of the tcgen05 jet kernel (diagnostic; counters cost a few % of kernel time)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import gpu_util as gu
from nsfnet_b200 import _capi
from oracle import jet_numpy as J
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
md, ed = J.NetDesc(2, 3, 6, 80), J.NetDesc(2, 1, 4, 40)
pm, pe = J.init_params(md, 1), J.init_params(ed, 2)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(n, device="cuda", generator=g); y = torch.rand(n, device="cuda", generator=g)
abi = gu.Abi((2, 3, 6, 80), (2, 1, 4, 40), path=2)
cp = _capi.physics(2000., alpha_evm=0.05, has_evm=True)
abi.step(pm, cp, x, y, params_evm=pe, want_resid=False)
abi.ctx.stage_cycles(read=False)
abi.step(pm, cp, x, y, params_evm=pe, want_resid=False)
c = abi.ctx.stage_cycles()
names = ["mma_wait", "work", "fence", "barrier", "steps"]
for w in range(15):
    r = c[w]
    if w == 3:
        print(f"warp {w:2d} issuer : " + "  ".join(f"{nm}={v / max(r[4], 1):8.1f}" for nm, v in zip(["w_wait", "issue", "mma_wait", "barrier"], r[:4])) + f"  steps={r[4]:.0f}")
    elif (w & 3) != 3:
        print(f"warp {w:2d} q{w & 3} sub{w >> 2}: fwd " + "  ".join(f"{nm}={v / max(r[4], 1):7.1f}" for nm, v in zip(names[:4], r[:4])) +
              "  | rev " + "  ".join(f"{nm}={v / max(r[9], 1):7.1f}" for nm, v in zip(names[:4], r[5:9])) + f"  steps={r[4]:.0f}+{r[9]:.0f}")
pairs = c[3][4] / 33.0   # stage-slots per tile group: 11 MMA stages x 3 slots
print("issuer operand wait per tile group by stage s (all slots):", "  ".join(f"s{i}={c[7][i] / max(pairs, 1):.0f}" for i in range(1, 12)))
