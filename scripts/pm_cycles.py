"""Per-role cycle counters of the points-on-M tcgen05 kernel (nsf_get_stage_cycles) for one workload."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from nsfnet_b200 import _capi
from oracle import jet_numpy as J
from tests import gpu_util as gu

wl = sys.argv[1] if len(sys.argv) > 1 else "ev"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
L, H = (6, 80) if wl == "ev" else (4, 120)
has_evm = wl == "ev"
abi = gu.Abi((2, 3, L, H), (2, 1, 4, 40) if has_evm else None, path=3)
pm, pe = J.init_params(J.NetDesc(2, 3, L, H), 1), J.init_params(J.NetDesc(2, 1, 4, 40), 2)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(n, device="cuda", generator=g); y = torch.rand(n, device="cuda", generator=g)
cp = _capi.physics(2000., alpha_evm=0.05, has_evm=has_evm)
abi.ctx.stage_cycles(False)
for _ in range(2):
    abi.step(pm, cp, x, y, blocks=[], params_evm=pe if has_evm else None, want_resid=False)
c = [v for row in abi.ctx.stage_cycles(True) for v in row]
tiles = c[3] / (2 * L - 1)
print(f"{wl} n={n}: tiles/CTA {tiles:.0f}, MMA stages {c[3]:.0f}")
print(f"issuer  per stage: wait-operands {c[0]/c[3]:.0f}  issue+weights {c[1]/c[3]:.0f}  (weight wait {c[2]/c[3]:.0f})   per tile total {(c[0]+c[1])/tiles:.0f}")
print(f"epi w0 per tile: stage0 {c[12]/tiles:.0f}  fwd hidden {c[13]/tiles:.0f}  output+bwd {c[14]/tiles:.0f}  rev phase A {c[15]/tiles:.0f}  rev phase B {c[16]/tiles:.0f}  last {c[17]/tiles:.0f}   wait-D fwd {c[18]/tiles:.0f}  wait-D rev {c[19]/tiles:.0f}")
print(f"epi wN per tile: stage0 {c[20]/tiles:.0f}  fwd hidden {c[21]/tiles:.0f}  output+bwd {c[22]/tiles:.0f}  rev phase A {c[23]/tiles:.0f}  rev phase B {c[24]/tiles:.0f}  last {c[25]/tiles:.0f}   wait-D fwd {c[26]/tiles:.0f}  wait-D rev {c[27]/tiles:.0f}")
for nm, o in (("epi w0 ", 4), ("epi wN ", 8)):
    print(f"{nm} per tile: wait-D {c[o]/tiles:.0f}  wait-wgrad {c[o+1]/tiles:.0f}  work {c[o+2]/tiles:.0f}")
