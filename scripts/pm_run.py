"""A few collocation steps of one workload on the tcgen05 points-on-M kernel (the command line profiled with ncu)."""
import sys
import torch
sys.path.insert(0, ".")
from nsfnet_b200 import _capi
from oracle import jet_numpy as J
from tests import gpu_util as gu

wl = sys.argv[1] if len(sys.argv) > 1 else "ev"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
path = int(sys.argv[4]) if len(sys.argv) > 4 else 3
L, H = (6, 80) if wl == "ev" else (4, 120)
has_evm = wl == "ev"
abi = gu.Abi((2, 3, L, H), (2, 1, 4, 40) if has_evm else None, path=path)
pm, pe = J.init_params(J.NetDesc(2, 3, L, H), 1), J.init_params(J.NetDesc(2, 1, 4, 40), 2)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(n, device="cuda", generator=g); y = torch.rand(n, device="cuda", generator=g)
cp = _capi.physics(2000., alpha_evm=0.05, has_evm=has_evm)
for _ in range(steps):
    o = abi.step(pm, cp, x, y, blocks=[], params_evm=pe if has_evm else None, want_resid=False)
print("ok", wl, n, float(o["loss_parts"][0]))
