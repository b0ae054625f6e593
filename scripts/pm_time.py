"""Kernel-only time of the collocation jet step (nsf_set_timing) for one workload: python scripts/pm_time.py ev|ns [n] [path]."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from nsfnet_b200 import _capi
from oracle import jet_numpy as J
from tests import gpu_util as gu

wl = sys.argv[1] if len(sys.argv) > 1 else "ev"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
path = int(sys.argv[3]) if len(sys.argv) > 3 else 3
L, H = (6, 80) if wl == "ev" else (4, 120)
has_evm = wl == "ev"
abi = gu.Abi((2, 3, L, H), (2, 1, 4, 40) if has_evm else None, path=path)
pm, pe = J.init_params(J.NetDesc(2, 3, L, H), 1), J.init_params(J.NetDesc(2, 1, 4, 40), 2)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(n, device="cuda", generator=g); y = torch.rand(n, device="cuda", generator=g)
cp = _capi.physics(2000., alpha_evm=0.05, has_evm=has_evm)
abi.ctx.set_timing(True)
ms = []
for i in range(12):
    abi.step(pm, cp, x, y, blocks=[], params_evm=pe if has_evm else None, want_resid=False)
    if i >= 2:
        ms.append(abi.ctx.last_kernel_ms())
import os
print(f"{wl} n={n} path={path} lib={os.path.basename(os.environ.get('NSF_B200_LIB', 'product'))}: kernel {np.median(ms):.3f} ms (min {min(ms):.3f})  -> {n / np.median(ms) / 1e3:.4g} pts/s")
