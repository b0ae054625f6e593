"""Gradient of the trained goldens on a chosen kernel path -> gpurun_out/trained_grad_<name>_p<path>.npy (offline error analysis)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nsfnet_b200 import _capi
from oracle import jet_numpy as J
from tests import gpu_util as gu
gd = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for name, path in (("trained_ev_re2000", 3), ("trained_ev_re2000", 1), ("trained_ns_re1000", 3), ("trained_ns_re1000", 1)):
    g = np.load(os.path.join(gd, name + ".npz"))
    xb, yb, ub, vb = J.cavity_boundary(int(g["n_side"]))
    nb = xb.size
    if "ev" in name:
        abi = gu.Abi((2, 3, 6, 80), (2, 1, 4, 40), path=path)
        cp = _capi.physics(float(g["Re"]), alpha_evm=float(g["alpha_evm"]), has_evm=True)
        o = abi.step(g["params_main"], cp, g["xf"], g["yf"], blocks=[(xb, yb, ub, vb, None, 10. / nb, 10. / nb, 0.)], params_evm=g["params_evm"], vtm_in=g["vis_t_minus"])
    else:
        abi = gu.Abi((2, 3, 4, 120), path=path)
        o = abi.step(g["params"], _capi.physics(float(g["Re"])), g["xf"], g["yf"], blocks=[(xb, yb, ub, vb, None, 10. / nb, 10. / nb, 0.)])
    np.save(f"gpurun_out/trained_grad_{name}_p{path}{tag}.npy", o["grad_main"])
    print(name, path, gu.rel(o["grad_main"], g["grad_f64"]))
