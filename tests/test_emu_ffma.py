"""Kernel index logic + C-ABI orchestration checked WITHOUT a GPU: the FFMA kernel body compiled
for the host SIMT emulation (tests/emu) against the numpy oracle and the reference's golden vectors.
The CUDA build of the very same source is checked on the B200 by tests/test_gpu_*.py."""
import os

import numpy as np
import pytest

from nsfnet_b200 import _capi
from oracle import jet_numpy as J
from tests.emu import emu


def rel(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="module")
def lib():
    return emu.load()


@pytest.mark.parametrize("L,H,n,nb", [(3, 16, 70, 21), (6, 80, 133, 40), (4, 120, 50, 33), (2, 10, 37, 5), (1, 8, 9, 3)])
@pytest.mark.parametrize("has_evm", [False, True])
@pytest.mark.parametrize("reverse", [0, 1])
def test_step_matches_oracle(lib, L, H, n, nb, has_evm, reverse):
    lib.nsf_emu_set_reverse(reverse)
    try:
        rng = np.random.default_rng(L * 100 + H)
        md, ed = J.NetDesc(2, 3, L, H), J.NetDesc(2, 1, 4, 40)
        pm, pe = J.init_params(md, 1) * 1.5, J.init_params(ed, 2)
        x, y = rng.random(n), rng.random(n)
        xb, yb, ub, vb = rng.random(nb), rng.random(nb), rng.random(nb), rng.random(nb)
        w = 0.5 + rng.random(n)
        vtm = rng.random(n) * 0.02
        cs = 1.3 if has_evm else 1.0
        phys = J.Physics(Re=1000., alpha_b=10., alpha_evm=0.05, has_evm=has_evm, evm_trainable=has_evm, coord_scale=cs)
        r = J.step(pm, md, phys, x.astype(np.float32), y.astype(np.float32), xb.astype(np.float32), yb.astype(np.float32),
                   ub.astype(np.float32), vb.astype(np.float32), evm_flat=pe if has_evm else None,
                   evm_desc=ed if has_evm else None, w=w.astype(np.float32), vis_t_minus=vtm.astype(np.float32) if has_evm else None)
        cp = _capi.physics(1000., alpha_evm=0.05, has_evm=has_evm, evm_trainable=has_evm, coord_scale=cs)
        o = emu.run_step(lib, (2, 3, L, H), pm, cp, x, y, blocks=[(xb, yb, ub, vb, None, 10. / nb, 10. / nb, 0.)],
                         evm_desc=(2, 1, 4, 40) if has_evm else None, params_evm=pe if has_evm else None, w=w,
                         vtm_in=vtm if has_evm else None)
        assert rel(o["grad_main"], r.grad_main) < 2e-6
        for k in range(4 if has_evm else 3):
            assert rel(o["resid"][k], r.eq[k]) < 5e-6
            assert abs(o["loss_parts"][k] / n - r.loss_eq[k]) < 3e-6 * r.loss_eq[k]
        assert o["loss_parts"][5] == n
        assert abs((o["loss_parts"][6] + o["loss_parts"][7]) / nb - r.loss_b) < 3e-6 * r.loss_b
        if has_evm:
            assert rel(o["grad_evm"], r.grad_evm) < 2e-6
            assert rel(o["e"], r.e) < 2e-6
            assert rel(o["vis_t"], r.vis_t) < 1e-6
            assert rel(o["vtm_out"], r.vis_t_minus_next) < 2e-6
    finally:
        lib.nsf_emu_set_reverse(0)


def test_golden_ev_through_emulated_abi(lib, golden_dir):
    """The reference's own outputs (ev-NSFnet, lagged viscosity over 3 steps) through the C ABI."""
    g = np.load(os.path.join(golden_dir, "ev_re2000_lag.npz"))
    xb, yb, ub, vb = J.cavity_boundary(int(g["n_side"]))
    nb = xb.size
    vtm = g["vis_t_minus_init"]
    for k in range(int(g["steps"])):
        cp = _capi.physics(float(g["Re"]), alpha_evm=float(g["alpha_evm"]), has_evm=True, evm_trainable=False)
        o = emu.run_step(lib, (2, 3, 6, 80), g[f"params_main_{k}"], cp, g["xf"], g["yf"],
                         blocks=[(xb, yb, ub, vb, None, 10. / nb, 10. / nb, 0.)], evm_desc=(2, 1, 4, 40),
                         params_evm=g[f"params_evm_{k}"], vtm_in=vtm)
        n = g["xf"].size
        assert rel(o["grad_main"], g[f"grad_main_{k}"]) < 5e-6
        assert rel(o["vis_t"], g[f"vis_t_{k}"]) < 1e-6
        for i in range(4):
            assert rel(o["resid"][i], g[f"eq{i+1}_{k}"]) < 5e-6
        lp = o["loss_parts"]
        loss = 10. * (lp[6] + lp[7]) / nb + (lp[0] + lp[1] + lp[2] + 0.1 * lp[3]) / n
        assert abs(loss - float(g[f"loss_{k}"])) < 5e-6 * float(g[f"loss_{k}"])
        vtm = o["vtm_out"]


def test_golden_supervised_sdf(lib, golden_dir):
    g = np.load(os.path.join(golden_dir, "ev_re3000_scale_sup.npz"))
    xb, yb, ub, vb = J.cavity_boundary(int(g["n_side"]))
    nb = xb.size
    s = g["sup"]
    ns, nps = s.shape[0], int(np.isfinite(s[:, 4]).sum())
    a_s = float(g["alpha_s"])
    cp = _capi.physics(float(g["Re"]), alpha_evm=float(g["alpha_evm"]), has_evm=True, coord_scale=float(g["coord_scale"]))
    o = emu.run_step(lib, (2, 3, 6, 80), g["params_main_0"], cp, g["xf"], g["yf"],
                     blocks=[(xb, yb, ub, vb, None, 10. / nb, 10. / nb, 0.),
                             (s[:, 0], s[:, 1], s[:, 2], s[:, 3], s[:, 4], a_s / ns, a_s / ns, a_s / nps)],
                     evm_desc=(2, 1, 4, 40), params_evm=g["params_evm_0"], vtm_in=g["vis_t_minus_init"])
    assert rel(o["grad_main"], g["grad_main_0"]) < 5e-6
    lp = o["loss_parts"]
    assert lp[13] == nps
    loss_s = (lp[10] + lp[11]) / ns + lp[12] / nps
    assert abs(loss_s - float(g["loss_s_0"])) < 5e-6 * float(g["loss_s_0"])

    g = np.load(os.path.join(golden_dir, "ev_re5000_sdf_unfrozen.npz"))
    xb, yb, ub, vb = J.cavity_boundary(int(g["n_side"]))
    nb = xb.size
    cp = _capi.physics(float(g["Re"]), alpha_evm=float(g["alpha_evm"]), has_evm=True, evm_trainable=True)
    o = emu.run_step(lib, (2, 3, 6, 80), g["params_main_0"], cp, g["xf"], g["yf"], blocks=[(xb, yb, ub, vb, None, 10. / nb, 10. / nb, 0.)],
                     evm_desc=(2, 1, 4, 40), params_evm=g["params_evm_0"], w=g["w"], vtm_in=g["vis_t_minus_init"])
    assert rel(o["grad_main"], g["grad_main_0"]) < 5e-6
    assert rel(o["grad_evm"], g["grad_evm_0"]) < 5e-6


def test_golden_ns(lib, golden_dir):
    for name in ("ns_re100_init", "ns_re1000_x2p5"):
        g = np.load(os.path.join(golden_dir, name + ".npz"))
        xb, yb, ub, vb = J.cavity_boundary(int(g["n_side"]))
        nb = xb.size
        o = emu.run_step(lib, (2, 3, 4, 120), g["params"], _capi.physics(float(g["Re"])), g["xf"], g["yf"],
                         blocks=[(xb, yb, ub, vb, None, 10. / nb, 10. / nb, 0.)])
        assert rel(o["grad_main"], g["grad"]) < 5e-6
        for i in range(3):
            assert rel(o["resid"][i], g[f"eq{i+1}"]) < 5e-6


def test_empty_and_errors(lib):
    md = J.NetDesc(2, 3, 2, 8)
    pm = J.init_params(md, 0)
    o = emu.run_step(lib, (2, 3, 2, 8), pm, _capi.physics(100.), np.zeros(0), np.zeros(0), blocks=[])
    assert np.all(o["grad_main"] == 0) and np.all(o["loss_parts"] == 0)
    with pytest.raises(_capi.NsfError):
        _capi.Context(lib, 0, (3, 3, 2, 8))           # n_in must be 2
    with pytest.raises(_capi.NsfError):
        _capi.Context(lib, 0, (2, 3, 2, 200))         # hidden too large
    with pytest.raises(_capi.NsfError):
        emu.run_step(lib, (2, 3, 2, 8), pm, _capi.physics(100., has_evm=True), np.zeros(4), np.zeros(4))  # EVM flag without EVM net
