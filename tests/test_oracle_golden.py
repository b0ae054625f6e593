"""Pin the CPU oracle against golden vectors produced by the reference's own code.

The goldens are fp32 results of /root/reference (tests/golden/make_golden.py).  The numpy
oracle runs in fp64, so agreement is limited by the reference's own fp32 noise (BASELINE.md 2:
loss 1e-7, residuals 1e-6, grads 1e-7..1e-6 rel-L2 at init)."""
import glob
import os

import numpy as np
import pytest

from oracle import jet_numpy as J

MAIN_NS = J.NetDesc(2, 3, 4, 120)
MAIN_EV = J.NetDesc(2, 3, 6, 80)
EVM = J.NetDesc(2, 1, 4, 40)


def rel(a, b):
    a = np.asarray(a, np.float64).reshape(-1); b = np.asarray(b, np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.mark.parametrize("name", ["ns_re100_init", "ns_re1000_x2p5"])
def test_ns_step_matches_reference(golden_dir, name):
    g = load(golden_dir, name)
    xb, yb, ub, vb = J.cavity_boundary(int(g["n_side"]))
    phys = J.Physics(Re=float(g["Re"]), alpha_b=float(g["alpha_b"]), alpha_e=float(g["alpha_e"]))
    r = J.step(g["params"], MAIN_NS, phys, g["xf"], g["yf"], xb, yb, ub, vb)
    assert abs(r.loss - float(g["loss"])) <= 2e-6 * abs(float(g["loss"]))
    assert abs(r.loss_b - float(g["loss_b"])) <= 2e-6 * abs(float(g["loss_b"]))
    for i in range(3):
        assert rel(r.eq[i], g[f"eq{i+1}"]) < 5e-6
        assert abs(r.loss_eq[i] - g["loss_eq"][i]) <= 3e-6 * abs(g["loss_eq"][i])
    assert rel(r.grad_main, g["grad"]) < 5e-6
    # value forward (neural_net_u)
    out, _ = J.mlp_forward(J.unpack(g["params"], MAIN_NS), np.concatenate([g["xf"], g["yf"]], 1).astype(np.float64))
    assert rel(out, g["uvp"]) < 2e-6


@pytest.mark.parametrize("name", ["ev_re2000_lag", "ev_re5000_sdf_unfrozen", "ev_re3000_scale_sup"])
def test_ev_steps_match_reference(golden_dir, name):
    g = load(golden_dir, name)
    xb, yb, ub, vb = J.cavity_boundary(int(g["n_side"]))
    phys = J.Physics(Re=float(g["Re"]), alpha_b=float(g["alpha_b"]), alpha_e=float(g["alpha_e"]),
                     alpha_evm=float(g["alpha_evm"]), alpha_s=float(g["alpha_s"]), coord_scale=float(g["coord_scale"]),
                     has_evm=True, evm_trainable=bool(int(g["unfreeze"])))
    w = g["w"] if "w" in g.files else None
    sup = None
    if "sup" in g.files:
        s = g["sup"]
        sup = (s[:, 0], s[:, 1], s[:, 2], s[:, 3], s[:, 4])
    vtm = g["vis_t_minus_init"]
    for k in range(int(g["steps"])):
        r = J.step(g[f"params_main_{k}"], MAIN_EV, phys, g["xf"], g["yf"], xb, yb, ub, vb,
                   evm_flat=g[f"params_evm_{k}"], evm_desc=EVM, w=w, vis_t_minus=vtm, sup=sup)
        assert rel(r.vis_t, g[f"vis_t_{k}"]) < 1e-6, "lagged entropy viscosity (ev :327-334)"
        assert rel(r.e, g[f"e_{k}"]) < 2e-6
        for i in range(4):
            assert rel(r.eq[i], g[f"eq{i+1}_{k}"]) < 5e-6, (k, i)
            assert abs(r.loss_eq[i] - g[f"loss_eq_{k}"][i]) <= 5e-6 * abs(g[f"loss_eq_{k}"][i])
        assert abs(r.loss - float(g[f"loss_{k}"])) <= 3e-6 * abs(float(g[f"loss_{k}"]))
        assert abs(r.loss_s - float(g[f"loss_s_{k}"])) <= 3e-6 * abs(float(g[f"loss_s_{k}"])) + 1e-12
        assert rel(r.grad_main, g[f"grad_main_{k}"]) < 5e-6, k
        if phys.evm_trainable:
            assert rel(r.grad_evm, g[f"grad_evm_{k}"]) < 5e-6, k
        else:
            assert np.all(g[f"grad_evm_{k}"] == 0)
        # hand the lag state on exactly as the reference does (ev :334), with the weights of step k
        vtm = r.vis_t_minus_next.astype(np.float32)


def test_fp32_oracle_close_to_fp64(golden_dir):
    g = load(golden_dir, "ns_re100_init")
    xb, yb, ub, vb = J.cavity_boundary(int(g["n_side"]))
    phys = J.Physics(Re=100.0)
    r64 = J.step(g["params"], MAIN_NS, phys, g["xf"], g["yf"], xb, yb, ub, vb, dtype=np.float64)
    r32 = J.step(g["params"], MAIN_NS, phys, g["xf"], g["yf"], xb, yb, ub, vb, dtype=np.float32)
    assert rel(r32.grad_main, r64.grad_main) < 1e-5


def test_autograd_port_matches_reference(golden_dir):
    """The torch port (what bench.py times as cpu_baseline) reproduces the reference bit-for-bit-ish."""
    import torch
    from oracle.autograd_port import RefSolver
    g = load(golden_dir, "ev_re2000_lag")
    s = RefSolver(2000, 6, 80, 4, 40, alpha_evm=0.05)
    s.net.load_flat(g["params_main_0"]); s.net_1.load_flat(g["params_evm_0"])
    s.set_boundary_data(J.cavity_boundary(int(g["n_side"])))
    s.set_eq_training_data((g["xf"].astype(np.float64), g["yf"].astype(np.float64)))
    assert rel(s.vis_t_minus, g["vis_t_minus_init"]) < 1e-6
    for k in range(int(g["steps"])):
        loss = s.adam_step()
        assert abs(loss - float(g[f"loss_{k}"])) <= 2e-6 * abs(loss)
    gn = load(golden_dir, "ns_re1000_x2p5")
    s = RefSolver(1000, 4, 120)
    s.net.load_flat(gn["params"])
    s.set_boundary_data(J.cavity_boundary(int(gn["n_side"])))
    s.set_eq_training_data((gn["xf"].astype(np.float64), gn["yf"].astype(np.float64)))
    loss = s.loss_fn(); s.opt.zero_grad(); loss.backward()
    assert rel(s.net.flat_grad(), gn["grad"]) < 2e-6
    assert abs(float(loss) - float(gn["loss"])) <= 2e-6 * float(gn["loss"])


def test_goldens_present(golden_dir):
    names = {os.path.basename(p) for p in glob.glob(os.path.join(golden_dir, "*.npz"))}
    for n in ["ns_re100_init", "ns_re1000_x2p5", "ev_re2000_lag", "ev_re5000_sdf_unfrozen", "ev_re3000_scale_sup",
              "curve_ns_re100", "curve_ns_re1000", "curve_ev_re2000"]:
        assert n + ".npz" in names


# ---- round 2: trained-weight goldens with an fp64 run of the reference graph as truth (tests/golden/make_golden_r2.py) --------
@pytest.mark.parametrize("name", ["trained_ns_re100", "trained_ns_re1000"])
def test_oracle_matches_reference_fp64_at_trained_weights_ns(golden_dir, name):
    """The numpy oracle (fp64) against the reference's own fp64 evaluation after 3000 Adam steps: the restated math is the
    reference's math where it matters most (large second derivatives, cancelling gradient terms)."""
    g = load(golden_dir, name)
    xb, yb, ub, vb = [a.astype(np.float32) for a in J.cavity_boundary(int(g["n_side"]))]   # the truth was fed the fp32-valued inputs
    r = J.step(g["params"], MAIN_NS, J.Physics(Re=float(g["Re"])), g["xf"], g["yf"], xb, yb, ub, vb)
    assert rel(r.grad_main, g["grad_f64"]) < 1e-9
    for i in range(3):
        assert rel(r.eq[i], g[f"eq{i+1}_f64"]) < 1e-9
    assert abs(r.loss - float(g["loss_f64"])) <= 1e-10 * float(g["loss_f64"])
    # and the distance the reference's fp32 path itself keeps from that truth (the yardstick of tests/test_gpu_trained.py)
    assert 1e-7 < rel(g["grad_f32"], g["grad_f64"]) < 1e-4


def test_oracle_matches_reference_fp64_at_trained_weights_ev(golden_dir):
    g = load(golden_dir, "trained_ev_re2000")
    xb, yb, ub, vb = [a.astype(np.float32) for a in J.cavity_boundary(int(g["n_side"]))]
    phys = J.Physics(Re=float(g["Re"]), alpha_evm=float(g["alpha_evm"]), has_evm=True)
    r = J.step(g["params_main"], MAIN_EV, phys, g["xf"], g["yf"], xb, yb, ub, vb, evm_flat=g["params_evm"], evm_desc=EVM,
               vis_t_minus=g["vis_t_minus"])
    # the reference rounds vis_t = min(20/Re, vis_t_minus) to fp32 even when its nets run in fp64 (ev :327-331): 1e-8 apart
    assert rel(r.grad_main, g["grad_f64"]) < 1e-7
    for i in range(4):
        assert rel(r.eq[i], g[f"eq{i+1}_f64"]) < 1e-7
    assert rel(r.e, g["e_f64"]) < 1e-12
    assert abs(r.loss - float(g["loss_f64"])) <= 1e-8 * float(g["loss_f64"])


def test_shipped_boundary_set_and_sdf_weights(golden_dir):
    """N_b = 2052 from the reference's DataLoader.loading_boundary_data(), SDF weights from its cKDTree: the oracle's and the
    package's boundary sets are the same points, and the oracle reproduces the reference's fp32 step on them."""
    from nsfnet_b200.cavity_data import cavity_boundary
    g = load(golden_dir, "ev_re5000_nb2052_sdf")
    for mine in (J.cavity_boundary(513), cavity_boundary(513)):
        for a, b in zip(mine, (g["xb"], g["yb"], g["ub"], g["vb"])):
            assert np.array_equal(np.asarray(a, np.float32).reshape(-1), b)
    phys = J.Physics(Re=float(g["Re"]), alpha_evm=float(g["alpha_evm"]), has_evm=True)
    r = J.step(g["params_main"], MAIN_EV, phys, g["xf"], g["yf"], g["xb"], g["yb"], g["ub"], g["vb"], evm_flat=g["params_evm"],
               evm_desc=EVM, w=g["w"], vis_t_minus=g["vis_t_minus"])
    assert rel(r.grad_main, g["grad_f32"]) < 5e-6
    for i in range(4):
        assert rel(r.eq[i], g[f"eq{i+1}_f32"]) < 5e-6
    assert abs(r.loss - float(g["loss_f32"])) <= 3e-6 * float(g["loss_f32"])
    # the package's SDF weights (vectorised host layer) equal the reference's cKDTree result
    from nsfnet_b200.cavity_data import sdf_weights
    w = sdf_weights(np.stack([g["xf"], g["yf"]], 1).astype(np.float64), np.stack([g["xb"], g["yb"]], 1).astype(np.float64), 0.2, 5.0)
    assert rel(w, g["w"]) < 1e-6


def test_full_size_curves_recorded(golden_dir):
    for name in ("curve_full_ns_re100", "curve_full_ns_re1000"):
        g = load(golden_dir, name)
        assert int(g["steps"]) == 5000 and g["xf"].size == 10000 and g["curve"].size == 50 and g["curve_fp64"].size == 50
        assert g["curve"][0] == pytest.approx(g["curve_fp64"][0], rel=1e-6)
