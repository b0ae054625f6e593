"""Round-2 golden vectors, produced by the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden_r2.py trained     # parity at TRAINED weights, with an fp64 run as truth
    python tests/golden/make_golden_r2.py nb2052      # shipped boundary set (DataLoader.loading_boundary_data) + SDF weights
    python tests/golden/make_golden_r2.py curves      # config 1 at full size: N_f = 10 000, 5 000 Adam steps

trained:  the reference's own loop (NSFnet/pinn_solver.py:250-254, ev-NSFnet/pinn_solver.py:440-487) runs a few
          thousand Adam steps (N_f = 4000, 516 boundary points, lr 1e-3).  At the state reached,
          one more evaluation of ``fwd_computing_loss_2d`` + ``loss.backward()`` is recorded twice: in fp32 (the
          reference as shipped) and in fp64 (the same objects after ``net.double()``, fed the same fp32-valued points,
          weights and lagged viscosity).  SURVEY 7 hard part 3: at trained weights two fp32 evaluation orders cannot
          agree better than either agrees with fp64, so the GPU test asserts
          err(kernel, fp64) <= max(1e-5, 2 * err(reference fp32, fp64)).
"""
import copy
import os
import sys
import time
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, REF, flat_grads, flat_params, load_reference, make_ev  # noqa: E402


def ref_dataloader(variant="ev-NSFnet", **kw):
    """The reference's own DataLoader (cavity_data.py); tools.py imports matplotlib, which is stubbed."""
    import importlib.util
    mpl = types.ModuleType("matplotlib"); mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules.setdefault("matplotlib", mpl); sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)
    d = f"{REF}/{variant}"
    sys.path.insert(0, d)
    try:
        for m in ("tools", "cavity_data"):
            sys.modules.pop(m, None)
        spec = importlib.util.spec_from_file_location("ref_cavity_data_" + variant.replace("-", "_"), f"{d}/cavity_data.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(d)
    return mod.DataLoader(**kw)


def boundary_sub(n_side):
    """4 x n_side points with the reference's lid profile (cavity_data.py:49-72 at a smaller Nx)."""
    s = np.linspace(0.0, 1.0, n_side)
    lid = 1 - np.cosh(10 * (s - 0.5)) / np.cosh(5.0)
    z, o = np.zeros(n_side), np.ones(n_side)
    return (np.concatenate([s, s, z, o]).reshape(-1, 1), np.concatenate([z, o, s, s]).reshape(-1, 1),
            np.concatenate([z, lid, z, z]).reshape(-1, 1), np.zeros((4 * n_side, 1)))


def _record(P, n_eq, tag, d, dtype):
    """One evaluation of the reference's loss + backward at the current state -> d[...]."""
    loss, (loss_e, loss_b) = P.fwd_computing_loss_2d()
    P.opt.zero_grad(); loss.backward()
    to = lambda t: t.detach().numpy().reshape(-1).astype(dtype)
    d[f"loss_{tag}"] = np.float64(loss.detach())
    d[f"loss_e_{tag}"] = np.float64(loss_e.detach()); d[f"loss_b_{tag}"] = np.float64(loss_b.detach())
    d[f"loss_eq_{tag}"] = np.array([float(getattr(P, f"loss_eq{i+1}").detach()) for i in range(n_eq)], np.float64)
    for i in range(n_eq):
        d[f"eq{i+1}_{tag}"] = to(getattr(P, f"eq{i+1}_pred"))
    d[f"grad_{tag}"] = np.concatenate([(p.grad if p.grad is not None else torch.zeros_like(p)).numpy().reshape(-1)
                                       for p in P.net.parameters()]).astype(dtype)
    return d


def _to_double(P, has_evm):
    """The same solver object evaluated in fp64: nets cast, point tensors rebuilt from their fp32 values."""
    Q = copy.copy(P)
    Q.net = copy.deepcopy(P.net).double()
    if has_evm:
        Q.net_1 = copy.deepcopy(P.net_1).double()
    for k in ("x_b", "y_b", "u_b", "v_b"):
        setattr(Q, k, getattr(P, k).detach().double())
    Q.x_f = P.x_f.detach().double().requires_grad_(True)
    Q.y_f = P.y_f.detach().double().requires_grad_(True)
    Q.opt = torch.optim.Adam(list(Q.net.parameters()) + (list(Q.net_1.parameters()) if has_evm else []), lr=1e-3)
    return Q


def _rel(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def trained_ns(ns, name, Re, seed, n_f=4000, n_side=129, steps=3000):
    torch.manual_seed(seed)
    P = ns.PysicsInformedNeuralNetwork(Re=Re, layers=4, hidden_size=120, N_f=n_f, bc_weight=10, eq_weight=1)
    rng = np.random.default_rng(5000 + seed)
    xf, yf = rng.random((n_f, 1)), rng.random((n_f, 1))
    P.set_boundary_data(X=boundary_sub(n_side)); P.set_eq_training_data(X=(xf, yf))
    P.opt.param_groups[0]["lr"] = 1e-3
    t0 = time.time()
    for k in range(steps):                       # NSFnet/pinn_solver.py:250-254
        loss, _ = P.fwd_computing_loss_2d()
        loss.backward(); P.opt.step(); P.opt.zero_grad()
        if k % 500 == 0:
            print(name, k, float(loss), f"{time.time() - t0:.0f}s", flush=True)
    d = dict(kind="ns_trained", Re=Re, alpha_b=10.0, alpha_e=1.0, n_side=n_side, steps=steps, params=flat_params(P.net),
             xf=xf.astype(np.float32), yf=yf.astype(np.float32))
    _record(P, 3, "f32", d, np.float32)
    _record(_to_double(P, False), 3, "f64", d, np.float64)
    np.savez_compressed(f"{OUT}/{name}.npz", **d)
    print(name, "loss", d["loss_f32"], "ref fp32 vs fp64: loss", abs(d["loss_f32"] - d["loss_f64"]) / d["loss_f64"],
          "grad", _rel(d["grad_f32"], d["grad_f64"]), "eq1", _rel(d["eq1_f32"], d["eq1_f64"]), flush=True)


def trained_ev(ev, name, Re, seed, n_f=4000, n_side=129, steps=2000, alpha_evm=0.05):
    P = make_ev(ev, Re, alpha_evm, seed)
    rng = np.random.default_rng(6000 + seed)
    xf, yf = rng.random((n_f, 1)), rng.random((n_f, 1))
    P.set_boundary_data(X=boundary_sub(n_side)); P.set_eq_training_data(X=(xf, yf))
    P.log_interval = 10 ** 9
    P.print_log = lambda *a, **k: None
    P.save = lambda *a, **k: None
    P.opt.param_groups[0]["lr"] = 1e-3
    P.solve_Adam(P.fwd_computing_loss_2d, steps)      # the reference's own loop incl. freeze bookkeeping (ev :440-487)
    vtm = np.asarray(P.vis_t_minus, np.float32).reshape(-1).copy()     # lag state the NEXT evaluation consumes
    d = dict(kind="ev_trained", Re=Re, alpha_b=10.0, alpha_e=1.0, alpha_evm=alpha_evm, n_side=n_side, steps=steps,
             params_main=flat_params(P.net), params_evm=flat_params(P.net_1), xf=xf.astype(np.float32), yf=yf.astype(np.float32),
             vis_t_minus=vtm)
    Q = _to_double(P, True)
    Q.vis_t_minus = vtm.reshape(-1, 1).copy()
    _record(P, 4, "f32", d, np.float32)
    d["e_f32"] = P.evm.detach().numpy().reshape(-1).astype(np.float32)
    _record(Q, 4, "f64", d, np.float64)
    d["e_f64"] = Q.evm.detach().numpy().reshape(-1)
    np.savez_compressed(f"{OUT}/{name}.npz", **d)
    print(name, "loss", d["loss_f32"], "ref fp32 vs fp64: loss", abs(d["loss_f32"] - d["loss_f64"]) / d["loss_f64"],
          "grad", _rel(d["grad_f32"], d["grad_f64"]), "eq1", _rel(d["eq1_f32"], d["eq1_f64"]), "eq4", _rel(d["eq4_f32"], d["eq4_f64"]),
          flush=True)


def nb2052(ev, name="ev_re5000_nb2052_sdf", Re=5000, seed=3, n_f=20000):
    """The shipped boundary set and SDF weights from the reference's OWN DataLoader (cavity_data.py:47-94,118-130),
    production.yaml's network and physics, one evaluation with the EVM net frozen."""
    sdf = types.SimpleNamespace(enabled=True, min_weight=0.2, decay=5.0)
    dl = ref_dataloader("ev-NSFnet", N_f=n_f, N_b=1000, sort_training_points=False, sdf_weighting=sdf)
    xb, yb, ub, vb = dl.loading_boundary_data()
    rng = np.random.default_rng(7000 + seed)
    pts = rng.random((n_f, 2))
    dl._compute_sdf_weights(pts)                       # cKDTree over the 2052 boundary points
    w = dl.get_sdf_weights()
    P = make_ev(ev, Re, 0.05, seed)
    P.set_boundary_data(X=(xb, yb, ub, vb)); P.set_eq_training_data(X=(pts[:, 0:1], pts[:, 1:2]), weights=w)
    P.freeze_evm_net(0)
    vtm = np.asarray(P.vis_t_minus, np.float32).reshape(-1).copy()
    d = dict(kind="ev_nb2052", Re=Re, alpha_b=10.0, alpha_e=1.0, alpha_evm=0.05, params_main=flat_params(P.net),
             params_evm=flat_params(P.net_1), xf=pts[:, 0].astype(np.float32), yf=pts[:, 1].astype(np.float32), w=w,
             xb=xb.astype(np.float32).reshape(-1), yb=yb.astype(np.float32).reshape(-1), ub=ub.astype(np.float32).reshape(-1),
             vb=vb.astype(np.float32).reshape(-1), vis_t_minus=vtm)
    _record(P, 4, "f32", d, np.float32)
    d["e_f32"] = P.evm.detach().numpy().reshape(-1).astype(np.float32)
    np.savez_compressed(f"{OUT}/{name}.npz", **d)
    print(name, "N_b", xb.shape[0], "loss", d["loss_f32"], flush=True)


def curve_full(ns, name, Re, seed=0, n_f=10000, steps=5000, every=100, threads=4):
    """BASELINE config 1: NSFnet 4x120, N_f = 10 000, the reference's 2052 boundary points, Adam lr 1e-3;
    loss every 100 steps from the reference's own loop body, then the same loop in fp64 (oracle port) as the envelope."""
    torch.set_num_threads(threads)
    dl = ref_dataloader("NSFnet", N_f=n_f, N_b=1000)
    xb, yb, ub, vb = dl.loading_boundary_data()
    torch.manual_seed(seed)
    P = ns.PysicsInformedNeuralNetwork(Re=Re, layers=4, hidden_size=120, N_f=n_f, bc_weight=10, eq_weight=1)
    rng = np.random.default_rng(8000 + seed)
    xf, yf = rng.random((n_f, 1)), rng.random((n_f, 1))
    P.set_boundary_data(X=(xb, yb, ub, vb)); P.set_eq_training_data(X=(xf, yf))
    params = flat_params(P.net)
    P.opt.param_groups[0]["lr"] = 1e-3
    curve, t0 = [], time.time()
    for k in range(steps):
        loss, _ = P.fwd_computing_loss_2d()
        loss.backward(); P.opt.step(); P.opt.zero_grad()
        if k % every == 0:
            curve.append(float(loss))
            if k % 500 == 0:
                print(name, "fp32", k, float(loss), f"{time.time() - t0:.0f}s", flush=True)
    g = dict(kind="ns_curve_full", Re=Re, alpha_b=10.0, alpha_e=1.0, params=params, xf=xf.astype(np.float32), yf=yf.astype(np.float32),
             steps=steps, every=every, lr=1e-3, curve=np.array(curve, np.float64), ref_seconds_per_step=(time.time() - t0) / steps,
             ref_threads=threads)
    np.savez_compressed(f"{OUT}/{name}.npz", **g)
    # fp64 envelope: the same loop through the oracle port (bit-identical to the reference in fp32, tests/test_oracle_golden.py)
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from oracle.autograd_port import RefSolver
    s = RefSolver(float(Re), 4, 120, dtype=torch.float64, lr=1e-3)
    s.net.load_flat(params)
    s.set_boundary_data((xb, yb, ub, vb))
    s.set_eq_training_data((xf.astype(np.float32).astype(np.float64), yf.astype(np.float32).astype(np.float64)))
    c64, t0 = [], time.time()
    for k in range(steps):
        loss = s.loss_fn(); loss.backward(); s.opt.step(); s.opt.zero_grad()
        if k % every == 0:
            c64.append(float(loss.detach()))
            if k % 500 == 0:
                print(name, "fp64", k, c64[-1], f"{time.time() - t0:.0f}s", flush=True)
    g["curve_fp64"] = np.array(c64, np.float64)
    np.savez_compressed(f"{OUT}/{name}.npz", **g)
    print(name, "fp64-vs-fp32 max deviation", (np.abs(g["curve_fp64"] - g["curve"]) / g["curve"]).max(), flush=True)


def curve_early(ns, name, Re, seed=0, n_f=10000, steps=800):
    """The first 800 Adam steps of the same run, EVERY step: the reference's loop body in fp32 with 4 host threads, again with 1 thread
    (another summation order of the same fp32 program: the yardstick for 'two fp32 evaluations of this trajectory'), and in fp64."""
    dl = ref_dataloader("NSFnet", N_f=n_f, N_b=1000)
    xb, yb, ub, vb = dl.loading_boundary_data()
    rng = np.random.default_rng(8000 + seed)
    xf, yf = rng.random((n_f, 1)), rng.random((n_f, 1))
    curves, params = {}, None
    for tag, threads in (("fp32_t4", 4), ("fp32_t1", 1)):
        torch.set_num_threads(threads)
        torch.manual_seed(seed)
        P = ns.PysicsInformedNeuralNetwork(Re=Re, layers=4, hidden_size=120, N_f=n_f, bc_weight=10, eq_weight=1)
        P.set_boundary_data(X=(xb, yb, ub, vb)); P.set_eq_training_data(X=(xf, yf))
        if params is None:
            params = flat_params(P.net)
        P.opt.param_groups[0]["lr"] = 1e-3
        c, t0 = [], time.time()
        for k in range(steps):
            loss, _ = P.fwd_computing_loss_2d()
            loss.backward(); P.opt.step(); P.opt.zero_grad()
            c.append(float(loss))
        curves[tag] = np.array(c, np.float64)
        print(name, tag, f"{time.time() - t0:.0f}s", c[-1], flush=True)
    torch.set_num_threads(4)
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from oracle.autograd_port import RefSolver
    s = RefSolver(float(Re), 4, 120, dtype=torch.float64, lr=1e-3)
    s.net.load_flat(params)
    s.set_boundary_data((xb, yb, ub, vb))
    s.set_eq_training_data((xf.astype(np.float32).astype(np.float64), yf.astype(np.float32).astype(np.float64)))
    c = []
    for k in range(steps):
        loss = s.loss_fn(); loss.backward(); s.opt.step(); s.opt.zero_grad()
        c.append(float(loss.detach()))
    curves["fp64"] = np.array(c, np.float64)
    np.savez_compressed(f"{OUT}/{name}.npz", kind="ns_curve_early", Re=Re, params=params, xf=xf.astype(np.float32), yf=yf.astype(np.float32),
                        steps=steps, lr=1e-3, **{"curve_" + k: v for k, v in curves.items()})
    d1 = np.abs(curves["fp32_t1"] - curves["fp32_t4"]) / curves["fp32_t4"]; d2 = np.abs(curves["fp64"] - curves["fp32_t4"]) / curves["fp32_t4"]
    print(name, "1-thread vs 4-thread fp32: max", d1.max(), "at 100/200/400:", d1[100], d1[200], d1[400], "| fp64 vs fp32:", d2[100], d2[200], d2[400], flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "trained"
    ns, ev = load_reference()
    if what == "trained":
        torch.set_num_threads(int(sys.argv[2]) if len(sys.argv) > 2 else 4)
        trained_ns(ns, "trained_ns_re100", 100, 0)
        trained_ns(ns, "trained_ns_re1000", 1000, 1)
        trained_ev(ev, "trained_ev_re2000", 2000, 0)
    elif what == "nb2052":
        torch.set_num_threads(4)
        nb2052(ev)
    elif what == "early":
        Re = int(sys.argv[2])
        curve_early(ns, f"curve_early_ns_re{Re}", Re)
    elif what == "curves":
        Re = int(sys.argv[2])
        curve_full(ns, f"curve_full_ns_re{Re}", Re)
