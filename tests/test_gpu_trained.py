"""Parity at TRAINED weights, judged against an fp64 run of the reference graph (SURVEY 7 hard part 3, 8c).

After a few thousand Adam steps the residuals are small sums of large cancelling terms (max |u_xx| ~ 40, residual rms ~ 3e-2)
and the gradient is a sum of cancelling per-point terms: the reference's OWN fp32 autograd path is then 1e-6 .. 1.4e-5 (rel-L2)
away from the same graph evaluated in fp64, so two fp32 evaluation orders cannot be asked to agree to 1e-5 with each other.
tests/golden/trained_*.npz (tests/golden/make_golden_r2.py) hold, for the state the UNMODIFIED reference reached after 3000 / 2000
Adam steps (N_f = 4000, 516 boundary points), the reference's fp32 outputs and its fp64 outputs on the same fp32-valued inputs.
The bar for every kernel path:   err(kernel, fp64) <= max(1e-5, 2 * err(reference fp32, fp64))
for residuals, loss and gradient (norm-wise), plus a floored element-wise check: no gradient entry further than
max(1e-3, 2 x reference) from the truth, relative to |truth| + 1 % of the gradient's rms (the maximum over 38 000 entries is a
4-sigma statistic of the error, i.e. ~4 x the rel-L2 figure scaled by rms / floor; 1e-3 is that for the 1e-5 norm-wise bar).  The measured triples go to
gpurun_out/r2_parity_trained.txt (committed as profiles/r2_parity_trained.txt).
"""
import os

import numpy as np
import pytest

from nsfnet_b200 import _capi
from oracle import jet_numpy as J

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH_NAME = {1: "ffma", 3: "tcgen05 points-on-M"}


def _bar(err_kernel, err_ref):
    return err_kernel <= max(1e-5, 2.0 * err_ref)


def _elem(a, truth):
    """worst element error relative to |truth| floored at 1 % of the tensor's rms"""
    a = np.asarray(a, np.float64).ravel(); t = np.asarray(truth, np.float64).ravel()
    floor = 0.01 * np.sqrt(np.mean(t * t))
    return float(np.max(np.abs(a - t) / (np.abs(t) + floor)))


def _report(lines):
    print("\n".join(lines))
    try:                                  # the record is a by-product: a read-only checkout must not fail the parity test
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "r2_parity_trained.txt"), "a") as f:
            f.write("\n".join(lines) + "\n")
    except OSError:
        pass


def _check(name, path, o, g, n_eq, n, nb):
    from tests import gpu_util as gu
    lines = [f"{name}  path {path} ({PATH_NAME[path]}):   kernel-vs-fp64 | reference-fp32-vs-fp64 | bar = max(1e-5, 2 x reference)"]
    ok = True
    kg, rg = gu.rel(o["grad_main"], g["grad_f64"]), gu.rel(g["grad_f32"], g["grad_f64"])
    lines.append(f"  gradient rel-L2      {kg:.3e} | {rg:.3e} | {max(1e-5, 2 * rg):.3e}")
    ok &= _bar(kg, rg)
    ke, re_ = _elem(o["grad_main"], g["grad_f64"]), _elem(g["grad_f32"], g["grad_f64"])
    lines.append(f"  gradient worst elem  {ke:.3e} | {re_:.3e} | {max(1e-3, 2 * re_):.3e}   (relative to |truth| + 1 % rms)")
    ok &= ke <= max(1e-3, 2.0 * re_)
    for i in range(n_eq):
        kr, rr = gu.rel(o["resid"][i], g[f"eq{i+1}_f64"]), gu.rel(g[f"eq{i+1}_f32"], g[f"eq{i+1}_f64"])
        lines.append(f"  eq{i+1} residual rel-L2 {kr:.3e} | {rr:.3e} | {max(1e-5, 2 * rr):.3e}")
        ok &= _bar(kr, rr)
    lp = o["loss_parts"]
    w4 = 0.1 * lp[3] if n_eq == 4 else 0.0
    loss = 10.0 * (lp[6] + lp[7]) / nb + (lp[0] + lp[1] + lp[2] + w4) / n
    kl = abs(loss - float(g["loss_f64"])) / float(g["loss_f64"]); rl = abs(float(g["loss_f32"]) - float(g["loss_f64"])) / float(g["loss_f64"])
    lines.append(f"  loss                 {kl:.3e} | {rl:.3e} | {max(1e-5, 2 * rl):.3e}")
    ok &= _bar(kl, rl)
    _report(lines)
    return ok


@pytest.mark.parametrize("name,path", [("trained_ns_re100", 3), ("trained_ns_re1000", 3), ("trained_ns_re100", 1), ("trained_ns_re1000", 1)])
def test_trained_ns(golden_dir, name, path):
    from tests import gpu_util as gu
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    xb, yb, ub, vb = J.cavity_boundary(int(g["n_side"]))
    nb, n = xb.size, g["xf"].size
    abi = gu.Abi((2, 3, 4, 120), path=path)
    o = abi.step(g["params"], _capi.physics(float(g["Re"])), g["xf"], g["yf"], blocks=[(xb, yb, ub, vb, None, 10. / nb, 10. / nb, 0.)])
    assert o["info"]["path"] == path
    assert _check(name, path, o, g, 3, n, nb)


@pytest.mark.parametrize("path", [3, 1])
def test_trained_ev(golden_dir, path):
    from tests import gpu_util as gu
    g = np.load(os.path.join(golden_dir, "trained_ev_re2000.npz"))
    xb, yb, ub, vb = J.cavity_boundary(int(g["n_side"]))
    nb, n = xb.size, g["xf"].size
    abi = gu.Abi((2, 3, 6, 80), (2, 1, 4, 40), path=path)
    cp = _capi.physics(float(g["Re"]), alpha_evm=float(g["alpha_evm"]), has_evm=True)
    o = abi.step(g["params_main"], cp, g["xf"], g["yf"], blocks=[(xb, yb, ub, vb, None, 10. / nb, 10. / nb, 0.)],
                 params_evm=g["params_evm"], vtm_in=g["vis_t_minus"])
    assert o["info"]["path"] == path
    assert _check("trained_ev_re2000", path, o, g, 4, n, nb)
    assert gu.rel(o["e"], g["e_f64"]) <= max(1e-5, 2 * gu.rel(g["e_f32"], g["e_f64"]))


@pytest.mark.parametrize("name,path", [("trained_ns_re1000", 3), ("trained_ev_re2000", 3), ("trained_ns_re1000", 1)])
def test_trained_weights_at_scale(golden_dir, name, path):
    """The goldens above hold 4000 points: one tile per CTA.  Here the trained weights meet 120 000 fresh points (25 / 50 tiles per CTA,
    so the per-CTA gradient rows are accumulated over many tiles); truth = the fp64 oracle, yardstick = the same oracle in fp32."""
    from tests import gpu_util as gu
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    ev = "params_main" in g.files
    rng = np.random.default_rng(5)
    n = 120000
    x, y = rng.random(n).astype(np.float32), rng.random(n).astype(np.float32)
    xb, yb, ub, vb = J.cavity_boundary(int(g["n_side"]))
    nb = xb.size
    md, ed = (J.NetDesc(2, 3, 6, 80), J.NetDesc(2, 1, 4, 40)) if ev else (J.NetDesc(2, 3, 4, 120), None)
    pm = g["params_main"] if ev else g["params"]
    pe = g["params_evm"] if ev else None
    vtm = (rng.random(n) * 0.02).astype(np.float32) if ev else None
    phys = J.Physics(Re=float(g["Re"]), alpha_b=10., alpha_evm=float(g["alpha_evm"]) if ev else 0.03, has_evm=ev)
    kw = dict(evm_flat=pe, evm_desc=ed, vis_t_minus=vtm)
    r64 = J.step(pm, md, phys, x, y, xb, yb, ub, vb, **kw)
    r32 = J.step(pm, md, phys, x, y, xb, yb, ub, vb, dtype=np.float32, **kw)
    abi = gu.Abi(tuple([2, 3, md.n_hidden_layers, md.hidden]), (2, 1, 4, 40) if ev else None, path=path)
    cp = _capi.physics(float(g["Re"]), alpha_evm=float(g["alpha_evm"]) if ev else 0.03, has_evm=ev)
    o = abi.step(pm, cp, x, y, blocks=[(xb, yb, ub, vb, None, 10. / nb, 10. / nb, 0.)], params_evm=pe, vtm_in=vtm)
    assert o["info"]["path"] == path
    kg, rg = gu.rel(o["grad_main"], r64.grad_main), gu.rel(r32.grad_main, r64.grad_main)
    lines = [f"{name} at scale ({n} points)  path {path} ({PATH_NAME[path]}): gradient rel-L2 kernel-vs-fp64 {kg:.3e} | fp32 oracle-vs-fp64 {rg:.3e}"]
    ok = _bar(kg, rg)
    for i in range(4 if ev else 3):
        kr, rr = gu.rel(o["resid"][i], r64.eq[i]), gu.rel(r32.eq[i], r64.eq[i])
        lines.append(f"  eq{i+1} residual rel-L2 {kr:.3e} | {rr:.3e}")
        ok &= _bar(kr, rr)
    _report(lines)
    assert ok


@pytest.mark.parametrize("path", [3, 1])
def test_golden_shipped_boundary_set_and_sdf_weights(golden_dir, path):
    """N_b = 2052: the boundary points come from the reference's own DataLoader.loading_boundary_data(), the SDF weights from
    its cKDTree over them (ev-NSFnet/cavity_data.py:47-94,118-130); production.yaml's nets, 20 000 collocation points."""
    from tests import gpu_util as gu
    from nsfnet_b200.cavity_data import cavity_boundary
    g = np.load(os.path.join(golden_dir, "ev_re5000_nb2052_sdf.npz"))
    nb, n = g["xb"].size, g["xf"].size
    assert nb == 2052
    mine = cavity_boundary(513)                      # the package's own boundary set is the reference's, bit for bit in fp32
    for a, b in zip(mine, (g["xb"], g["yb"], g["ub"], g["vb"])):
        assert np.array_equal(np.asarray(a, np.float32).reshape(-1), b)
    abi = gu.Abi((2, 3, 6, 80), (2, 1, 4, 40), path=path)
    cp = _capi.physics(float(g["Re"]), alpha_evm=float(g["alpha_evm"]), has_evm=True)
    o = abi.step(g["params_main"], cp, g["xf"], g["yf"], blocks=[(g["xb"], g["yb"], g["ub"], g["vb"], None, 10. / nb, 10. / nb, 0.)],
                 params_evm=g["params_evm"], w=g["w"], vtm_in=g["vis_t_minus"])
    assert gu.rel(o["grad_main"], g["grad_f32"]) < 1e-5
    for i in range(4):
        assert gu.rel(o["resid"][i], g[f"eq{i+1}_f32"]) < 1e-5
    lp = o["loss_parts"]
    loss = 10.0 * (lp[6] + lp[7]) / nb + (lp[0] + lp[1] + lp[2] + 0.1 * lp[3]) / n
    assert abs(loss - float(g["loss_f32"])) < 1e-5 * float(g["loss_f32"])
    for i in range(4):
        assert abs(lp[i] / n - g["loss_eq_f32"][i]) < 1e-5 * g["loss_eq_f32"][i]


def _run_ns_curve(g, steps):
    import torch
    from nsfnet_b200.nsfnet import PysicsInformedNeuralNetwork
    from nsfnet_b200.cavity_data import cavity_boundary
    P = PysicsInformedNeuralNetwork(Re=float(g["Re"]), layers=4, hidden_size=120, N_f=g["xf"].size, bc_weight=10, eq_weight=1)
    off = 0
    flat = torch.as_tensor(g["params"]).cuda()
    with torch.no_grad():
        for p in P.net.parameters():
            p.copy_(flat[off:off + p.numel()].view(p.shape)); off += p.numel()
    P.verbose = False
    P.set_boundary_data(cavity_boundary(513)); P.set_eq_training_data((g["xf"], g["yf"]))
    P.opt.param_groups[0]["lr"] = float(g["lr"])
    curve = []
    for k in range(steps):                                   # the reference's loop body, NSFnet/pinn_solver.py:250-254
        loss, _ = P.fwd_computing_loss_2d()
        loss.backward(); P.opt.step(); P.opt.zero_grad()
        curve.append(loss.detach())
    return torch.stack(curve).double().cpu().numpy()


def _curve_report(lines):
    print("\n".join(lines))
    try:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "r2_curves.txt"), "a") as f:
            f.write("\n".join(lines) + "\n")
    except OSError:
        pass


@pytest.mark.parametrize("name", ["curve_early_ns_re100", "curve_early_ns_re1000"])
def test_loss_curves_track_reference_step_by_step(golden_dir, name):
    """BASELINE config 1: NSFnet 4x120, N_f = 10 000, the reference's 2052 boundary points, Adam lr 1e-3, EVERY one of the
    first 800 steps against the UNMODIFIED reference's own loop (tests/golden/make_golden_r2.py `early`).

    Adam at lr 1e-3 amplifies rounding differences by ~x30 per 100 steps: the reference run with 1 host thread instead of 4 (same
    fp32 program, another summation order) is 3e-7 away from itself at step 100, 5e-5 at step 200, 4e-3 at step 400 and 6 % at step
    800; its fp64 twin likewise.  So 'tracks the reference' cannot be a fixed tolerance; it is: at every step the kernel's curve is
    no further from the reference's than TWO legitimate evaluations of the reference are from each other,
        dev_kernel(t) <= max(2e-5, 3 x envelope(t + 25)),   envelope = running max of (1-thread fp32, fp64) deviations,
    (a factor 3 and 25 steps of slack on an exponentially growing envelope), and bit-level agreement at the start."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    T = int(g["steps"])
    curve = _run_ns_curve(g, T)
    ref = g["curve_fp32_t4"]
    dev = np.abs(curve - ref) / ref
    env = np.maximum.accumulate(np.maximum(np.abs(g["curve_fp32_t1"] - ref) / ref, np.abs(g["curve_fp64"] - ref) / ref))
    shifted = np.concatenate([env[25:], np.full(25, env[-1])])
    bound = np.maximum(2e-5, 3.0 * shifted)
    ks = [0, 50, 100, 150, 200, 300, 400, 600, 799]
    _curve_report([f"{name}: first {T} Adam steps, every step, N_f = {g['xf'].size}",
                   "  step                      " + " ".join(f"{k:9d}" for k in ks),
                   "  kernel vs reference       " + " ".join(f"{dev[k]:9.2e}" for k in ks),
                   "  reference vs itself (env) " + " ".join(f"{env[k]:9.2e}" for k in ks),
                   f"  worst ratio dev / bound {np.max(dev / bound):.3f} at step {int(np.argmax(dev / bound))}; loss at step {T - 1}: kernel {curve[-1]:.4e}, reference {ref[-1]:.4e} "
                   f"(1 thread {g['curve_fp32_t1'][-1]:.4e}, fp64 {g['curve_fp64'][-1]:.4e})"])
    assert dev[0] < 2e-6 and np.all(dev[:50] < 1e-5)
    assert np.all(dev <= bound), (int(np.argmax(dev / bound)), float(np.max(dev / bound)))


@pytest.mark.parametrize("name", ["curve_full_ns_re100", "curve_full_ns_re1000"])
def test_full_size_curves_track_reference(golden_dir, name):
    """The same run to 5000 steps (loss every 100 steps from the reference's own loop, and its fp64 twin).  Past step ~500 the
    trajectories of ANY two fp32 evaluations have separated (test above), so what is compared is what survives chaos: the loss
    level.  Per sample the kernel's loss stays within the band the reference's own fp32 / fp64 pair spans, widened by a factor 3
    (Adam spikes make single samples of the reference's own pair differ by up to 5x), and the median of the last 1000 steps
    within max(25 %, 2 x the reference's own fp32-vs-fp64 gap)."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    every = int(g["every"])
    curve = _run_ns_curve(g, int(g["steps"]))[::every]
    ref, c64 = g["curve"], g["curve_fp64"]
    err = np.abs(curve - ref) / ref
    env = np.maximum.accumulate(np.abs(c64 - ref) / ref)
    tail = slice(-10, None)
    lvl, lvl_ref, lvl64 = np.median(curve[tail]), np.median(ref[tail]), np.median(c64[tail])
    lines = [f"{name}: {int(g['steps'])} Adam steps, N_f = {g['xf'].size}, sampled every {every}",
             f"  max deviation from the reference curve  {err.max():.3e} (at sample {int(err.argmax())}); reference's own fp32-vs-fp64 envelope {env.max():.3e}",
             f"  first 10 samples: deviation {np.array2string(err[:10], precision=2)}",
             f"                    envelope  {np.array2string(env[:10], precision=2)}",
             f"  median loss of the last 1000 steps: kernel {lvl:.4e}, reference fp32 {lvl_ref:.4e}, reference fp64 {lvl64:.4e}"]
    _curve_report(lines)
    assert err[0] < 1e-5
    lo, hi = np.minimum(ref, c64), np.maximum(ref, c64)
    # smoothed band: the reference pair over a window of +-2 samples (a spike lands one sample earlier or later in another run)
    pad = lambda a, f: np.array([f(a[max(0, i - 2):i + 3]) for i in range(a.size)])
    assert np.all(curve <= 3.0 * pad(hi, np.max)) and np.all(curve >= pad(lo, np.min) / 3.0), (curve, lo, hi)
    assert abs(np.log(lvl / lvl_ref)) <= max(np.log(1.25), 2 * abs(np.log(lvl64 / lvl_ref)))
